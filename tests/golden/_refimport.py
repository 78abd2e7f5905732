"""Import shim for the UNMODIFIED reference at /root/reference (golden generation only).

Only `tests/golden/make_golden.py` uses this, and only in the build container: the
reference is a pure-Python tree that cannot travel to the GPU box, so its outputs are
frozen into the .npz fixtures next to this file.  Nothing under tests/ that runs in
the normal suites imports this module.

The reference's driver modules import matplotlib / jsonpickle / gym / pyglet at module
level (for plotting and the Duckietown simulator).  None of that is on the planner hot
path, so permissive stub modules are injected before the import.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FRENET_REFERENCE_ROOT", "/root/reference")


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        mod = _Stub(self.__name__ + "." + name)
        setattr(self, name, mod)
        return mod

    def __call__(self, *args, **kwargs):
        return _Stub("stub_call")


_STUBBED = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.patches",
    "matplotlib.transforms", "matplotlib.lines", "matplotlib.gridspec",
    "matplotlib.collections", "jsonpickle", "gym", "gym.spaces", "gym_duckietown",
    "gym_duckietown.envs", "gym_duckietown.simulator", "pyglet", "pyglet.window",
]


def install():
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "lib", "planner")):
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the mount is read-only
    for name in _STUBBED:
        sys.modules.setdefault(name, _Stub(name))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
