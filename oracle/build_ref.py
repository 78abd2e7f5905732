"""Recipe for `oracle/_ref/`: a verbatim, git-ignored copy of the reference's own pure-Python package, made at build time.

TEST INFRASTRUCTURE - the product (`otg_b200`) never imports anything from here.

The reference (pepes97/Optimal-trajectory-generation-in-the-Duckietown-environment) has no native code to compile: its hot path
is `lib/planner`, `lib/trajectory`, `lib/sensor`, ....  To time and cross-check the UNMODIFIED reference on
the GPU box's host cores (bench.py's `cpu_baseline` and `--impl reference` legs), `__graft_entry__.build()` calls `build_ref()`
here, in the build container where `/root/reference` is mounted.  It copies the `lib/` package (90 .py files, 0.6 MB - the test
drivers import the whole package through `lib/tests/__init__.py`) to `oracle/_ref/lib/`, unmodified, and writes a manifest with
one sha256 per file.  `oracle/_ref/` is listed in .gitignore (reference sources never enter this repository's history) but not
in .gpurunignore, so it travels to the GPU box like the built `.so` does.  When `/root/reference` is absent (the GPU box) the
function is a no-op and whatever `oracle/_ref/` already holds is used.
"""
from __future__ import annotations

import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("FRENET_REFERENCE_ROOT", "/root/reference")


def build_ref(verbose: bool = False) -> str | None:
    """Returns the path of oracle/_ref (or None when neither the reference nor an earlier copy exists)."""
    src = os.path.join(REFERENCE_ROOT, "lib")
    if not os.path.isdir(os.path.join(src, "planner")):
        return REF_DIR if os.path.isdir(os.path.join(REF_DIR, "lib", "planner")) else None
    dst = os.path.join(REF_DIR, "lib")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    lines = []
    for root, _, files in sorted(os.walk(dst)):
        for f in sorted(files):
            p = os.path.join(root, f)
            with open(p, "rb") as fh:
                lines.append(f"{hashlib.sha256(fh.read()).hexdigest()}  {os.path.relpath(p, REF_DIR)}")
    with open(os.path.join(REF_DIR, "MANIFEST.sha256"), "w") as fh:
        fh.write(f"# verbatim copy of {src} made by oracle/build_ref.py; {len(lines)} files\n" + "\n".join(lines) + "\n")
    if verbose:
        print(f"oracle/_ref: {len(lines)} files copied from {src}")
    return REF_DIR


if __name__ == "__main__":
    print(build_ref(verbose=True))
