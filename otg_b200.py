"""Import alias for the package directory.

The package lives in `optimal-trajectory-generation-in-the-duckietown-environment_b200/`
(the layout this repository is required to have); hyphens make that name unusable in an
`import` statement, so this module loads the directory as the package `otg_b200` and
replaces itself in `sys.modules`.  After `import otg_b200`, sub-modules import normally
(`from otg_b200.planner import TrajectoryPlannerV1`).
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "optimal-trajectory-generation-in-the-duckietown-environment_b200")
_spec = importlib.util.spec_from_file_location(
    "otg_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_pkg = importlib.util.module_from_spec(_spec)
sys.modules["otg_b200"] = _pkg
_spec.loader.exec_module(_pkg)
