"""Importable alias for the product package, whose directory name `ubpl-poseestimation_b200`
(fixed by the repo layout contract) is not a valid Python identifier.  `import ubpl_b200` loads
that directory as the package `ubpl_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ubpl-poseestimation_b200")
_spec = importlib.util.spec_from_file_location(
    "ubpl_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ubpl_b200"] = _mod
_spec.loader.exec_module(_mod)
