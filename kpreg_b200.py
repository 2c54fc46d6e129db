"""Import alias: ``import kpreg_b200`` loads the package directory
``boosting-fine-grained-feature-fusion-in-3d-point-cloud-registration_b200/`` (whose name is not a
valid Python identifier) under the module name ``kpreg_b200``."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "boosting-fine-grained-feature-fusion-in-3d-point-cloud-registration_b200")
_spec = importlib.util.spec_from_file_location(
    "kpreg_b200", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["kpreg_b200"] = _mod
_spec.loader.exec_module(_mod)
