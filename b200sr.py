"""Import shim: exposes the package directory `multi-image-super-resolution-for-medical-images_b200/`
(whose name is not a valid Python identifier) as the module `b200sr`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "multi-image-super-resolution-for-medical-images_b200")
_spec = importlib.util.spec_from_file_location("b200sr", os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b200sr"] = _mod
_spec.loader.exec_module(_mod)
