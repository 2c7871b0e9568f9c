"""Import shim: the package lives in ``x-detector_b200/`` (the name the project layout fixes),
which is not a valid Python identifier.  ``import xdet_b200`` loads that directory as the
package ``xdet_b200`` (sub-modules resolve inside it: ``xdet_b200.ops``, ``xdet_b200.net`` ...)."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "x-detector_b200")
_spec = importlib.util.spec_from_file_location(
    "xdet_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["xdet_b200"] = _mod
_spec.loader.exec_module(_mod)
