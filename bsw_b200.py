"""Import shim: the package directory is named after the reference repo (`bwa-mem-sw_b200/`), which is not a valid
Python identifier, so it is loaded here under the module name `bsw_b200`."""
import importlib.util
import os
import sys

_pkg = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bwa-mem-sw_b200")
_spec = importlib.util.spec_from_file_location("bsw_b200", os.path.join(_pkg, "__init__.py"),
                                               submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["bsw_b200"] = _mod
_spec.loader.exec_module(_mod)
