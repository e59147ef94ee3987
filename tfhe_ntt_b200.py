"""Import shim: the product package lives in `tfhe-rs-main_modified_b200/` (hyphens are not
importable), this module loads it under the name `tfhe_ntt_b200`."""
import importlib.util as _u
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "tfhe-rs-main_modified_b200")
_spec = _u.spec_from_file_location(__name__, _os.path.join(_pkg_dir, "__init__.py"),
                                   submodule_search_locations=[_pkg_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
