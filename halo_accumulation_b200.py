"""Import alias: the package directory is `halo-accumulation_b200/` (the name the project layout
prescribes); a hyphen cannot appear in a Python module name, so this module loads that directory
under the importable name `halo_accumulation_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "halo-accumulation_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
