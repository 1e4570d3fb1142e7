"""Import shim: the package directory is `noaa-apt_b200/` (the name the build
contract asks for), which is not a valid Python identifier.  `import
noaa_apt_b200` lands here and swaps itself for the real package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "noaa-apt_b200")
_spec = importlib.util.spec_from_file_location(
    "noaa_apt_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["noaa_apt_b200"] = _mod
_spec.loader.exec_module(_mod)
