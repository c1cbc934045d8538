"""Import shim: the package directory is `avbd-demo3d_b200/` (the name the build contract fixes), which is not a
valid Python identifier, so `import avbd_demo3d_b200` lands here and loads that directory as the package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "avbd-demo3d_b200")
_spec = importlib.util.spec_from_file_location("avbd_demo3d_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["avbd_demo3d_b200"] = _mod
_spec.loader.exec_module(_mod)
