"""Import alias for the product package, whose directory name (`dataframedbs.jl_b200/`, fixed by the
repo layout) is not a valid Python identifier.  `import dfdb_b200` loads that directory as a package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dataframedbs.jl_b200")
_spec = importlib.util.spec_from_file_location("dfdb_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dfdb_b200"] = _mod
_spec.loader.exec_module(_mod)
