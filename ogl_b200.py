"""Import shim: the product package lives in `online-gnn-learning_b200/` (a directory name that is not
a valid Python identifier), so it is loaded here under the importable name `ogl_b200`."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "online-gnn-learning_b200")
_NAME = "ogl_b200"

_spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[_NAME] = _mod          # replaces this shim module object under the same name
_spec.loader.exec_module(_mod)
