"""Import alias: ``import flownet2_b200`` -> the package in ``flownet2-pytorch_b200/`` (hyphenated
directory names cannot be imported directly)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "flownet2-pytorch_b200")
_spec = importlib.util.spec_from_file_location(
    "flownet2_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["flownet2_b200"] = _mod
_spec.loader.exec_module(_mod)
