"""Import shim: exposes the package in `bodyct-dram-emph-subtype_b200/` as `dram_b200`.

    import dram_b200                      # the package
    from dram_b200 import ops, med3d      # its submodules
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bodyct-dram-emph-subtype_b200")
_spec = importlib.util.spec_from_file_location(
    "dram_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dram_b200"] = _mod
_spec.loader.exec_module(_mod)
