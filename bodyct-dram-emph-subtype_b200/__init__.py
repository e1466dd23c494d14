"""B200-native Med3D + dRAM inference hot path (drop-in for the reference's med3d/models/processor).

The directory name carries a hyphen, so it is imported through the root-level `dram_b200`
shim (or by adding this directory's parent to sys.path and using importlib); inside, modules
use relative imports.  The reference's flat module names (`med3d`, `models`, `utils`, ...)
are kept as submodule names so a `_target_: med3d.resnet34segreg` config resolves here.
"""
__all__ = ["_capi", "ops", "build"]
