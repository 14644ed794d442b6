"""Import alias: the package directory is named `holistic-robot-pose-estimation_b200/` (not a valid Python
identifier), so this module loads it and registers it as `horopose_b200`."""
import importlib.util
import sys
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent / "holistic-robot-pose-estimation_b200"
_spec = importlib.util.spec_from_file_location(
    "horopose_b200", _pkg_dir / "__init__.py", submodule_search_locations=[str(_pkg_dir)]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["horopose_b200"] = _mod
_spec.loader.exec_module(_mod)
