"""Import shim: registers the package directory ``maniac-mc.github.io_b200/`` (not a valid
Python identifier) as the importable package ``maniac_b200``."""
import importlib.util
import sys
from pathlib import Path

_PKG_DIR = Path(__file__).resolve().parent / "maniac-mc.github.io_b200"
_spec = importlib.util.spec_from_file_location(
    "maniac_b200", _PKG_DIR / "__init__.py", submodule_search_locations=[str(_PKG_DIR)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["maniac_b200"] = _mod
_spec.loader.exec_module(_mod)
