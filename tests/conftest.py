import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import maniac_b200  # noqa: E402,F401  (registers the package)

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def kat():
    return json.loads((GOLDEN / "kat.json").read_text())


@pytest.fixture(scope="session")
def load():
    from maniac_b200.snapshot import load_snapshot

    def _load(name):
        return load_snapshot(GOLDEN / f"{name}.npz")
    return _load


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (C) and make sure the CUDA library exists (cross-compiled by build())."""
    from oracle import oracle as orc
    orc.build()
    lib = ROOT / "maniac-mc.github.io_b200" / "libmaniac_gpu.so"
    if not lib.exists():
        import __graft_entry__ as ge
        ge.build()
    return True
