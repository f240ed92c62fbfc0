import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def tables():
    data = ROOT / "vgpmp_b200" / "data"
    return json.loads((data / "robots.json").read_text()), json.loads((data / "problemsets.json").read_text())


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
