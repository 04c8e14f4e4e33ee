import gzip
import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def load_golden(name):
    with gzip.open(ROOT / "tests" / "golden" / name, "rb") as f:
        return json.loads(f.read())


@pytest.fixture(scope="session")
def acvp_keygen():
    return load_golden("acvp_keygen.json.gz")


@pytest.fixture(scope="session")
def acvp_siggen():
    return load_golden("acvp_siggen.json.gz")


@pytest.fixture(scope="session")
def acvp_sigver():
    return load_golden("acvp_sigver.json.gz")


@pytest.fixture(scope="session")
def reference_kats():
    return load_golden("reference_kats.json.gz")


PS_OF = {"ML-DSA-44": 44, "ML-DSA-65": 65, "ML-DSA-87": 87}
