import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def golden_cases():
    # brain cases only: the simulator recordings (sim_*.npz, tests/golden/make_env_golden.py) have their own tests
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith("sim_"))


@pytest.fixture(scope="session")
def v2v():
    import importlib
    return importlib.import_module("globecom2020-resourceallocationgnn_b200")
