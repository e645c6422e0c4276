import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # built artefacts are git-ignored: compile libsos_b200.so (nvcc cross-compiles without a GPU) if it is missing
    lib = os.path.join(ROOT, "sos-radiative-transfer_b200", "libsos_b200.so")
    if not os.path.isfile(lib):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "sos-radiative-transfer_b200", "csrc")], check=True,
                       stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


def relmax(a, b):
    """max|a-b| / max|b| -- the parity metric of SURVEY.md 7 (hard part 4)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def relelem(a, b, floor=1e-6):
    """max elementwise relative error where |b| > floor * max|b|."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    m = np.abs(b) > floor * np.max(np.abs(b))
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m]) / np.abs(b[m])))
