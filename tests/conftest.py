import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import kp_oracle
    kp_oracle.build()
    return kp_oracle


@pytest.fixture(scope="session")
def golden_modelnet():
    return dict(np.load(os.path.join(GOLDEN, "modelnet_pair.npz")))


@pytest.fixture(scope="session")
def golden_3dmatch():
    return dict(np.load(os.path.join(GOLDEN, "threedmatch_small_pyramid.npz")))


# measured parity errors of the run, written to gpurun_out/parity_errors.json at session end (committed under profiles/)
PARITY = {}


def record_parity(name, value, bound):
    PARITY[name] = {"measured": float(value), "bound": float(bound)}


def pytest_sessionfinish(session, exitstatus):
    if PARITY:
        import json
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_errors.json"), "w") as fh:
            json.dump(PARITY, fh, indent=1, sort_keys=True)


def rel_err(a, b):
    """max |a-b| / max |b| — the norm-wise relative error the 1e-4 feature tolerance is stated in."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="session")
def golden_encoder_r2():
    return dict(np.load(os.path.join(GOLDEN, "encoder_r2.npz")))


@pytest.fixture(scope="session")
def golden_overlaps_r2():
    return dict(np.load(os.path.join(GOLDEN, "overlaps_r2.npz")))
