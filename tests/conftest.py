import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_count():
    try:
        import rimphony_b200 as R
        return R.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    # GPU tests are skipped only when there is no CUDA device at all.  On a GPU box a missing
    # or broken librimphony_b200.so is a FAILURE, not a skip: there is no fallback path.
    have_gpu = None
    for item in items:
        if "gpu" in item.keywords:
            if have_gpu is None:
                have_gpu = _cuda_device_count() > 0 or os.path.exists("/dev/nvidia0")
            if not have_gpu:
                item.add_marker(pytest.mark.skip(reason="no CUDA device in this container"))


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure), built on demand."""
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def golden():
    def load(name):
        fx = np.load(os.path.join(GOLDEN, name + ".npz"))
        return {"kind": int(fx["kind"]), "s": fx["s"], "theta": fx["theta"], "params": list(fx["params"]),
                "out": fx["out"], "lobes": fx["lobes"]}
    return load


@pytest.fixture(scope="session")
def symphony_rows():
    """The reference's golden vectors tests/symphony-powerlaw.txt (200 x [s, theta, p, J_I, A_I, J_Q, A_Q,
    J_V, A_V], computed with Symphony), stored as a binary table by tests/golden/make_golden.py."""
    return np.load(os.path.join(GOLDEN, "symphony_golden.npz"))["table"]
