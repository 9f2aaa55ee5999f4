import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def check_digest(grads: dict, gold: dict, prefix: str = "", rtol=2e-4, atol=2e-6):
    """Compares parameter grads against a make_golden.grad_digest (sum, abs-sum, 64 samples)."""
    names = sorted(k[len(prefix):-4] for k in gold if k.startswith(prefix) and k.endswith("/sum"))
    assert names, "empty digest"
    for name in names:
        g = np.asarray(grads[name], dtype=np.float32).reshape(-1)
        pos = gold[f"{prefix}{name}/pos"]
        val = gold[f"{prefix}{name}/val"]
        scale = float(gold[f"{prefix}{name}/abssum"]) / g.size + 1e-12
        np.testing.assert_allclose(g[pos], val, rtol=rtol, atol=atol + rtol * scale, err_msg=name)
        np.testing.assert_allclose(np.abs(g.astype(np.float64)).sum(), gold[f"{prefix}{name}/abssum"], rtol=rtol,
                                   err_msg=name)
