import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def hostcheck_lib():
    """Test-only host build of the plan compiler (tests/hostcheck)."""
    import subprocess
    import iexa_b200 as ex
    d = os.path.join(ROOT, "tests", "hostcheck")
    subprocess.check_call(["make", "-C", d, "libiexa_hostcheck.so"], stdout=subprocess.DEVNULL)
    return ex.lib.load(os.path.join(d, "libiexa_hostcheck.so"))


def eval_point(core, seed=0, scale=0.1):
    """x = x0 + scale·U(−1,1) clipped to the bounds, y ~ U(−1,1)  (SURVEY.md §8(d))."""
    rng = np.random.default_rng(seed)
    x = core.x0_vec + scale * rng.uniform(-1, 1, core.nvar)
    x = np.minimum(np.maximum(x, core.lvar_vec), core.uvar_vec)
    y = rng.uniform(-1, 1, core.ncon)
    return x, y


# tolerance stated by BASELINE.json north_star: 1e-12 relative, 1e-14 absolute
RTOL, ATOL = 1e-12, 1e-14


def assert_close(a, b, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    err = np.abs(a - b)
    tol = ATOL + RTOL * np.maximum(np.abs(a), np.abs(b))
    bad = ~(err <= tol) & ~(np.isnan(a) & np.isnan(b))
    assert not bad.any(), f"{what}: {bad.sum()} entries differ; worst abs {err[bad].max():.3e} at {np.argmax(bad)}"
