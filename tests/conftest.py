import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

URDF = str(ROOT / "dyros_robot_controller_b200" / "robots" / "fr3" / "fr3.urdf")
SRDF = str(ROOT / "dyros_robot_controller_b200" / "robots" / "fr3" / "fr3.srdf")
LINK = "fr3_link8"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.c_oracle import Oracle
    return Oracle(URDF, SRDF, threads=8)


@pytest.fixture(scope="session")
def emu():
    from tests.emu import Emu
    return Emu(URDF, SRDF)


def workload(model_like, B, seed, stress=False):
    """BASELINE.md section 4 inputs.  model_like needs q_lo, q_hi, v_lim (arrays)."""
    rng = np.random.default_rng(seed)
    lo, hi, vl = model_like.q_lo, model_like.q_hi, model_like.v_lim
    n = len(lo)
    q = lo + (0.1 + 0.8 * rng.random((B, n))) * (hi - lo)
    if stress:
        k = B // 10  # 10 % within 0.02 rad of a joint limit
        j = rng.integers(0, n, k)
        side = rng.random(k) < 0.5
        q[np.arange(k), j] = np.where(side, lo[j] + 0.02 * rng.random(k), hi[j] - 0.02 * rng.random(k))
        k2 = B // 20  # 5 % near-singular (elbow stretched, wrist aligned)
        q[k:k + k2, 3] = -0.16 - 0.02 * rng.random(k2)
        q[k:k + k2, 5] = 0.55 + 0.02 * rng.random(k2)
    qd = rng.uniform(-0.5, 0.5, (B, n)) * vl
    q_t = q + 0.05 * rng.normal(size=(B, n))
    xdot_t = 0.05 * rng.normal(size=(B, 6))
    return q, qd, q_t, xdot_t


@pytest.fixture(scope="session")
def gpu_ctx():
    import dyros_robot_controller_b200 as drc
    if drc.device_count() < 1:
        pytest.fail("GPU tests need a CUDA device; the product path has no CPU fallback")
    model = drc.Model(URDF, SRDF)
    ctx = drc.Context(model, 65536, device=0)
    return model, ctx
