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


# ---- mobile manipulators (synthesized URDFs, tools/make_moma_urdf.py): joints [virtual 0..2, wheels 3..3+w-1, arm]
ROBOTS = ROOT / "dyros_robot_controller_b200" / "robots"
MOMA = {
    "husky_fr3": dict(urdf=str(ROBOTS / "husky_fr3" / "husky_fr3.urdf"), srdf=str(ROBOTS / "husky_fr3" / "husky_fr3.srdf"),
                      kin=dict(type="Differential", wheel_radius=0.1651, base_width=0.555),
                      joint_idx=dict(virtual_start=0, mobi_start=3, mani_start=5), actuator_idx=dict(mobi_start=0, mani_start=2)),
    # mecanum parameters of the reference's XLS example (examples/C++/src/xls_controller.cpp:18-27)
    "xls_fr3": dict(urdf=str(ROBOTS / "xls_fr3" / "xls_fr3.urdf"), srdf=str(ROBOTS / "xls_fr3" / "xls_fr3.srdf"),
                    kin=dict(type="Mecanum", wheel_radius=0.120, roller_angles=[-np.pi / 4, np.pi / 4, np.pi / 4, -np.pi / 4],
                             base2wheel_positions=[(0.2225, 0.2045), (0.2225, -0.2045), (-0.2225, 0.2045), (-0.2225, -0.2045)],
                             base2wheel_angles=[0.0, 0.0, 0.0, 0.0]),
                    joint_idx=dict(virtual_start=0, mobi_start=3, mani_start=7), actuator_idx=dict(mobi_start=0, mani_start=4)),
    # powered casters (mobile/robot_data.cpp:179-203): two casters, wheel joints ordered (steer, roll) per caster
    "pcv_fr3": dict(urdf=str(ROBOTS / "pcv_fr3" / "pcv_fr3.urdf"), srdf=str(ROBOTS / "pcv_fr3" / "pcv_fr3.srdf"),
                    kin=dict(type="Caster", wheel_radius=0.055, wheel_offset=0.020,
                             base2wheel_positions=[(0.215, 0.125), (-0.215, -0.125)]),
                    joint_idx=dict(virtual_start=0, mobi_start=3, mani_start=7), actuator_idx=dict(mobi_start=0, mani_start=4)),
}


def moma_workload(model_like, w, B, seed):
    """base pose U([-2,2]^2 x [-pi,pi]), wheel angles U(-pi,pi), wheel speeds U(-2,2); arm as config 1."""
    rng = np.random.default_rng(seed)
    n = len(model_like.q_lo)
    lo, hi, vl = model_like.q_lo[3 + w:], model_like.q_hi[3 + w:], model_like.v_lim[3 + w:]
    q, qd = np.zeros((B, n)), np.zeros((B, n))
    q[:, 0:2] = rng.uniform(-2, 2, (B, 2)); q[:, 2] = rng.uniform(-np.pi, np.pi, B)
    q[:, 3:3 + w] = rng.uniform(-np.pi, np.pi, (B, w))
    q[:, 3 + w:] = lo + (0.1 + 0.8 * rng.random((B, n - 3 - w))) * (hi - lo)
    qd[:, 3:3 + w] = rng.uniform(-2, 2, (B, w))
    qd[:, 3 + w:] = rng.uniform(-0.5, 0.5, (B, n - 3 - w)) * vl
    q_t = q.copy()
    q_t[:, 3 + w:] += 0.05 * rng.normal(size=(B, n - 3 - w))
    q_t[:, 0:2] += 0.05 * rng.normal(size=(B, 2))
    xdot_t = 0.05 * rng.normal(size=(B, 6))
    return q, qd, q_t, xdot_t


def struct_to_ref(kind, dbg, n_mani=None, am=0):
    """Map the product's structured primal / dual vectors (Context.qp_debug) to the reference's QP order
    (x as laid out by the QP classes, rows = [bounds; inequalities; equalities], QP_base.h:204-226).

    kind "ik":      x = [qdot s_qmin s_qmax s_sing s_col]                          (QP_IK.cpp:12-34)
    kind "id":      x = [qddot tau s_q- s_q+ s_v- s_v+ s_sing s_col]               (QP_ID.cpp:12-65)
    kind "moma_ik": x = eta; rows = [free bounds (act); qmin, qmax (mani); sing; col]   (mobile_manipulator/QP_IK.cpp:14-28)
    kind "moma_id": x = [etadot tau]; rows = [q-, q+, v-, v+ (mani); sing; col; dynamics (act)]   (QP_ID.cpp:14-36)
    n_mani / am: manipulator dof and its first actuated index (whole-body QPs)."""
    x, y, nc, ku, nr = dbg["x"], dbg["y"], dbg["nc"], dbg["ku"], dbg["nr"]
    xc, xs, xr = x[:, :nc], x[:, nc:nc * (1 + ku)].reshape(-1, ku, nc), x[:, nc * (1 + ku):]
    yb = y[:, :nc]
    yu = y[:, nc:nc * (1 + ku)].reshape(-1, ku, nc)
    ysb = y[:, nc * (1 + ku):nc * (1 + 2 * ku)].reshape(-1, ku, nc)
    yr = y[:, nc * (1 + 2 * ku):nc * (1 + 2 * ku) + nr]
    yrb = y[:, nc * (1 + 2 * ku) + nr:]
    B = x.shape[0]
    if kind == "ik":
        xr_ = np.concatenate([xc, xs.reshape(B, -1), xr], axis=1)
        yr_ = np.concatenate([yb, ysb.reshape(B, -1), yrb, yu.reshape(B, -1), yr], axis=1)
    elif kind == "id":
        tau, s2 = xr[:, 2:], xr[:, :2]
        xr_ = np.concatenate([xc, tau, xs.reshape(B, -1), s2], axis=1)
        yr_ = np.concatenate([yb, yrb[:, 2:], ysb.reshape(B, -1), yrb[:, :2], yu.reshape(B, -1), yr[:, :2], yr[:, 2:]], axis=1)
    elif kind == "moma_ik":
        m = n_mani
        xr_ = xc
        yr_ = np.concatenate([yb, yu[:, :, am:am + m].reshape(B, -1), yr], axis=1)
    else:
        m = n_mani
        xr_ = np.concatenate([xc, xr[:, 2:]], axis=1)
        yr_ = np.concatenate([yu[:, :, am:am + m].reshape(B, -1), yr[:, :2], yr[:, 2:]], axis=1)
    return xr_, yr_
