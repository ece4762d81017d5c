"""BASELINE config 2 (UR5e, 6 dof, CLIK + OSF, no QP) on the CPU: the synthesized model's anchors, the oracle against the
numpy restatement, and the product's kernel bodies (NV = 6 instantiation) against the oracle."""
import numpy as np
import pytest

from oracle import np_oracle
from tests.conftest import ROBOTS, workload

URDF, SRDF, LINK = str(ROBOTS / "ur5e" / "ur5e.urdf"), str(ROBOTS / "ur5e" / "ur5e.srdf"), "tool0"


@pytest.fixture(scope="module")
def rig():
    from oracle.c_oracle import Oracle
    from tests.emu import Emu
    return Oracle(URDF, SRDF, threads=4), Emu(URDF, SRDF)


def test_ur5e_anchors(rig):
    o, e = rig
    assert o.nv == 6 and e.nv == 6
    f = o.frame_id(LINK)
    T = o.update_state(np.zeros((1, 6)), np.zeros((1, 6)), f)["pose"][0].reshape(3, 4)
    # public UR5e DH table: x = a2 + a3, y = -(d4 + d6), z = d1 - d5 at q = 0
    assert np.allclose(T[:, 3], [-0.425 - 0.3922, -0.1333 - 0.0996, 0.1625 - 0.0997], atol=1e-12)
    assert abs(o.model.mass.sum() - (3.761 + 8.058 + 2.846 + 1.37 + 1.3 + 0.365)) < 1e-12   # base link is welded to the world
    q, qd, _, _ = workload(o.model, 3, 1)
    st = o.update_state(q, qd, f)
    for b in range(3):
        assert np.abs(st["J"][b] - np_oracle.frame_jacobian(o.model, q[b], f)).max() < 1e-12
        assert np.abs(st["M"][b] - np_oracle.mass_matrix(o.model, q[b])).max() < 1e-11
        assert np.abs(st["g"][b] - np_oracle.gravity(o.model, q[b])).max() < 1e-11


def test_ur5e_kernel_bodies(rig):
    o, e = rig
    B = 256
    q, qd, q_t, xdot_t = workload(o.model, B, 2)
    f, fe = o.frame_id(LINK), e.frame_id(LINK)
    ref = o.update_state(q, qd, f)
    r = e.update_and_get(q, qd, fe)
    rel = lambda a, b: np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
    assert rel(r["pose"], ref["pose"]) < 1e-12 and rel(r["J"], ref["J"]) < 1e-12 and rel(r["Jdot"], ref["Jdot"]) < 1e-11
    assert rel(r["M"], ref["M"]) < 1e-9 and rel(r["g"], ref["g"]) < 1e-9 and rel(r["nle"], ref["nle"]) < 1e-9
    x_t = o.update_state(q_t, qd, f)["pose"]
    null = np.random.default_rng(0).normal(size=(B, 6))
    for nv in (None, null):
        for mode in (0, 1):          # CLIKStep, OSFStep  (robot_controller.cpp:156-171, 232-240)
            a = e.taskspace(mode, q, qd, x_t, xdot_t, fe, aux=nv)
            b = o.taskspace(mode, q, qd, x_t, xdot_t, f, null_vec=nv)
            # a 6x6 Jacobian near a wrist singularity amplifies rounding: relative to the output scale
            assert np.abs(a - b).max() < 1e-6 * max(1.0, np.abs(b).max())
    # the QP controllers exist for the 6-dof shape too (QPIK 20 vars / 34 rows)
    ref_c = o.cycle(1, q, qd, x_t, xdot_t, f)
    c = e.cycle(1, q, qd, x_t, xdot_t, fe)
    same = (c["iters"] == ref_c["iters"]) & (c["status"] == ref_c["status"])
    assert same.mean() > 0.97
    # random UR5e postures over +-324 deg are often in (or near) self-collision: the active collision row carries the GJK
    # witness noise of its cylinder pairs into the command ("GJK witness precision")
    err = np.abs(c["out"] - ref_c["out"]).max(axis=1)[same]
    assert (err < 1e-5).mean() > 0.9 and (err < 1e-3).mean() > 0.99
