"""Mesh collision geometry on the GPU (reference src/manipulator/robot_data.cpp:24-34): the CUDA path through the
C ABI on robots/fr3_mesh (binary STL box, OBJ prism, COLLADA icosphere -> convex hulls) against the oracle (independent loader,
scipy / Qhull hulls), and the box mesh against the primitive it was generated from."""
import numpy as np
import pytest

from tests.conftest import LINK, SRDF, URDF, workload
from tests.test_mesh_cpu import BOXMESH, MESH, MESH_SRDF, PKGS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mesh_ctx():
    import dyros_robot_controller_b200 as drc
    model = drc.Model(MESH, MESH_SRDF, PKGS)
    assert model.info["mesh_geoms"] == 3 and model.info["hull_vertices"] == 8 + 64 + 162
    return model, drc.Context(model, 8192, device=0)


@pytest.fixture(scope="module")
def mesh_oracle():
    from oracle.c_oracle import Oracle
    return Oracle(MESH, MESH_SRDF, threads=8, packages_path=PKGS)


def test_box_mesh_equals_box_primitive_on_gpu(gpu_ctx):
    import dyros_robot_controller_b200 as drc
    model, ctx = gpu_ctx
    mbox = drc.Model(BOXMESH, SRDF, PKGS)
    cbox = drc.Context(mbox, 4096, device=0)
    q, qd, _, _ = workload(type("M", (), dict(q_lo=model.q_lower, q_hi=model.q_upper, v_lim=model.v_limit)), 4096, 21, stress=True)
    ctx.update_state(q, qd); cbox.update_state(q, qd)
    (da, ga, _, pa), (db, gb, _, pb) = ctx.get_min_distance(with_graddot=True), cbox.get_min_distance(with_graddot=True)
    np.testing.assert_allclose(db, da, atol=1e-8)   # GJK stops at a 1e-10 duality gap
    same = pa == pb
    assert same.mean() > 0.995
    np.testing.assert_allclose(gb[same], ga[same], atol=2e-4)


def test_mesh_min_distance_matches_oracle(mesh_ctx, mesh_oracle):
    model, ctx = mesh_ctx
    o = mesh_oracle
    B = 4096
    q, qd, _, _ = workload(o.model, B, 8, stress=True)
    ctx.update_state(q, qd)
    (d, g, gd, pair), ref = ctx.get_min_distance(with_graddot=True), o.min_distance(q, qd)
    same = pair == ref["pair"]
    assert same.mean() > 0.995
    np.testing.assert_allclose(d, ref["d"], atol=1e-7)
    gt = o.model.geom_type
    pa, pb = o.model.pairs[ref["pair"], 0], o.model.pairs[ref["pair"], 1]
    hull = (gt[pa] == 4) | (gt[pb] == 4)
    assert hull.sum() > 100, "the batch never has a mesh hull as the closest geometry"
    # primitive pairs: as in test_gpu_parity; hull pairs: the witness points of (nearly) parallel facets are not unique, so a few per
    # cent of them may pick another point of the same closest feature
    prim = same & ~hull
    assert np.abs(g - ref["grad"])[prim].max() < 1e-4 and np.abs(gd - ref["grad_dot"])[prim].max() < 1e-4
    hs = same & hull
    assert (np.abs(g - ref["grad"])[hs].max(1) < 1e-4).mean() > 0.9
    assert ((d < 0) == (ref["d"] < 0)).all()     # overlapping hull pairs: EPA on both sides


def test_mesh_control_cycle_matches_oracle(mesh_ctx, mesh_oracle):
    model, ctx = mesh_ctx
    o = mesh_oracle
    B = 2048
    q, qd, q_t, xd = workload(o.model, B, 9, stress=True)
    fid = o.frame_id(LINK)
    x_t = o.update_state(q_t, qd, fid)["pose"]
    r = ctx.cycle_qpik_step(q, qd, x_t, xd, LINK)
    ref = o.cycle(1, q, qd, x_t, xd, fid)
    same = (r["status"] == ref["status"]) & (r["iters"] == ref["iters"])
    assert same.mean() > 0.98
    err = np.abs(r["out"] - ref["out"]).max(1)
    assert (err[same] < 1e-4).mean() > 0.99
