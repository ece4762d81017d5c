"""Mesh collision geometry (reference src/manipulator/robot_data.cpp:24-34: buildGeom with packages_path).

CPU suite: the model compiler's mesh readers / hull selection (through the host emulation of the kernel bodies) against the
oracle's independent loader (numpy readers + scipy / Qhull hull), and the narrow phase on hull geometry against the oracle and
against the primitive model the box mesh was generated from.  Robot: robots/fr3_mesh (tools/make_mesh_robot.py)."""
from pathlib import Path

import numpy as np
import pytest

from tests.conftest import ROOT, SRDF, URDF, workload

MESH_DIR = ROOT / "dyros_robot_controller_b200" / "robots" / "fr3_mesh"
PKGS = str(MESH_DIR / "packages")
BOXMESH, MESH, MESH_SRDF = str(MESH_DIR / "fr3_boxmesh.urdf"), str(MESH_DIR / "fr3_mesh.urdf"), str(MESH_DIR / "fr3_mesh.srdf")


@pytest.fixture(scope="module")
def emus():
    from tests.emu import Emu
    return Emu(URDF, SRDF), Emu(BOXMESH, SRDF, PKGS), Emu(MESH, MESH_SRDF, PKGS)


@pytest.fixture(scope="module")
def mesh_oracle():
    from oracle.c_oracle import Oracle
    return Oracle(MESH, MESH_SRDF, threads=8, packages_path=PKGS)


def test_hull_selection_matches_qhull(emus, mesh_oracle):
    _, ebox, emesh = emus
    bi, mi = ebox.mesh_info(), emesh.mesh_info()
    assert bi["mesh_geoms"] == 1 and bi["hull_vertices"] == 8
    assert mi["mesh_geoms"] == 3 and sorted(mi["vert_n"][mi["vert_n"] > 0]) == [8, 64, 162]   # box, 32-gon prism, icosphere (2 subdivisions)
    # same vertex SETS as the oracle's loader (Qhull), up to the product's shift of each hull to its box centre
    om = mesh_oracle.model
    assert om.hull_n.sum() == mi["hull_vertices"]
    off = 0
    for g in range(len(om.hull_n)):
        n = int(om.hull_n[g])
        if not n:
            continue
        ref = om.hull[om.hull_off[g]:om.hull_off[g] + n]
        mine = mi["hull"][off:off + n]
        off += n
        ref_c = ref - 0.5 * (ref.min(0) + ref.max(0))
        key = lambda a: a[np.lexsort(np.round(a, 9).T[::-1])]
        np.testing.assert_allclose(key(mine), key(ref_c), atol=1e-12)


def test_box_mesh_equals_box_primitive(emus):
    eprim, ebox, _ = emus
    q, qd, _, _ = workload(type("M", (), dict(q_lo=eprim.model_arrays()["q_lo"], q_hi=eprim.model_arrays()["q_hi"], v_lim=eprim.model_arrays()["v_lim"])), 512, 3, stress=True)
    a, b = eprim.min_distance(q, qd), ebox.min_distance(q, qd)
    # the hull of the 8-vertex mesh IS the box: same minimum, same pair, same gradient (support points coincide; the iterates differ
    # only through the centre-based bound and the GJK start)
    np.testing.assert_allclose(b["d"], a["d"], atol=1e-9)
    assert (a["pair"] == b["pair"]).mean() > 0.995
    same = a["pair"] == b["pair"]
    np.testing.assert_allclose(b["grad"][same], a["grad"][same], atol=2e-4)
    # some states do have the hand box as the closest geometry, otherwise the test says nothing
    names = eprim.model_arrays()
    assert same.sum() > 400


def test_mesh_robot_matches_oracle(emus, mesh_oracle):
    _, _, emesh = emus
    o = mesh_oracle
    q, qd, _, _ = workload(o.model, 768, 5, stress=True)
    mine, ref = emesh.min_distance(q, qd), o.min_distance(q, qd)
    same = mine["pair"] == ref["pair"]
    assert same.mean() > 0.995
    np.testing.assert_allclose(mine["d"], ref["d"], atol=1e-7)
    # faceted hulls: the witness point of a (nearly) parallel face pair is not unique, so gradients are compared where the
    # witnesses agree, and by finite differences of d for all
    close = same & (np.abs(mine["witness"][:, :3] - ref["pa"]).max(1) < 1e-5) & (np.abs(mine["witness"][:, 3:] - ref["pb"]).max(1) < 1e-5)
    assert close.mean() > 0.9
    np.testing.assert_allclose(mine["grad"][close], ref["grad"][close], atol=1e-4)
    # hull geometry takes part: some minima come from pairs with a mesh geometry
    gt = o.model.geom_type
    pa, pb = o.model.pairs[ref["pair"], 0], o.model.pairs[ref["pair"], 1]
    assert ((gt[pa] == 4) | (gt[pb] == 4)).sum() > 20
    # overlapping hull pairs went through EPA
    assert (mine["d"] < 0).sum() == (ref["d"] < 0).sum()


def test_mesh_distance_against_constrained_minimisation(emus, mesh_oracle):
    """Independent truth for hull pairs: min |x_a - x_b| over convex combinations of the two vertex sets (scipy SLSQP)."""
    from scipy.optimize import minimize
    o = mesh_oracle
    om = o.model
    rng = np.random.default_rng(11)
    q = om.q_lo + (0.2 + 0.6 * rng.random(om.nv)) * (om.q_hi - om.q_lo)
    d_all, pa, pb, _ = o.pair_distances(q)
    # joint placements by plain numpy forward kinematics on the oracle's flat model (revolute chain)
    oR, op = [], []
    for i in range(om.nv):
        a = om.axis[i]
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        Rq = np.eye(3) + np.sin(q[i]) * K + (1 - np.cos(q[i])) * K @ K
        Rp, pp = (np.eye(3), np.zeros(3)) if om.parent[i] < 0 else (oR[om.parent[i]], op[om.parent[i]])
        oR.append(Rp @ om.jR[i] @ Rq); op.append(Rp @ om.jp[i] + pp)
    checked = 0
    for k, (ga, gb) in enumerate(om.pairs):
        if om.geom_type[ga] != 4 or om.geom_type[gb] != 4 or d_all[k] <= 0:
            continue
        def world(g):
            j = om.geom_parent[g]
            R, p = (np.eye(3), np.zeros(3)) if j < 0 else (oR[j], op[j])
            v = om.hull[om.hull_off[g]:om.hull_off[g] + om.hull_n[g]]
            return (v @ om.geom_R[g].T + om.geom_p[g]) @ R.T + p
        A, B = world(ga), world(gb)
        na, nb = len(A), len(B)
        f = lambda w: np.sum((w[:na] @ A - w[na:] @ B) ** 2)
        cons = [dict(type="eq", fun=lambda w: w[:na].sum() - 1), dict(type="eq", fun=lambda w: w[na:].sum() - 1)]
        w0 = np.concatenate([np.full(na, 1 / na), np.full(nb, 1 / nb)])
        r = minimize(f, w0, bounds=[(0, 1)] * (na + nb), constraints=cons, method="SLSQP", options=dict(maxiter=500, ftol=1e-16))
        assert abs(np.sqrt(r.fun) - d_all[k]) < 1e-5, (k, np.sqrt(r.fun), d_all[k])
        checked += 1
    assert checked >= 1


def test_mesh_errors_and_lookup(tmp_path):
    from tests.emu import Emu
    with pytest.raises(RuntimeError, match="packages_path"):
        Emu(MESH, MESH_SRDF)                       # package:// without a package directory
    with pytest.raises(RuntimeError, match="cannot open"):
        Emu(MESH, MESH_SRDF, str(tmp_path))        # package directory without the file
    # ASCII STL and file:// names
    stl = tmp_path / "tet.stl"
    stl.write_text("solid t\nfacet normal 0 0 1\nouter loop\nvertex 0 0 0\nvertex 0.1 0 0\nvertex 0 0.1 0\nendloop\nendfacet\n"
                   "facet normal 0 0 1\nouter loop\nvertex 0 0 0\nvertex 0 0 0.1\nvertex 0.1 0 0\nendloop\nendfacet\nendsolid t\n")
    urdf = Path(BOXMESH).read_text().replace("package://fr3_mesh_description/meshes/hand_box.stl", f"file://{stl}")
    u = tmp_path / "tet.urdf"
    u.write_text(urdf)
    e = Emu(str(u), MESH_SRDF)
    assert e.mesh_info()["hull_vertices"] == 4
