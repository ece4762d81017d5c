"""ORACLE (test infrastructure, never on the product path) -- URDF/SRDF -> flat rigid-body model.

PARITY UNPINNED: the reference delegates this step to Pinocchio
(`pinocchio::urdf::buildModel`, `buildGeom`, `addAllCollisionPairs`,
`srdf::removeCollisionPairs`; reference src/manipulator/robot_data.cpp:21-43), an
un-vendored, un-pinned dependency that is not installable here.  This file restates
the conventions of those calls:

  * the root link is welded to the universe; its inertia is dropped;
  * fixed joints are merged: the child link's inertia is lumped into the parent
    joint's body, its frames/geometries hang off the parent joint with the
    accumulated placement;
  * `rpy` -> R = Rz(yaw) Ry(pitch) Rx(roll);
  * revolute/continuous -> 1-dof revolute about `axis`; prismatic -> 1-dof prismatic;
  * joints are numbered depth-first in URDF child order (chain order for FR3);
  * dof = number of moving joints (reference robot_data.cpp:52);
  * limits: lower/upper position, +-velocity (robot_data.cpp:59-62);
  * every <collision> becomes a geometry object attached to the link's joint;
    all pairs (i<j) with different parent joints are enabled, then every pair whose
    links appear in an SRDF <disable_collisions> is removed, order preserved.

It is deliberately independent of the product's C++ model compiler
(dyros_robot_controller_b200/csrc/model.cpp); tests compare the two blobs.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional

from pathlib import Path

import numpy as np

GEOM_SPHERE, GEOM_CYLINDER, GEOM_BOX, GEOM_CAPSULE = 0, 1, 2, 3
JOINT_REVOLUTE, JOINT_PRISMATIC = 0, 1


def rpy_to_R(rpy) -> np.ndarray:
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _vec(s: Optional[str], n=3, default=0.0) -> np.ndarray:
    if s is None:
        return np.full(n, default)
    return np.array([float(t) for t in s.split()], dtype=np.float64)


def _origin(el) -> (np.ndarray, np.ndarray):
    if el is None:
        return np.eye(3), np.zeros(3)
    o = el.find("origin")
    if o is None:
        return np.eye(3), np.zeros(3)
    return rpy_to_R(_vec(o.get("rpy"))), _vec(o.get("xyz"))


@dataclass
class FlatModel:
    name: str = ""
    nv: int = 0
    joint_names: List[str] = field(default_factory=list)
    parent: np.ndarray = None      # (nv,) int32, -1 = universe
    jtype: np.ndarray = None       # (nv,) int32
    axis: np.ndarray = None        # (nv,3)
    jR: np.ndarray = None          # (nv,3,3) placement in parent joint frame
    jp: np.ndarray = None          # (nv,3)
    mass: np.ndarray = None        # (nv,)
    com: np.ndarray = None         # (nv,3) in joint frame
    inertia: np.ndarray = None     # (nv,3,3) about com, joint-frame axes
    q_lo: np.ndarray = None
    q_hi: np.ndarray = None
    v_lim: np.ndarray = None
    effort: np.ndarray = None
    frame_names: List[str] = field(default_factory=list)
    frame_parent: np.ndarray = None  # (nf,) int32 (-1 universe)
    frame_R: np.ndarray = None
    frame_p: np.ndarray = None
    geom_names: List[str] = field(default_factory=list)
    geom_link: List[str] = field(default_factory=list)
    geom_type: np.ndarray = None   # (ng,) int32
    geom_param: np.ndarray = None  # (ng,3): sphere r,-,- | cylinder r,halflen,- | box hx,hy,hz
    geom_parent: np.ndarray = None
    geom_R: np.ndarray = None
    geom_p: np.ndarray = None
    pairs: np.ndarray = None       # (np,2) int32
    gravity: np.ndarray = None     # (3,)

    def frame_id(self, name: str) -> int:
        return self.frame_names.index(name) if name in self.frame_names else -1



GEOM_CONVEX = 4


def resolve_mesh(filename: str, urdf_dir: str, packages_path: str) -> str:
    """package://pkg/rest -> packages_path/pkg/rest; file://abs; otherwise relative to the URDF (pinocchio's lookup)."""
    if filename.startswith("package://"):
        return str(Path(packages_path) / filename[len("package://"):])
    if filename.startswith("file://"):
        return filename[len("file://"):]
    return filename if filename.startswith("/") else str(Path(urdf_dir) / filename)


def read_mesh_vertices(path: str) -> np.ndarray:
    """(n,3) vertex coordinates of an STL (binary / ASCII), OBJ or COLLADA file."""
    raw = Path(path).read_bytes()
    ext = path.rsplit(".", 1)[-1].lower()
    if ext == "stl":
        if len(raw) >= 84:
            n = int(np.frombuffer(raw[80:84], "<u4")[0])
            if len(raw) == 84 + 50 * n:
                rec = np.frombuffer(raw[84:], np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
                return rec["v"].reshape(-1, 3).astype(np.float64)
        tok = raw.decode(errors="ignore").split()
        v = [tok[i + 1:i + 4] for i, t in enumerate(tok) if t == "vertex"]
        return np.array(v, np.float64)
    if ext == "obj":
        v = [ln.split()[1:4] for ln in raw.decode(errors="ignore").splitlines() if ln.startswith("v ")]
        return np.array(v, np.float64)
    if ext == "dae":
        root = ET.fromstring(raw)
        strip = lambda t: t.split("}")[-1]
        unit = 1.0
        out = []
        src = {}
        for el in root.iter():
            tag = strip(el.tag)
            if tag == "unit" and el.get("meter"):
                unit = float(el.get("meter"))
            if tag == "source":
                fa = [c for c in el if strip(c.tag) == "float_array"]
                if fa and fa[0].text:
                    src[el.get("id")] = np.array(fa[0].text.split(), np.float64)
        for el in root.iter():
            if strip(el.tag) == "vertices":
                for inp in el:
                    if strip(inp.tag) == "input" and inp.get("semantic") == "POSITION":
                        out.append(src[inp.get("source").lstrip("#")].reshape(-1, 3))
        return unit * np.concatenate(out)
    raise ValueError(f"unsupported mesh format: {path}")


def load(urdf_path: str, srdf_path: str = "", packages_path: str = "") -> FlatModel:
    root = ET.parse(urdf_path).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    link_order = [l.get("name") for l in root.findall("link")]
    joints = root.findall("joint")
    children: Dict[str, list] = {n: [] for n in links}
    has_parent = set()
    for j in joints:
        children[j.find("parent").get("link")].append(j)
        has_parent.add(j.find("child").get("link"))
    roots = [n for n in link_order if n not in has_parent]
    assert len(roots) == 1, "URDF must have exactly one root link"

    m = FlatModel(name=root.get("name", ""))
    jn, par, jt, ax, jR, jp = [], [], [], [], [], []
    mass, mc, Io = [], [], []   # accumulate mass, first moment and inertia about the joint origin
    qlo, qhi, vl, ef = [], [], [], []
    fn, fpar, fR, fp = [], [], [], []
    gn, gl, gt, gpar_, gparam, gR, gp = [], [], [], [], [], [], []
    hull_pts, hull_off = [], {}
    urdf_dir = str(Path(urdf_path).resolve().parent)

    def add_body(jidx, R, p, link):
        """Lump `link`'s inertia (placed at (R,p) in joint jidx's frame) into that joint's body."""
        ine = link.find("inertial")
        if ine is None or jidx < 0:
            return
        Ri, pi = _origin(ine)
        mval = float(ine.find("mass").get("value"))
        it = ine.find("inertia")
        I = np.array([[float(it.get("ixx")), float(it.get("ixy")), float(it.get("ixz"))],
                      [float(it.get("ixy")), float(it.get("iyy")), float(it.get("iyz"))],
                      [float(it.get("ixz")), float(it.get("iyz")), float(it.get("izz"))]])
        Rw = R @ Ri
        c = R @ pi + p
        Ic = Rw @ I @ Rw.T
        mass[jidx] += mval
        mc[jidx] += mval * c
        Io[jidx] += Ic + mval * (np.dot(c, c) * np.eye(3) - np.outer(c, c))

    def add_link_items(jidx, R, p, lname):
        link = links[lname]
        fn.append(lname); fpar.append(jidx); fR.append(R.copy()); fp.append(p.copy())
        add_body(jidx, R, p, link)
        for k, c in enumerate(link.findall("collision")):
            Rc, pc = _origin(c)
            g = list(c.find("geometry"))[0]
            if g.tag == "sphere":
                t, prm = GEOM_SPHERE, [float(g.get("radius")), 0, 0]
            elif g.tag == "cylinder":
                t, prm = GEOM_CYLINDER, [float(g.get("radius")), 0.5 * float(g.get("length")), 0]
            elif g.tag == "box":
                s = _vec(g.get("size"))
                t, prm = GEOM_BOX, list(0.5 * s)
            elif g.tag == "capsule":
                t, prm = GEOM_CAPSULE, [float(g.get("radius")), 0.5 * float(g.get("length")), 0]
            elif g.tag == "mesh":
                # robot_data.cpp:24-34 (buildGeom with packages_path): the mesh's vertices -> convex hull (scipy / Qhull here,
                # independent of the product's GJK-based hull selection); vertices stay in the mesh's own frame
                pts = read_mesh_vertices(resolve_mesh(g.get("filename"), urdf_dir, packages_path)) * _vec(g.get("scale"), 3, 1.0)
                from scipy.spatial import ConvexHull
                hv = pts[ConvexHull(pts).vertices]
                t, prm = GEOM_CONVEX, [float(np.linalg.norm(hv, axis=1).max()), 0, 0]
                hull_off[len(gn)] = (len(hull_pts), len(hv))
                hull_pts.extend(hv.tolist())
            else:
                continue
            gn.append(f"{lname}_{k}"); gl.append(lname); gt.append(t); gparam.append(prm)
            gpar_.append(jidx); gR.append(R @ Rc); gp.append(R @ pc + p)

    def visit(lname, jidx, R, p):
        add_link_items(jidx, R, p, lname)
        for j in children[lname]:
            Rj, pj = _origin(j)
            Rn, pn = R @ Rj, R @ pj + p
            typ = j.get("type")
            child = j.find("child").get("link")
            if typ == "fixed":
                visit(child, jidx, Rn, pn)
            elif typ in ("revolute", "continuous", "prismatic"):
                new = len(jn)
                jn.append(j.get("name")); par.append(jidx)
                jt.append(JOINT_PRISMATIC if typ == "prismatic" else JOINT_REVOLUTE)
                a = _vec(j.find("axis").get("xyz")) if j.find("axis") is not None else np.array([1.0, 0, 0])
                ax.append(a / np.linalg.norm(a))
                jR.append(Rn); jp.append(pn)
                mass.append(0.0); mc.append(np.zeros(3)); Io.append(np.zeros((3, 3)))
                lim = j.find("limit")
                if lim is not None:
                    qlo.append(float(lim.get("lower", "-inf")) if typ != "continuous" else -np.inf)
                    qhi.append(float(lim.get("upper", "inf")) if typ != "continuous" else np.inf)
                    vl.append(float(lim.get("velocity", "inf"))); ef.append(float(lim.get("effort", "inf")))
                else:
                    qlo.append(-np.inf); qhi.append(np.inf); vl.append(np.inf); ef.append(np.inf)
                visit(child, new, np.eye(3), np.zeros(3))
            else:
                raise ValueError(f"unsupported joint type {typ}")

    visit(roots[0], -1, np.eye(3), np.zeros(3))

    nv = len(jn)
    m.nv = nv
    m.joint_names = jn
    m.parent = np.array(par, np.int32)
    m.jtype = np.array(jt, np.int32)
    m.axis = np.array(ax).reshape(nv, 3)
    m.jR = np.array(jR).reshape(nv, 3, 3)
    m.jp = np.array(jp).reshape(nv, 3)
    m.mass = np.array(mass)
    m.com = np.zeros((nv, 3))
    m.inertia = np.zeros((nv, 3, 3))
    for i in range(nv):
        if mass[i] > 0:
            c = mc[i] / mass[i]
            m.com[i] = c
            m.inertia[i] = Io[i] - mass[i] * (np.dot(c, c) * np.eye(3) - np.outer(c, c))
    m.q_lo, m.q_hi, m.v_lim, m.effort = (np.array(a, np.float64) for a in (qlo, qhi, vl, ef))
    m.frame_names = fn
    m.frame_parent = np.array(fpar, np.int32)
    m.frame_R = np.array(fR).reshape(len(fn), 3, 3)
    m.frame_p = np.array(fp).reshape(len(fn), 3)
    ng = len(gn)
    m.geom_names, m.geom_link = gn, gl
    m.geom_type = np.array(gt, np.int32)
    m.geom_param = np.array(gparam, np.float64).reshape(ng, 3)
    m.geom_parent = np.array(gpar_, np.int32)
    m.geom_R = np.array(gR).reshape(ng, 3, 3)
    m.geom_p = np.array(gp).reshape(ng, 3)
    m.gravity = np.array([0.0, 0.0, -9.81])
    m.hull = np.array(hull_pts, np.float64).reshape(-1, 3)
    m.hull_off = np.array([hull_off.get(g, (0, 0))[0] for g in range(ng)], np.int32)
    m.hull_n = np.array([hull_off.get(g, (0, 0))[1] for g in range(ng)], np.int32)

    disabled = set()
    if srdf_path:
        try:
            sroot = ET.parse(srdf_path).getroot()
            for d in sroot.findall("disable_collisions"):
                disabled.add(frozenset((d.get("link1"), d.get("link2"))))
        except (FileNotFoundError, ET.ParseError):
            pass
    pairs = []
    for i in range(ng):
        for j in range(i + 1, ng):
            if m.geom_parent[i] == m.geom_parent[j]:
                continue
            if frozenset((gl[i], gl[j])) in disabled:
                continue
            pairs.append((i, j))
    m.pairs = np.array(pairs, np.int32).reshape(-1, 2)
    return m
