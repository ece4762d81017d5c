#!/usr/bin/env python3
"""Synthesises robots/fr3_mesh: the FR3 with some collision primitives replaced by MESH files, to exercise the mesh branch of the
model compiler (reference src/manipulator/robot_data.cpp:24-34: buildGeom with a packages_path).

  fr3_boxmesh.urdf  the hand's box primitive -> an 8-vertex binary STL of the same box (package:// name).  Its convex hull IS the
                    box, so every distance must equal the primitive model's.
  fr3_mesh.urdf     additionally: the first cylinder of link 5 -> a 32-sided prism (OBJ, name relative to the URDF, with a scale),
                    the first sphere of link 0 -> an icosphere (COLLADA in millimetres, package:// name).

Meshes are generated here (nothing is taken from the reference's assets); SRDF = the FR3's plus one disabled link pair.

    python tools/make_mesh_robot.py
"""
import re
import struct
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "dyros_robot_controller_b200" / "robots" / "fr3"
OUT = ROOT / "dyros_robot_controller_b200" / "robots" / "fr3_mesh"
PKG = OUT / "packages" / "fr3_mesh_description" / "meshes"


def write_stl(path, tris):
    with open(path, "wb") as f:
        f.write(b"drc_b200 synthetic mesh".ljust(80, b" "))
        f.write(struct.pack("<I", len(tris)))
        for t in tris:
            n = np.cross(t[1] - t[0], t[2] - t[0])
            n = n / (np.linalg.norm(n) or 1.0)
            f.write(struct.pack("<12fH", *n, *t[0], *t[1], *t[2], 0))


def box_tris(h):
    c = np.array([[sx * h[0], sy * h[1], sz * h[2]] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)])
    faces = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    return [np.array([c[a], c[b], c[d]]) for a, b, d, e in faces] + [np.array([c[a], c[d], c[e]]) for a, b, d, e in faces]


def icosphere(r, sub=2):
    t = (1 + 5 ** 0.5) / 2
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    v = [np.array(p, float) / np.linalg.norm(p) for p in v]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(sub):
        cache, nf = {}, []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[k] = len(v) - 1
            return cache[k]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return r * np.array(v), f


def main():
    PKG.mkdir(parents=True, exist_ok=True)
    (OUT / "meshes").mkdir(parents=True, exist_ok=True)
    urdf = (SRC / "fr3.urdf").read_text()
    # 1. the hand box (size 0.071 0.071 0.041)
    m = re.search(r'<geometry><box size="([^"]+)"/></geometry>', urdf)
    size = np.array(m.group(1).split(), float)
    write_stl(PKG / "hand_box.stl", box_tris(0.5 * size))
    box_urdf = urdf.replace(m.group(0), '<geometry><mesh filename="package://fr3_mesh_description/meshes/hand_box.stl"/></geometry>')
    box_urdf = box_urdf.replace('<robot name="fr3"', '<robot name="fr3_boxmesh"')
    (OUT / "fr3_boxmesh.urdf").write_text(box_urdf)
    # 2. first cylinder of link 5 -> prism OBJ (unit radius / unit half length, scaled in the URDF)
    l5 = box_urdf.index('<link name="fr3_link5">')
    mc = re.compile(r'<geometry><cylinder radius="([^"]+)" length="([^"]+)"/></geometry>').search(box_urdf, l5)
    r, length = float(mc.group(1)), float(mc.group(2))
    n = 32
    ang = 2 * np.pi * np.arange(n) / n
    ring = np.stack([np.cos(ang), np.sin(ang)], 1)
    with open(OUT / "meshes" / "prism32.obj", "w") as f:
        f.write("# drc_b200 synthetic mesh: 32-sided unit prism (radius 1, half length 1, axis z)\n")
        for z in (-1.0, 1.0):
            for x, y in ring:
                f.write(f"v {x:.17g} {y:.17g} {z:.17g}\n")
        for i in range(n):
            j = (i + 1) % n
            f.write(f"f {i + 1} {j + 1} {n + j + 1}\nf {i + 1} {n + j + 1} {n + i + 1}\n")
        for i in range(1, n - 1):
            f.write(f"f 1 {i + 2} {i + 1}\nf {n + 1} {n + i + 1} {n + i + 2}\n")
    mesh_urdf = box_urdf[:mc.start()] + f'<geometry><mesh filename="meshes/prism32.obj" scale="{r} {r} {0.5 * length}"/></geometry>' + box_urdf[mc.end():]
    # 3. first sphere of link 0 -> icosphere DAE in millimetres
    l0 = mesh_urdf.index('<link name="fr3_link0">')
    ms = re.compile(r'<geometry><sphere radius="([^"]+)"/></geometry>').search(mesh_urdf, l0)
    rs = float(ms.group(1))
    v, f = icosphere(1000.0 * rs, 2)
    pos = " ".join(f"{x:.17g}" for x in v.ravel())
    idx = " ".join(str(i) for tri in f for i in tri)
    dae = f"""<?xml version="1.0" encoding="utf-8"?>
<COLLADA xmlns="http://www.collada.org/2005/11/COLLADASchema" version="1.4.1">
  <asset><unit name="millimeter" meter="0.001"/><up_axis>Z_UP</up_axis></asset>
  <library_geometries>
    <geometry id="ico" name="ico">
      <mesh>
        <source id="ico-positions">
          <float_array id="ico-positions-array" count="{v.size}">{pos}</float_array>
          <technique_common><accessor source="#ico-positions-array" count="{len(v)}" stride="3"><param name="X" type="float"/><param name="Y" type="float"/><param name="Z" type="float"/></accessor></technique_common>
        </source>
        <vertices id="ico-vertices"><input semantic="POSITION" source="#ico-positions"/></vertices>
        <triangles count="{len(f)}"><input semantic="VERTEX" source="#ico-vertices" offset="0"/><p>{idx}</p></triangles>
      </mesh>
    </geometry>
  </library_geometries>
</COLLADA>
"""
    (PKG / "icosphere.dae").write_text(dae)
    mesh_urdf = mesh_urdf[:ms.start()] + '<geometry><mesh filename="package://fr3_mesh_description/meshes/icosphere.dae"/></geometry>' + mesh_urdf[ms.end():]
    mesh_urdf = mesh_urdf.replace('<robot name="fr3_boxmesh"', '<robot name="fr3_mesh"')
    (OUT / "fr3_mesh.urdf").write_text(mesh_urdf)
    # mesh hulls have no closed forms: every pair with one goes through GJK, and the narrow phase tracks at most 64 such pairs per
    # robot -- one more link pair is disabled to stay below that
    srdf = (SRC / "fr3.srdf").read_text().replace("</robot>", '  <disable_collisions link1="fr3_link2" link2="fr3_link6" reason="synthetic"/>\n</robot>')
    (OUT / "fr3_mesh.srdf").write_text(srdf)
    (OUT / "README.md").write_text("Synthesised by tools/make_mesh_robot.py: the FR3 with collision primitives replaced by generated mesh files "
                                   "(binary STL box, OBJ prism, COLLADA icosphere) for the mesh branch of the model compiler.\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
