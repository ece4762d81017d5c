#!/usr/bin/env python3
"""Import the FR3 robot description (model DATA, not code) from the reference checkout.

Reads   /root/reference/examples/robots/fr3/fr3.urdf  and  fr3.srdf
Writes  dyros_robot_controller_b200/robots/fr3/fr3.urdf  and  fr3.srdf

Only what the hot path consumes is kept: links (inertial + primitive collision
geometry), joints (origin, axis, limits) and the SRDF disabled pairs.  Visual
meshes, safety controllers and joint-dynamics tags are dropped (Pinocchio's
URDF loader ignores them too, see reference src/manipulator/robot_data.cpp:21-43).
The numbers are written verbatim (string copy) so no precision is lost.

Run in the build container only (the GPU box has no /root/reference).
"""
import sys
import xml.etree.ElementTree as ET
from pathlib import Path

REF = Path("/root/reference/examples/robots/fr3")
OUT = Path(__file__).resolve().parents[1] / "dyros_robot_controller_b200" / "robots" / "fr3"


def strip_urdf(src: Path) -> str:
    root = ET.parse(src).getroot()
    out = ['<?xml version="1.0"?>',
           '<!-- FR3 model data imported by tools/import_reference_model.py (inertial + primitive collision + joints only) -->',
           f'<robot name="{root.get("name")}">']
    for el in root:
        if el.tag == "link":
            out.append(f'  <link name="{el.get("name")}">')
            ine = el.find("inertial")
            if ine is not None:
                o = ine.find("origin")
                i = ine.find("inertia")
                out.append("    <inertial>")
                out.append(f'      <origin xyz="{o.get("xyz")}" rpy="{o.get("rpy")}"/>')
                out.append(f'      <mass value="{ine.find("mass").get("value")}"/>')
                out.append("      <inertia " + " ".join(f'{k}="{i.get(k)}"' for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz")) + "/>")
                out.append("    </inertial>")
            for c in el.findall("collision"):
                o = c.find("origin")
                g = list(c.find("geometry"))[0]
                attrs = " ".join(f'{k}="{v.strip()}"' for k, v in g.attrib.items())
                out.append("    <collision>")
                out.append(f'      <origin xyz="{o.get("xyz")}" rpy="{o.get("rpy")}"/>')
                out.append(f"      <geometry><{g.tag} {attrs}/></geometry>")
                out.append("    </collision>")
            out.append("  </link>")
        elif el.tag == "joint":
            out.append(f'  <joint name="{el.get("name")}" type="{el.get("type")}">')
            o = el.find("origin")
            out.append(f'    <origin xyz="{o.get("xyz")}" rpy="{o.get("rpy")}"/>')
            out.append(f'    <parent link="{el.find("parent").get("link")}"/>')
            out.append(f'    <child link="{el.find("child").get("link")}"/>')
            if el.find("axis") is not None:
                out.append(f'    <axis xyz="{el.find("axis").get("xyz")}"/>')
            lim = el.find("limit")
            if lim is not None:
                out.append("    <limit " + " ".join(f'{k}="{lim.get(k)}"' for k in ("lower", "upper", "velocity", "effort")) + "/>")
            out.append("  </joint>")
    out.append("</robot>")
    return "\n".join(out) + "\n"


def strip_srdf(src: Path) -> str:
    root = ET.parse(src).getroot()
    out = ['<?xml version="1.0"?>', f'<robot name="{root.get("name")}">']
    for gs in root.findall("group_state"):
        out.append(f'  <group_state name="{gs.get("name")}" group="{gs.get("group")}">')
        for j in gs.findall("joint"):
            out.append(f'    <joint name="{j.get("name")}" value="{j.get("value")}"/>')
        out.append("  </group_state>")
    for d in root.findall("disable_collisions"):
        out.append(f'  <disable_collisions link1="{d.get("link1")}" link2="{d.get("link2")}"/>')
    out.append("</robot>")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    if not REF.exists():
        sys.exit("reference checkout not present; nothing to import")
    OUT.mkdir(parents=True, exist_ok=True)
    (OUT / "fr3.urdf").write_text(strip_urdf(REF / "fr3.urdf"))
    (OUT / "fr3.srdf").write_text(strip_srdf(REF / "fr3.srdf"))
    print("wrote", OUT)
