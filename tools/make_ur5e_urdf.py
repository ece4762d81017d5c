#!/usr/bin/env python3
"""Synthesize robots/ur5e/ur5e.urdf for BASELINE config 2 (UR5e CLIK + OSF).  NOT from the reference (it ships no UR5e
model): kinematics from the public UR5e DH table (d1 .1625, a2 -.425, a3 -.3922, d4 .1333, d5 .0997, d6 .0996), link
masses 3.761 / 8.058 / 2.846 / 1.37 / 1.3 / 0.365 kg with solid-cylinder inertias about illustrative centres of mass,
primitive collision shapes (spheres / cylinders) for the self-distance getter.  Joint range +-2*pi*0.9, 3.14 rad/s."""
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "dyros_robot_controller_b200" / "robots" / "ur5e"
PI2 = 1.5707963267948966


def cyl_inertia(m, r, l):
    ixx = m * (3 * r * r + l * l) / 12.0
    return ixx, ixx, 0.5 * m * r * r


def link(name, mass, com, r, l, shapes):
    ixx, iyy, izz = cyl_inertia(mass, r, l)
    s = [f'  <link name="{name}">',
         f'    <inertial><origin xyz="{com[0]} {com[1]} {com[2]}" rpy="0 0 0"/><mass value="{mass}"/>'
         f'<inertia ixx="{ixx:.8f}" ixy="0" ixz="0" iyy="{iyy:.8f}" iyz="0" izz="{izz:.8f}"/></inertial>']
    for kind, xyz, rpy, dims in shapes:
        geo = f'<sphere radius="{dims[0]}"/>' if kind == "sphere" else f'<cylinder radius="{dims[0]}" length="{dims[1]}"/>'
        s.append(f'    <collision><origin xyz="{xyz[0]} {xyz[1]} {xyz[2]}" rpy="{rpy[0]} {rpy[1]} {rpy[2]}"/><geometry>{geo}</geometry></collision>')
    s.append("  </link>")
    return "\n".join(s)


def joint(name, parent, child, xyz, rpy, lim=5.6549, vel=3.14, eff=150):
    return (f'  <joint name="{name}" type="revolute"><parent link="{parent}"/><child link="{child}"/>'
            f'<origin xyz="{xyz[0]} {xyz[1]} {xyz[2]}" rpy="{rpy[0]} {rpy[1]} {rpy[2]}"/><axis xyz="0 0 1"/>'
            f'<limit lower="{-lim}" upper="{lim}" velocity="{vel}" effort="{eff}"/></joint>')


def main():
    parts = ['<?xml version="1.0"?>', '<!-- synthesized by tools/make_ur5e_urdf.py: NOT part of the reference -->', '<robot name="ur5e">',
             link("base_link", 4.0, (0, 0, 0.05), 0.075, 0.1, [("cylinder", (0, 0, 0.05), (0, 0, 0), (0.075, 0.1))]),
             joint("shoulder_pan_joint", "base_link", "shoulder_link", (0, 0, 0.1625), (0, 0, 0)),
             link("shoulder_link", 3.761, (0, -0.00193, -0.02561), 0.06, 0.15, [("sphere", (0, 0, 0), (0, 0, 0), (0.075,))]),
             joint("shoulder_lift_joint", "shoulder_link", "upper_arm_link", (0, 0, 0), (PI2, 0, 0)),
             link("upper_arm_link", 8.058, (-0.2125, 0, 0.11336), 0.055, 0.425,
                  [("cylinder", (-0.2125, 0, 0.138), (0, PI2, 0), (0.055, 0.425)), ("sphere", (0, 0, 0.138), (0, 0, 0), (0.07,))]),
             joint("elbow_joint", "upper_arm_link", "forearm_link", (-0.425, 0, 0), (0, 0, 0), lim=3.1416 * 0.9),
             link("forearm_link", 2.846, (-0.15, 0, 0.0265), 0.04, 0.3922,
                  [("cylinder", (-0.1961, 0, 0.007), (0, PI2, 0), (0.04, 0.3922)), ("sphere", (0, 0, 0.007), (0, 0, 0), (0.06,))]),
             joint("wrist_1_joint", "forearm_link", "wrist_1_link", (-0.3922, 0, 0.1333), (0, 0, 0), vel=3.14, eff=28),
             link("wrist_1_link", 1.37, (0, -0.01634, -0.0018), 0.04, 0.12, [("sphere", (0, 0, 0), (0, 0, 0), (0.05,))]),
             joint("wrist_2_joint", "wrist_1_link", "wrist_2_link", (0, -0.0997, 0), (PI2, 0, 0), eff=28),
             link("wrist_2_link", 1.3, (0, 0.01634, -0.0018), 0.04, 0.12, [("sphere", (0, 0, 0), (0, 0, 0), (0.05,))]),
             joint("wrist_3_joint", "wrist_2_link", "wrist_3_link", (0, 0.0996, 0), (-PI2, 0, 0), eff=28),
             link("wrist_3_link", 0.365, (0, 0, -0.001159), 0.035, 0.04, [("cylinder", (0, 0, -0.02), (0, 0, 0), (0.035, 0.04))]),
             '  <joint name="flange" type="fixed"><parent link="wrist_3_link"/><child link="tool0"/><origin xyz="0 0 0" rpy="0 0 0"/></joint>',
             '  <link name="tool0"/>', "</robot>"]
    OUT.mkdir(exist_ok=True)
    (OUT / "ur5e.urdf").write_text("\n".join(parts) + "\n")
    adj = [("base_link", "shoulder_link"), ("shoulder_link", "upper_arm_link"), ("upper_arm_link", "forearm_link"),
           ("forearm_link", "wrist_1_link"), ("wrist_1_link", "wrist_2_link"), ("wrist_2_link", "wrist_3_link"),
           ("base_link", "upper_arm_link"), ("wrist_1_link", "wrist_3_link"), ("forearm_link", "wrist_2_link")]
    srdf = ['<?xml version="1.0"?>', '<robot name="ur5e">'] + [f'  <disable_collisions link1="{a}" link2="{b}"/>' for a, b in adj] + ["</robot>"]
    (OUT / "ur5e.srdf").write_text("\n".join(srdf) + "\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
