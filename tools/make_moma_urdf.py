#!/usr/bin/env python3
"""Synthesize the two mobile-manipulator models of BASELINE configs 4-5 (NOT from the reference, which ships no
mobile-manipulator URDF): the reference's FR3 arm (examples/robots/fr3/fr3.urdf) mounted on
  * husky_fr3: differential-drive base, 2 wheel joints, r = 0.1651 m, track 0.555 m (public Husky A200 figures, synthetic inertia)
  * pcv_fr3:   powered-caster base, 2 casters = 4 joints (steer about z, roll about y, offset 0.020 m, r = 0.055 m); synthetic
  * xls_fr3:   mecanum base, 4 wheel joints; base mass/inertia and wheel mass from the reference's
               examples/robots/xls/summit_xls.xml:38,58-61, wheel parameters from examples/C++/src/xls_controller.cpp:18-27
behind three virtual joints (prismatic x, prismatic y, revolute z) as MobileManipulator::RobotData requires
(src/mobile_manipulator/robot_data.cpp:19,115).  Joint order: virtual(0..2), wheels(3..3+w-1), arm(3+w..).

    python tools/make_moma_urdf.py     # rewrites dyros_robot_controller_b200/robots/{husky_fr3,xls_fr3}/
"""
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
ROBOTS = ROOT / "dyros_robot_controller_b200" / "robots"


def inertial(mass, ixx, iyy, izz, xyz="0 0 0"):
    return (f'<inertial><origin xyz="{xyz}" rpy="0 0 0"/><mass value="{mass}"/>'
            f'<inertia ixx="{ixx}" ixy="0" ixz="0" iyy="{iyy}" iyz="0" izz="{izz}"/></inertial>')


def build(name, base_mass, base_inertia, base_box, base_z, wheels, wheel_mass, wheel_r, mount_xyz, caster_offset=None):
    fr3 = (ROBOTS / "fr3" / "fr3.urdf").read_text()
    body = fr3[fr3.index("<robot"):]
    body = body[body.index(">") + 1: body.rindex("</robot>")].replace('"base_link"', '"fr3_base"')
    srdf = (ROBOTS / "fr3" / "fr3.srdf").read_text()
    dis = [l for l in srdf.splitlines() if "disable_collisions" in l]
    out = ['<?xml version="1.0"?>', f'<!-- synthesized by tools/make_moma_urdf.py: NOT part of the reference -->', f'<robot name="{name}">',
           '  <link name="world"/>', '  <link name="virtual_x_link"/>', '  <link name="virtual_y_link"/>',
           '  <joint name="virtual_x" type="prismatic"><parent link="world"/><child link="virtual_x_link"/><axis xyz="1 0 0"/>'
           '<limit lower="-100" upper="100" velocity="2" effort="1000"/></joint>',
           '  <joint name="virtual_y" type="prismatic"><parent link="virtual_x_link"/><child link="virtual_y_link"/><axis xyz="0 1 0"/>'
           '<limit lower="-100" upper="100" velocity="2" effort="1000"/></joint>',
           '  <joint name="virtual_yaw" type="continuous"><parent link="virtual_y_link"/><child link="mobile_base"/><axis xyz="0 0 1"/>'
           '<limit velocity="2" effort="1000"/></joint>',
           f'  <link name="mobile_base">{inertial(base_mass, *base_inertia, xyz=f"0 0 {base_z}")}'
           f'<collision><origin xyz="0 0 {base_z}" rpy="0 0 0"/><geometry><box size="{base_box[0]} {base_box[1]} {base_box[2]}"/></geometry></collision></link>']
    iw = 0.5 * wheel_mass * wheel_r ** 2
    for wname, (x, y) in wheels:
        if caster_offset is not None:
            # powered caster: steering joint about z on the base, rolling joint about y on the steering fork, trailing the
            # steering axis by the caster offset b (joint order steer, roll = the reference's wheel_pos layout,
            # mobile/robot_data.cpp:181-182)
            h = 2.0 * wheel_r
            out.append(f'  <joint name="{wname}_steer" type="continuous"><parent link="mobile_base"/><child link="{wname}_fork"/>'
                       f'<origin xyz="{x} {y} {h}" rpy="0 0 0"/><axis xyz="0 0 1"/><limit velocity="20" effort="100"/></joint>')
            out.append(f'  <link name="{wname}_fork">{inertial(0.5, 0.001, 0.001, 0.001)}</link>')
            out.append(f'  <joint name="{wname}_roll" type="continuous"><parent link="{wname}_fork"/><child link="{wname}_link"/>'
                       f'<origin xyz="{-caster_offset} 0 {-(h - wheel_r)}" rpy="0 0 0"/><axis xyz="0 1 0"/><limit velocity="40" effort="100"/></joint>')
            out.append(f'  <link name="{wname}_link">{inertial(wheel_mass, 0.5 * iw + wheel_mass * 0.001, iw, 0.5 * iw + wheel_mass * 0.001)}</link>')
            continue
        out.append(f'  <joint name="{wname}" type="continuous"><parent link="mobile_base"/><child link="{wname}_link"/>'
                   f'<origin xyz="{x} {y} {wheel_r}" rpy="0 0 0"/><axis xyz="0 1 0"/><limit velocity="20" effort="100"/></joint>')
        out.append(f'  <link name="{wname}_link">{inertial(wheel_mass, 0.5 * iw + wheel_mass * 0.01, iw, 0.5 * iw + wheel_mass * 0.01)}</link>')
    out.append(f'  <joint name="arm_mount" type="fixed"><parent link="mobile_base"/><child link="fr3_base"/>'
               f'<origin xyz="{mount_xyz[0]} {mount_xyz[1]} {mount_xyz[2]}" rpy="0 0 0"/></joint>')
    out.append(body)
    out.append("</robot>")
    d = ROBOTS / name
    d.mkdir(exist_ok=True)
    (d / f"{name}.urdf").write_text("\n".join(out) + "\n")
    extra = [f'  <disable_collisions link1="mobile_base" link2="fr3_link{i}"/>' for i in (0, 1, 2)]
    (d / f"{name}.srdf").write_text('<?xml version="1.0"?>\n<robot name="%s">\n%s\n%s\n</robot>\n' % (name, "\n".join(dis), "\n".join(extra)))
    print("wrote", d)


if __name__ == "__main__":
    build("husky_fr3", 46.0, (0.6022, 1.7386, 2.0296), (0.99, 0.67, 0.25), 0.20,
          [("wheel_left", (0.0, 0.2775)), ("wheel_right", (0.0, -0.2775))], 2.637, 0.1651, (0.30, 0.0, 0.34))
    build("xls_fr3", 125.0, (1.391, 6.853, 6.125), (0.72, 0.61, 0.30), 0.25,
          [("wheel_fl", (0.2225, 0.2045)), ("wheel_fr", (0.2225, -0.2045)), ("wheel_rl", (-0.2225, 0.2045)),
           ("wheel_rr", (-0.2225, -0.2045))], 6.5, 0.120, (0.25, 0.0, 0.42))
    # powered-caster base (Holmberg & Khatib's PCV class): two casters, 4 joints (steer, roll each); all figures synthetic
    build("pcv_fr3", 80.0, (1.2, 3.4, 4.1), (0.70, 0.50, 0.25), 0.25,
          [("caster_front", (0.215, 0.125)), ("caster_rear", (-0.215, -0.125))], 1.2, 0.055, (0.20, 0.0, 0.40),
          caster_offset=0.020)
