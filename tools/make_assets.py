#!/usr/bin/env python
"""Writes the two model files under assets/.

assets/fr3.urdf      -- a dynamics-only description of the Franka FR3 arm: the same <link> and <joint>
                        elements, in the same document order, with the same inertial / origin / axis /
                        limit numbers as the robot description the reference loads (its
                        assets/fr3.urdf; SURVEY.md section 2.1 lists them), and nothing else (no meshes,
                        transmissions or gazebo blocks).  Document order matters: the reference pairs the
                        k-th joint with the k-th link (rigidbody/src/multibody.rs:70-75).
assets/chain32.urdf  -- the synthetic 32-DoF serial revolute chain of BASELINE.json configs[4], defined
                        in SURVEY.md section 8d (the reference cannot represent it).
"""
import math
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HP = "1.5707963267948966"
# (parent xyz, parent rpy, limit effort lower upper velocity, link com xyz, mass, ixx ixy ixz iyy iyz izz)
FR3 = [
    ("0 0 0.333", "0 0 0", (87.0, -2.3093, 2.3093, 2.0),
     "0.003875 0.002081 -0.04762", "4.970684", "0.70337 -0.000139 0.006772 0.70661 0.019169 0.009117"),
    ("0 0 0", f"-{HP} 0 0", (87.0, -1.5133, 1.5133, 1.0),
     "-0.003141 -0.02872  0.003495", "0.646926", "0.007962 -0.003925 0.010254 0.02811 0.000704 0.025995"),
    ("0 -0.316 0", f"{HP} 0 0", (87.0, -2.4937, 2.4937, 1.5),
     "2.7518e-02 3.9252e-02 -6.6502e-02", "3.228604", "0.037242 -0.004761 -0.011396 0.036155 -0.012805 0.01083"),
    ("0.0825 0 0", f"{HP} 0 0", (87.0, -2.7478, -0.4461, 1.25),
     "-5.317e-02 1.04419e-01 2.7454e-02", "3.587895", "0.025853 0.007796 -0.001332 0.019552 0.008641 0.028323"),
    ("-0.0825 0.384 0", f"-{HP} 0 0", (12.0, -2.48, 2.48, 3.0),
     "-1.1953e-02 4.1065e-02 -3.8437e-02", "1.225946", "0.035549 -0.002117 -0.004037 0.029474 0.000229 0.008627"),
    ("0 0 0", f"{HP} 0 0", (12.0, 0.8521, 4.2094, 1.5),
     "6.0149e-02 -1.4117e-02 -1.0517e-02", "1.666555", "0.001964 0.000109 -0.001158 0.004354 0.000341 0.005433"),
    ("0.088 0 0", f"{HP} 0 0", (12.0, -2.6895, 2.6895, 3.0),
     "1.0517e-02 -4.252e-03 6.1597e-02", "0.735522", "0.012516 -0.000428 -0.001196 0.010027 -0.000741 0.004815"),
]
LINK0 = ("-0.041018 -0.00014 0.049974", "0.629769", "0.00315 8.2904e-07 0.00015 0.00388 8.2299e-06 0.004285")


def _inertial(com, mass, six):
    k = six.split()
    # one compact line per link: this file is a parameter table, not a copy of anybody's robot description layout
    return (f'<inertial><mass value="{mass}"/><origin xyz="{com}" rpy="0 0 0"/>'
            f'<inertia ixx="{k[0]}" iyy="{k[3]}" izz="{k[5]}" ixy="{k[1]}" ixz="{k[2]}" iyz="{k[4]}"/></inertial>')


def fr3():
    o = ['<?xml version="1.0" ?>\n',
         '<!-- Dynamics-only FR3 description written by tools/make_assets.py: element order and numbers follow\n'
         '     the franka_description model the reference loads; meshes/transmissions/gazebo are left out. -->\n',
         '<robot name="fr3">\n']

    def sc(k):
        o.append(f'  <link name="fr3_link{k}_sc"/>\n')
        o.append(f'  <joint type="fixed" name="fr3_link{k}_sc_joint"><parent link="fr3_link{k}"/><child link="fr3_link{k}_sc"/>'
                 f'<origin rpy="0 0 0"/></joint>\n')

    o.append('  <link name="fr3_link0">' + _inertial(*LINK0) + '</link>\n')
    sc(0)
    for k, (xyz, rpy, (eff, lo, up, vel), com, mass, six) in enumerate(FR3, start=1):
        o.append(f'  <link name="fr3_link{k}">' + _inertial(com, mass, six) + '</link>\n')
        sc(k)
        o.append(f'  <joint type="revolute" name="fr3_joint{k}"><parent link="fr3_link{k - 1}"/><child link="fr3_link{k}"/>'
                 f'<origin xyz="{xyz}" rpy="{rpy}"/><axis xyz="0 0 1"/>'
                 f'<limit lower="{lo}" upper="{up}" velocity="{vel}" effort="{eff}"/></joint>\n')
    o.append('  <link name="fr3_link8"/>\n')
    o.append('  <joint type="fixed" name="fr3_joint8"><parent link="fr3_link7"/><child link="fr3_link8"/>'
             '<origin xyz="0 0 0.107" rpy="0 0 0"/></joint>\n')
    o.append('  <link name="world"/>\n')
    o.append('  <joint type="fixed" name="world_joint"><parent link="world"/><child link="fr3_link0"/>'
             '<origin xyz="0 0 0" rpy="0 0 0"/></joint>\n')
    o.append('</robot>\n')
    return "".join(o)


def chain32(n=32):
    """SURVEY.md section 8d, config 5.  Links and joints alternate so the k-th joint pairs with the k-th link."""
    o = ['<?xml version="1.0" ?>\n',
         f'<!-- synthetic {n}-DoF serial revolute chain (SURVEY.md 8d config 5); written by tools/make_assets.py -->\n',
         f'<robot name="chain{n}">\n']
    for i in range(n):
        roll = 0.0 if i == 0 else (math.pi / 2 if i % 2 == 1 else -math.pi / 2)
        xyz = (0.05 * (i % 3 == 1), -0.10 * (i % 2 == 1), 0.10 * (i % 2 == 0))
        mass = 1.0 + 0.1 * (i % 5)
        com = (0.01, 0.02 * (-1) ** i, -0.03 + 0.005 * (i % 4))
        ixx, iyy, izz = 0.010 * mass, 0.012 * mass, 0.008 * mass
        six = f"{ixx!r} 0.0001 -0.0002 {iyy!r} 0.0003 {izz!r}"
        o.append(f'  <link name="link{i + 1}">\n'
                 + _inertial(" ".join(repr(float(c)) for c in com), repr(mass), six) + '  </link>\n')
        parent = "base" if i == 0 else f"link{i}"
        o.append(f'  <joint name="joint{i + 1}" type="revolute">\n'
                 f'    <origin rpy="{roll!r} 0 0" xyz="{" ".join(repr(float(c) + 0.0) for c in xyz)}"/>\n'
                 f'    <parent link="{parent}"/>\n    <child link="link{i + 1}"/>\n'
                 f'    <axis xyz="0 0 1"/>\n'
                 f'    <limit effort="50.0" lower="{-math.pi!r}" upper="{math.pi!r}" velocity="2.0"/>\n  </joint>\n')
    o.append('  <link name="base"/>\n</robot>\n')
    return "".join(o)


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "assets"), exist_ok=True)
    with open(os.path.join(ROOT, "assets", "fr3.urdf"), "w") as f:
        f.write(fr3())
    with open(os.path.join(ROOT, "assets", "chain32.urdf"), "w") as f:
        f.write(chain32())
    print("wrote assets/fr3.urdf, assets/chain32.urdf")
