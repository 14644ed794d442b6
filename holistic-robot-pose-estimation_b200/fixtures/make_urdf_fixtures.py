"""Authoring script for the synthetic, mesh-free URDF fixtures in fixtures/urdf/.

The reference does not ship its URDFs (README.md:32 points to a download; lib/config.py:33-36 holds the paths),
so the oracle and the CUDA path share these hand-written kinematic trees.  They are kinematically plausible
(DH-like origins of the public robot descriptions) and carry exactly the link / joint names the reference
looks up: lib/dataset/const.py:58-84 (LINK_NAMES, JOINT_NAMES) and lib/utils/urdf_robot.py:61-65 (Baxter).
Run from this directory: `python make_urdf_fixtures.py`.
"""
from pathlib import Path

OUT = Path(__file__).resolve().parent / "urdf"


def link(n):
    return f'  <link name="{n}"/>\n'


def joint(name, typ, parent, child, xyz="0 0 0", rpy="0 0 0", axis=None, limit=None, mimic=None):
    s = (f'  <joint name="{name}" type="{typ}">\n    <origin xyz="{xyz}" rpy="{rpy}"/>\n'
         f'    <parent link="{parent}"/>\n    <child link="{child}"/>\n')
    if axis:
        s += f'    <axis xyz="{axis}"/>\n'
    if limit:
        s += f'    <limit lower="{limit[0]}" upper="{limit[1]}" effort="100" velocity="2"/>\n'
    if mimic:
        s += f'    <mimic joint="{mimic[0]}" multiplier="{mimic[1]}" offset="{mimic[2]}"/>\n'
    return s + '  </joint>\n'


H = "1.5707963267948966"
Q = "0.7853981633974483"
PI = "3.141592653589793"


def panda():
    s = '<?xml version="1.0"?>\n<robot name="panda">\n'
    for n in ["panda_link%d" % i for i in range(9)] + ["panda_hand", "panda_leftfinger", "panda_rightfinger"]:
        s += link(n)
    P = [("0 0 0.333", "0 0 0", (-2.8973, 2.8973)), ("0 0 0", f"-{H} 0 0", (-1.7628, 1.7628)),
         ("0 -0.316 0", f"{H} 0 0", (-2.8973, 2.8973)), ("0.0825 0 0", f"{H} 0 0", (-3.0718, -0.0698)),
         ("-0.0825 0.384 0", f"-{H} 0 0", (-2.8973, 2.8973)), ("0 0 0", f"{H} 0 0", (-0.0175, 3.7525)),
         ("0.088 0 0", f"{H} 0 0", (-2.8973, 2.8973))]
    for i, (xyz, rpy, lim) in enumerate(P, start=1):
        s += joint(f"panda_joint{i}", "revolute", f"panda_link{i-1}", f"panda_link{i}", xyz, rpy, "0 0 1", lim)
    s += joint("panda_joint8", "fixed", "panda_link7", "panda_link8", "0 0 0.107")
    s += joint("panda_hand_joint", "fixed", "panda_link8", "panda_hand", "0 0 0", f"0 0 -{Q}")
    s += joint("panda_finger_joint1", "prismatic", "panda_hand", "panda_leftfinger", "0 0 0.0584", "0 0 0", "0 1 0",
               (0.0, 0.04))
    s += joint("panda_finger_joint2", "prismatic", "panda_hand", "panda_rightfinger", "0 0 0.0584", "0 0 0", "0 -1 0",
               (0.0, 0.04), ("panda_finger_joint1", 1.0, 0.0))
    return s + "</robot>\n"


def kuka():
    s = '<?xml version="1.0"?>\n<robot name="iiwa7">\n'
    for n in ["iiwa_link_%d" % i for i in range(8)] + ["iiwa_link_ee"]:
        s += link(n)
    K = [("0 0 0.15", "0 0 0", 2.9671), ("0 0 0.19", f"{H} 0 {PI}", 2.0944), ("0 0.21 0", f"{H} 0 {PI}", 2.9671),
         ("0 0 0.19", f"{H} 0 0", 2.0944), ("0 0.21 0", f"-{H} {PI} 0", 2.9671),
         ("0 0.0607 0.19", f"{H} 0 0", 2.0944), ("0 0.081 0.0607", f"-{H} {PI} 0", 3.0543)]
    for i, (xyz, rpy, lim) in enumerate(K, start=1):
        s += joint(f"iiwa_joint_{i}", "revolute", f"iiwa_link_{i-1}", f"iiwa_link_{i}", xyz, rpy, "0 0 1", (-lim, lim))
    s += joint("iiwa_joint_ee", "fixed", "iiwa_link_7", "iiwa_link_ee", "0 0 0.045")
    return s + "</robot>\n"


def baxter():
    s = '<?xml version="1.0"?>\n<robot name="baxter">\n'
    arm_links = ["arm_mount", "upper_shoulder", "lower_shoulder", "upper_elbow", "lower_elbow", "upper_forearm",
                 "lower_forearm", "wrist", "hand"]
    for n in ["base", "torso", "head"] + [f"{sd}_{l}" for sd in ("right", "left") for l in arm_links]:
        s += link(n)
    s += joint("torso_t0", "fixed", "base", "torso")
    s += joint("head_pan", "revolute", "torso", "head", "0.06 0 0.686", "0 0 0", "0 0 1", (-1.5708, 1.5708))
    A = [("s0", "0.055695 0 0.011038", "0 0 0", (-1.7017, 1.7017)), ("s1", "0.069 0 0.27035", f"-{H} 0 0", (-2.147, 1.047)),
         ("e0", "0.102 0 0", f"{H} 0 {H}", (-3.0542, 3.0542)), ("e1", "0.069 0 0.26242", f"-{H} -{H} 0", (-0.05, 2.618)),
         ("w0", "0.10359 0 0", f"{H} 0 {H}", (-3.059, 3.059)), ("w1", "0.01 0 0.2707", f"-{H} -{H} 0", (-1.5708, 2.094)),
         ("w2", "0.115975 0 0", f"{H} 0 {H}", (-3.059, 3.059))]
    for sd, ysign, yaw in (("right", "-", f"-{Q}"), ("left", "", Q)):
        s += joint(f"{sd}_torso_arm_mount", "fixed", "torso", f"{sd}_arm_mount",
                   f"0.024645 {ysign}0.219645 0.118588", f"0 0 {yaw}")
        for i, (jn, xyz, rpy, lim) in enumerate(A):
            s += joint(f"{sd}_{jn}", "revolute", f"{sd}_{arm_links[i]}", f"{sd}_{arm_links[i+1]}", xyz, rpy, "0 0 1", lim)
        s += joint(f"{sd}_hand", "fixed", f"{sd}_wrist", f"{sd}_hand", "0 0 0.11355")
    return s + "</robot>\n"


if __name__ == "__main__":
    OUT.mkdir(exist_ok=True)
    (OUT / "panda.urdf").write_text(panda())
    (OUT / "iiwa7.urdf").write_text(kuka())
    (OUT / "baxter.urdf").write_text(baxter())
