"""Robot constants of the hot path (values of lib/dataset/const.py:58-84,168-246 and lib/utils/urdf_robot.py:61-65
of the reference; these are data, the reference's names are part of the drop-in contract)."""

LINK_NAMES = {
    "panda": ["panda_link0", "panda_link2", "panda_link3", "panda_link4", "panda_link6", "panda_link7",
              "panda_hand"],
    "kuka": ["iiwa_link_0", "iiwa_link_1", "iiwa_link_2", "iiwa_link_3", "iiwa_link_4", "iiwa_link_5",
             "iiwa_link_6", "iiwa_link_7"],
    "baxter": ["torso", "right_upper_shoulder", "left_upper_shoulder", "right_lower_shoulder",
               "left_lower_shoulder", "right_upper_elbow", "left_upper_elbow", "right_lower_elbow",
               "left_lower_elbow", "right_upper_forearm", "left_upper_forearm", "right_lower_forearm",
               "left_lower_forearm", "right_wrist", "left_wrist", "right_hand", "left_hand"],
}

JOINT_NAMES = {
    "panda": ["panda_joint1", "panda_joint2", "panda_joint3", "panda_joint4", "panda_joint5", "panda_joint6",
              "panda_joint7", "panda_finger_joint1"],
    "kuka": ["iiwa_joint_1", "iiwa_joint_2", "iiwa_joint_3", "iiwa_joint_4", "iiwa_joint_5", "iiwa_joint_6",
             "iiwa_joint_7"],
    "baxter": ["head_pan", "right_s0", "left_s0", "right_s1", "left_s1", "right_e0", "left_e0", "right_e1",
               "left_e1", "right_w0", "left_w0", "right_w1", "left_w1", "right_w2", "left_w2"],
}

# joints whose parent link / origin define Baxter's keypoints (urdf_robot.py:61-65)
BAXTER_KEYPOINT_JOINTS = ["torso_t0", "right_s0", "left_s0", "right_s1", "left_s1", "right_e0", "left_e0",
                          "right_e1", "left_e1", "right_w0", "left_w0", "right_w1", "left_w1", "right_w2",
                          "left_w2", "right_hand", "left_hand"]

# INITIAL_JOINT_ANGLE["mean"] in JOINT_NAMES order (const.py:168-211)
INIT_POSE_MEAN = {
    "panda": [0.0, 0.0, 0.0, -1.52715, 0.0, 1.8675, 0.0, 0.02],
    "kuka": [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
    "baxter": [0.0, 0.0, 0.0, -0.5499999999999999, -0.5499999999999999, 0.0, 0.0, 1.284, 1.284, 0.0, 0.0,
               0.2616018366049999, 0.2616018366049999, 0.0, 0.0],
}
INIT_POSE_ZERO = {k: [0.0] * len(v) for k, v in INIT_POSE_MEAN.items()}

JOINT_BOUNDS = {
    "panda": [[-2.9671, 2.9671], [-1.8326, 1.8326], [-2.9671, 2.9671], [-3.1416, 0.0873], [-2.9671, 2.9671],
              [-0.0873, 3.8223], [-2.9671, 2.9671], [0.0, 0.04]],
    "kuka": [[-2.9671, 2.9671], [-2.0944, 2.0944], [-2.9671, 2.9671], [-2.0944, 2.0944], [-2.9671, 2.9671],
             [-2.0944, 2.0944], [-3.0543, 3.0543]],
    "baxter": [[-1.5708, 1.5708], [-1.7017, 1.7017], [-1.7017, 1.7017], [-2.147, 1.047], [-2.147, 1.047],
               [-3.0542, 3.0542], [-3.0542, 3.0542], [-0.05, 2.618], [-0.05, 2.618], [-3.059, 3.059],
               [-3.059, 3.059], [-1.5708, 2.094], [-1.5708, 2.094], [-3.059, 3.059], [-3.059, 3.059]],
}
