"""CPU oracle for the HoRoPose inference forward pass -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain PyTorch-CPU / numpy restatement of the reference's algorithm for the hot path,
driven directly by a reference-keyed `state_dict`.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` leg may import this module; the product path never does.

Parity pinning: the reference ships no tests or golden vectors, so this oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF, run in the build container through oracle/ref_harness.py; the
vectors live in tests/golden/*.npz with the generating script tests/golden/make_golden.py, and
tests/test_oracle_golden.py re-checks this file against them on every run.

Every function cites the reference file:line it restates (paths relative to the reference root).
"""
from __future__ import annotations

import math
import xml.etree.ElementTree as ET
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------------------
# tables: lib/dataset/const.py:58-84 (names), full_net.py:42-51 (DoF / nkpt), configs/*/full.yaml (ref kpt)
# ----------------------------------------------------------------------------------------------------------
LINK_NAMES = {
    "panda": ["panda_link0", "panda_link2", "panda_link3", "panda_link4", "panda_link6", "panda_link7", "panda_hand"],
    "kuka": ["iiwa_link_%d" % i for i in range(8)],
}
JOINT_NAMES = {
    "panda": ["panda_joint%d" % i for i in range(1, 8)] + ["panda_finger_joint1"],
    "kuka": ["iiwa_joint_%d" % i for i in range(1, 8)],
    "baxter": ["head_pan", "right_s0", "left_s0", "right_s1", "left_s1", "right_e0", "left_e0", "right_e1", "left_e1",
               "right_w0", "left_w0", "right_w1", "left_w1", "right_w2", "left_w2"],
}
BAXTER_KP_JOINTS = ["torso_t0", "right_s0", "left_s0", "right_s1", "left_s1", "right_e0", "left_e0", "right_e1",
                    "left_e1", "right_w0", "left_w0", "right_w1", "left_w1", "right_w2", "left_w2", "right_hand",
                    "left_hand"]  # lib/utils/urdf_robot.py:61-65
ROBOTS = {"panda": (8, 7, 3), "kuka": (7, 8, 3), "baxter": (15, 17, 0)}
INIT_POSE_MEAN = {  # lib/dataset/const.py:168-211 ("mean"), in JOINT_NAMES order
    "panda": [0.0, 0.0, 0.0, -1.52715, 0.0, 1.8675, 0.0, 0.02],
    "kuka": [0.0] * 7,
    "baxter": [0.0, 0.0, 0.0, -0.5499999999999999, -0.5499999999999999, 0.0, 0.0, 1.284, 1.284, 0.0, 0.0,
               0.2616018366049999, 0.2616018366049999, 0.0, 0.0],
}


# ----------------------------------------------------------------------------------------------------------
# URDF -> kinematic tree (lib/utils/urdfpytorch/urdf.py:2399-2409, 3721-3751; utils.py:22-51,142-167)
# ----------------------------------------------------------------------------------------------------------
def rpy_to_matrix(rpy):
    """urdfpytorch/utils.py:22-51 (Z1-Y2-X3 Tait-Bryan), float64."""
    c3, c2, c1 = np.cos(np.asarray(rpy, dtype=np.float64))
    s3, s2, s1 = np.sin(np.asarray(rpy, dtype=np.float64))
    return np.array([
        [c1 * c2, (c1 * s2 * s3) - (c3 * s1), (s1 * s3) + (c1 * c3 * s2)],
        [c2 * s1, (c1 * c3) + (s1 * s2 * s3), (c3 * s1 * s2) - (c1 * s3)],
        [-s2, c2 * s3, c2 * c3]], dtype=np.float64)


class OracleRobot:
    """Restates URDF.load + graph build (urdf.py:2746-2772), actuated-joint ordering (urdf.py:3795-3813, with the
    stable sort the authors' numpy produced -- SURVEY.md fact 11) and URDFRobot's keypoint table
    (lib/utils/urdf_robot.py:22-80)."""

    def __init__(self, robot_type: str, urdf_path: str):
        self.robot_type = robot_type
        self.dof = ROBOTS[robot_type][0]
        root = ET.parse(urdf_path).getroot()
        self.links = [l.attrib["name"] for l in root.findall("link")]
        self.joints = []
        for j in root.findall("joint"):
            origin = np.eye(4, dtype=np.float64)
            o = j.find("origin")
            if o is not None:  # utils.py:142-167
                if "xyz" in o.attrib:
                    origin[:3, 3] = np.array(o.attrib["xyz"].split(), dtype=np.float64)
                if "rpy" in o.attrib:
                    origin[:3, :3] = rpy_to_matrix(np.array(o.attrib["rpy"].split(), dtype=np.float64))
            ax = j.find("axis")
            axis = np.array(ax.attrib["xyz"].split(), dtype=np.float64) if ax is not None else np.array([1.0, 0, 0])
            axis = axis / np.linalg.norm(axis)  # urdf.py:2169
            mim = j.find("mimic")
            self.joints.append(dict(
                name=j.attrib["name"], type=j.attrib["type"], parent=j.find("parent").attrib["link"],
                child=j.find("child").attrib["link"], origin=origin, axis=axis,
                mimic=None if mim is None else (mim.attrib["joint"], float(mim.attrib.get("multiplier", 1.0)),
                                                float(mim.attrib.get("offset", 0.0)))))
        self.joint_of_child = {j["child"]: j for j in self.joints}
        children = set(self.joint_of_child)
        bases = [l for l in self.links if l not in children]
        assert len(bases) == 1
        self.base = bases[0]

        def depth(link):
            d = 1
            while link != self.base:
                link = self.joint_of_child[link]["parent"]
                d += 1
            return d

        act = [j for j in self.joints if j["mimic"] is None and j["type"] != "fixed"]  # urdf.py:3788-3790
        order = np.argsort([depth(j["child"]) for j in act], kind="stable")          # urdf.py:3807-3813
        self.actuated = [act[i] for i in order]
        self.qcol = {j["name"]: i for i, j in enumerate(self.actuated)}
        # topological order: parents before children (urdf.py:2772, reverse topological sort)
        self.topo = sorted(self.links, key=depth)
        # keypoints (urdf_robot.py:52-74)
        if robot_type in ("panda", "kuka"):
            self.link_names = list(LINK_NAMES[robot_type])
            self.offsets = np.zeros((len(self.link_names), 3), dtype=np.float64)
        else:
            jm = {j["name"]: j for j in self.joints}
            self.link_names = [jm[n]["parent"] for n in BAXTER_KP_JOINTS]
            self.offsets = np.stack([jm[n]["origin"][:3, 3] for n in BAXTER_KP_JOINTS])

    # urdf.py:2427-2462 (Rodrigues), :2344-2396 (child pose), :3061-3149 (tree walk)
    def link_fk(self, q: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
        B, dt = q.shape[0], q.dtype
        fk = OrderedDict()
        for link in self.topo:
            if link == self.base:
                fk[link] = torch.eye(4, dtype=dt).repeat(B, 1, 1)
                continue
            j = self.joint_of_child[link]
            origin = torch.as_tensor(j["origin"]).to(dt)
            cfg = None
            if j["mimic"] is not None:
                mj, mult, off = j["mimic"]
                cfg = mult * q[:, self.qcol[mj]] + off                       # urdf.py:3128-3132
            elif j["name"] in self.qcol:
                cfg = q[:, self.qcol[j["name"]]]
            if cfg is None or j["type"] == "fixed":
                child = origin.repeat(B, 1, 1)                               # urdf.py:2377-2379
            elif j["type"] in ("revolute", "continuous"):
                axis = j["axis"]
                sina, cosa = torch.sin(cfg), torch.cos(cfg)
                M = torch.eye(4, dtype=dt).repeat(B, 1, 1)
                M[:, 0, 0] = cosa
                M[:, 1, 1] = cosa
                M[:, 2, 2] = cosa
                M[:, :3, :3] += torch.as_tensor(np.outer(axis, axis)).to(dt) * (1.0 - cosa)[:, None, None]
                M[:, :3, :3] += torch.as_tensor(np.array([[0.0, -axis[2], axis[1]], [axis[2], 0.0, -axis[0]],
                                                          [-axis[1], axis[0], 0.0]])).to(dt) * sina[:, None, None]
                child = torch.matmul(origin, M)                              # urdf.py:2383
            elif j["type"] == "prismatic":
                T = torch.eye(4, dtype=dt).repeat(B, 1, 1)
                T[:, :3, 3] = torch.as_tensor(j["axis"]).to(dt) * cfg[:, None]
                child = torch.matmul(origin, T)                              # urdf.py:2388-2390
            else:
                raise NotImplementedError(j["type"])
            fk[link] = torch.matmul(fk[j["parent"]], child)                  # urdf.py:3137-3139
        return fk

    def get_TWL(self, q):  # urdf_robot.py:107-111
        fk = self.link_fk(q)
        return torch.stack([fk[n] for n in self.link_names]).permute(1, 0, 2, 3)

    def _base2cam(self, rot, trans):  # urdf_robot.py:86-100
        B = rot.shape[0]
        if rot.shape[1] == 6:
            R = rot6d_to_rotmat(rot)
        elif rot.shape[1] == 4:
            R = quat_to_rotmat(rot)
        else:
            raise NotImplementedError
        T = torch.zeros(B, 4, 4, dtype=rot.dtype)
        T[:, :3, :3] = R
        T[:, :3, 3] = trans
        T[:, 3, 3] = 1.0
        return T.unsqueeze(1)

    def _pts(self, TWL):  # urdf_robot.py:104,198
        off = torch.as_tensor(self.offsets).to(TWL.dtype)[None, :, :, None]
        return (TWL[:, :, :3, :3] @ off + TWL[:, :, :3, 3:4]).squeeze(-1)

    def get_keypoints(self, q, rot, trans):  # urdf_robot.py:82-105
        return self._pts(self._base2cam(rot, trans) @ self.get_TWL(q))

    def get_keypoints_only_fk(self, q):  # urdf_robot.py:141-149
        return self._pts(self.get_TWL(q))

    def get_keypoints_root(self, q, rot, trans, root=0):  # urdf_robot.py:169-199
        if root == 0:
            return self.get_keypoints(q, rot, trans)
        TWL = self.get_TWL(q)
        TWL = torch.linalg.inv(TWL[:, root:root + 1]) @ TWL
        return self._pts(self._base2cam(rot, trans) @ TWL)

    def get_rotation_at_specific_root(self, q, rot, trans, root=0):  # urdf_robot.py:113-138
        if root == 0:
            return rot
        TWL = self._base2cam(rot, trans) @ self.get_TWL(q)
        return TWL[:, root, :2, :3].reshape(-1, 6)


# ----------------------------------------------------------------------------------------------------------
# rotation representations (lib/utils/geometries.py)
# ----------------------------------------------------------------------------------------------------------
def rot6d_to_rotmat(poses):  # geometries.py:100-115 -- rows of the result are x, y, z
    x_raw, y_raw = poses[..., 0:3], poses[..., 3:6]
    x = x_raw / torch.norm(x_raw, p=2, dim=-1, keepdim=True)
    z = torch.cross(x, y_raw, dim=-1)
    z = z / torch.norm(z, p=2, dim=-1, keepdim=True)
    y = torch.cross(z, x, dim=-1)
    return torch.stack((x, y, z), -1).transpose(-2, -1)


def rotmat_to_rot6d(m):  # geometries.py:117-132
    return m[..., :2, :].clone().reshape(*m.size()[:-2], 6)


def quat_to_rotmat(quat):  # geometries.py:21-41
    nq = quat / (quat.norm(p=2, dim=1, keepdim=True) + 1e-9)
    w, x, y, z = nq[:, 0], nq[:, 1], nq[:, 2], nq[:, 3]
    w2, x2, y2, z2 = w.pow(2), x.pow(2), y.pow(2), z.pow(2)
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).view(-1, 3, 3)


# ----------------------------------------------------------------------------------------------------------
# transforms (lib/utils/transforms.py) and the heatmap integral (lib/utils/integral.py)
# ----------------------------------------------------------------------------------------------------------
def inv_intrinsics(K):  # transforms.py:145-162 / integral.py:56-73: divisions in float64, stored float32
    B = K.shape[0]
    m = torch.zeros(B, 3, 3, dtype=torch.float32)
    fx, fy, cx, cy = (K[:, 0, 0].double(), K[:, 1, 1].double(), K[:, 0, 2].double(), K[:, 1, 2].double())
    m[:, 0, 0] = 1.0 / fx
    m[:, 0, 2] = -cx / fx
    m[:, 1, 1] = 1.0 / fy
    m[:, 1, 2] = -cy / fy
    m[:, 2, 2] = 1
    return m


def uvd_to_xyz(uvd, image_size, inv_k, root_trans, depth_factor):  # transforms.py:33-73
    new = uvd.clone()
    new[:, :, 0] = (uvd[:, :, 0] + 0.5) * image_size
    new[:, :, 1] = (uvd[:, :, 1] + 0.5) * image_size
    new[:, :, 2] = uvd[:, :, 2] * depth_factor
    dz = new[:, :, 2]
    uv_homo = torch.cat((new[:, :, :2], torch.ones_like(new)[:, :, 2:]), dim=2)
    xyz = torch.matmul(inv_k.unsqueeze(1), uv_homo.unsqueeze(-1)).squeeze(3)
    abs_z = dz + root_trans[:, 2].unsqueeze(-1)
    return xyz * abs_z.unsqueeze(-1)


def xyz_to_uvd(xyz, image_size, K, root_trans, depth_factor):  # transforms.py:76-107
    uvz = torch.matmul(K.unsqueeze(1), xyz.unsqueeze(-1)).squeeze(3)
    uv = uvz / uvz[:, :, 2:3]
    out = torch.empty_like(xyz)
    out[:, :, 2] = (xyz[:, :, 2] - root_trans[:, 2:3]) / depth_factor
    out[:, :, 0] = uv[:, :, 0] / float(image_size) - 0.5
    out[:, :, 1] = uv[:, :, 1] / float(image_size) - 0.5
    return out


def uvz2xyz_singlepoint(uv, z, K):  # transforms.py:133-143
    inv_k = inv_intrinsics(K)
    v = torch.cat([uv * z, z], dim=1)
    return torch.matmul(inv_k, v.unsqueeze(-1)).squeeze(-1)


def point_projection_from_3d(K, pts):  # transforms.py:7-21 (numpy and tensor variants are the same maths)
    out = []
    for k, p in zip(K, pts):
        v = torch.matmul(k, p.T)
        out.append((v / v[-1])[:-1].T)
    return torch.stack(out)


def heatmap_integral(out, nkpt, K, root_trans, rootid, fixroot=True, image_size=256.0, depth_factor=1.3,
                     D=64, H=64, W=64, renorm=True):
    """integral.py:97-145 (resnet branch: softmax then /sum renorm; hrnet branch :147-186 has renorm=False)."""
    B = out.shape[0]
    inv_k = inv_intrinsics(K)
    hm = F.softmax(out.reshape(B, nkpt, -1), 2)
    if renorm:
        hm = hm / hm.sum(dim=2, keepdim=True)
    hm = hm.reshape(B, nkpt, D, H, W)
    hm_x0, hm_y0, hm_z0 = hm.sum((2, 3)), hm.sum((2, 4)), hm.sum((3, 4))
    r = torch.arange(W, dtype=torch.float32)
    cx = (hm_x0 * r).sum(dim=2, keepdim=True) / float(W) - 0.5
    cy = (hm_y0 * r).sum(dim=2, keepdim=True) / float(H) - 0.5
    cz = (hm_z0 * r).sum(dim=2, keepdim=True) / float(D) - 0.5
    uvd = torch.cat((cx, cy, cz), dim=2)
    if fixroot:
        uvd[:, rootid, 2] = 0.0
    xyz = uvd_to_xyz(uvd, image_size, inv_k, root_trans, depth_factor)
    return uvd, xyz


# ----------------------------------------------------------------------------------------------------------
# networks, driven by a reference-keyed state_dict
# ----------------------------------------------------------------------------------------------------------
class _Ctx:
    """Functional conv/BN evaluator.  `calib` != None switches BatchNorm to batch statistics and records them
    (the BN-calibration pass of the synthetic-weight recipe)."""

    def __init__(self, sd, prefix="", calib=None, taps=None):
        self.sd, self.prefix, self.calib, self.taps = sd, prefix, calib, taps

    def conv(self, x, name, stride=1, pad=0):
        w = self.sd[self.prefix + name + ".weight"]
        b = self.sd.get(self.prefix + name + ".bias")
        return F.conv2d(x, w, b, stride, pad)

    def bn(self, x, name):
        p = self.prefix + name
        if self.calib is not None:
            mean = x.mean(dim=(0, 2, 3))
            var = x.var(dim=(0, 2, 3), unbiased=False)
            mean, var = self.calib(p, mean, var)
            self.sd[p + ".running_mean"], self.sd[p + ".running_var"] = mean, var
        return F.batch_norm(x, self.sd[p + ".running_mean"], self.sd[p + ".running_var"], self.sd[p + ".weight"],
                            self.sd[p + ".bias"], False, 0.0, 1e-5)

    def tap(self, name, x):
        if self.taps is not None:
            self.taps[self.prefix + name] = x
        return x


def _bottleneck(c, x, name, stride=1):  # Resnet.py:96-135 / HRnet.py:60-98 (stride on the 3x3)
    out = F.relu(c.bn(c.conv(x, name + ".conv1"), name + ".bn1"))
    out = F.relu(c.bn(c.conv(out, name + ".conv2", stride, 1), name + ".bn2"))
    out = c.bn(c.conv(out, name + ".conv3"), name + ".bn3")
    if (c.prefix + name + ".downsample.0.weight") in c.sd:
        x = c.bn(c.conv(x, name + ".downsample.0", stride, 0), name + ".downsample.1")
    return F.relu(out + x)


def _basic(c, x, name):  # HRnet.py:28-57
    out = F.relu(c.bn(c.conv(x, name + ".conv1", 1, 1), name + ".bn1"))
    out = c.bn(c.conv(out, name + ".conv2", 1, 1), name + ".bn2")
    return F.relu(out + x)


def resnet50_forward(sd, x, prefix="", calib=None, taps=None):
    """Resnet.py:56-67."""
    c = _Ctx(sd, prefix, calib, taps)
    x = F.relu(c.bn(c.conv(x, "conv1", 2, 3), "bn1"))
    x = c.tap("stem", x)
    x = F.max_pool2d(x, 3, 2, 1)
    for li, blocks in enumerate((3, 4, 6, 3), start=1):
        for b in range(blocks):
            x = _bottleneck(c, x, f"layer{li}.{b}", stride=2 if (b == 0 and li > 1) else 1)
        x = c.tap(f"layer{li}", x)
    return x


def hrnet32_forward(sd, x, prefix="", calib=None, taps=None):
    """HRnet.py:499-570 with generate_feat=True, generate_hm=False -> (B,2048) pooled feature."""
    c = _Ctx(sd, prefix, calib, taps)
    x = F.relu(c.bn(c.conv(x, "conv1", 2, 1), "bn1"))
    x = F.relu(c.bn(c.conv(x, "conv2", 2, 1), "bn2"))
    for b in range(4):
        x = _bottleneck(c, x, f"layer1.{b}")
    x = c.tap("layer1", x)
    ys = [F.relu(c.bn(c.conv(x, "transition1.0.0", 1, 1), "transition1.0.1")),
          F.relu(c.bn(c.conv(x, "transition1.1.0.0", 2, 1), "transition1.1.0.1"))]
    for stage, nmod in ((2, 1), (3, 4), (4, 3)):
        nb = stage
        if stage > 2:  # HRnet.py:516-521 / :524-529: new branch from the previous stage's LAST output
            t = f"transition{stage - 1}.{nb - 1}.0"
            ys = ys + [F.relu(c.bn(c.conv(ys[-1], t + ".0", 2, 1), t + ".1"))]
        for m in range(nmod):
            mod = f"stage{stage}.{m}"
            xs = []
            for br in range(nb):
                v = ys[br]
                for blk in range(4):
                    v = _basic(c, v, f"{mod}.branches.{br}.{blk}")
                xs.append(v)
            fused = []
            for i in range(nb):  # HRnet.py:254-263
                y = None
                for j in range(nb):
                    f = f"{mod}.fuse_layers.{i}.{j}"
                    if j == i:
                        t_ = xs[j]
                    elif j > i:
                        t_ = c.bn(c.conv(xs[j], f + ".0"), f + ".1")
                        t_ = F.interpolate(t_, scale_factor=2 ** (j - i), mode="nearest")
                    else:
                        t_ = xs[j]
                        for k in range(i - j):
                            t_ = c.bn(c.conv(t_, f"{f}.{k}.0", 2, 1), f"{f}.{k}.1")
                            if k != i - j - 1:
                                t_ = F.relu(t_)
                    y = t_ if y is None else y + t_
                fused.append(F.relu(y))
            ys = fused
        for i, v in enumerate(ys):
            c.tap(f"stage{stage}.out{i}", v)
    # classification head, HRnet.py:557-568
    y = _bottleneck(c, ys[0], "incre_modules.0.0")
    for i in range(3):
        d = F.relu(c.bn(c.conv(y, f"downsamp_modules.{i}.0", 2, 1), f"downsamp_modules.{i}.1"))
        y = _bottleneck(c, ys[i + 1], f"incre_modules.{i + 1}.0") + d
    y = F.relu(c.bn(c.conv(y, "final_feat_layer.0"), "final_feat_layer.1"))
    c.tap("final_feat", y)
    return F.avg_pool2d(y, kernel_size=y.shape[2:]).view(y.size(0), -1)


def depthnet_forward(sd, x, k_value, calib=None):
    """RootNet.forward, depth_net.py:92-137 (hrnet32, no xy / offset / fc branches) -> depth in millimetres."""
    feat = hrnet32_forward(sd, x.float(), "backbone.", calib)
    gamma = F.conv2d(feat[:, :, None, None], sd["depth_layer.weight"], sd["depth_layer.bias"]).view(-1, 1)
    return gamma * k_value.view(-1, 1)


def full_features(sd, x_reg, x_root, calib=None, taps=None):
    """The convolutional part of RootNetwithRegInt.forward (full_net.py:251-296): pooled HRNet feature (B,2048),
    ResNet trunk output (B,2048,8,8) and the heatmap logits (B,nkpt*64,64,64)."""
    x_reg, x_root = x_reg.float(), x_root.float()
    feat = hrnet32_forward(sd, x_root, "rootnet_backbone.", calib, taps)
    x_out = resnet50_forward(sd, x_reg, "reg_backbone.", calib, taps)
    c = _Ctx(sd, "", calib, taps)
    out = x_out
    for i in range(3):
        out = F.conv_transpose2d(out, sd[f"deconv_layers.{3 * i}.weight"], None, stride=2, padding=1)
        out = F.relu(c.bn(out, f"deconv_layers.{3 * i + 1}"))
    c.tap("deconv", out)
    out = F.conv2d(out, sd["final_layer.weight"], sd["final_layer.bias"])
    c.tap("heatmap", out)
    return feat, x_out, out


def full_head(sd, robot: OracleRobot, feat, x_out, heat, k_value, K, n_iter=4, init_pose=None, init_rot=None):
    """Everything after the convolutions (full_net.py:271-287,294,297-397), fp32."""
    dof, nkpt, ref = ROBOTS[robot.robot_type]
    B = feat.shape[0]
    init_pose = sd["init_pose"].expand(B, -1) if init_pose is None else init_pose
    init_rot = sd["init_rot"].expand(B, -1) if init_rot is None else init_rot
    # A. root depth, full_net.py:271-287
    gamma = F.conv2d(feat[:, :, None, None], sd["depth_layer.weight"], sd["depth_layer.bias"]).view(-1, 1)
    pred_depth = (gamma * k_value.view(-1, 1)).reshape(B, 1) / 1000.0
    root_trans = torch.zeros(B, 3)
    root_trans[:, 2:3] = pred_depth
    # B. keypoints, full_net.py:294-298
    xf = F.avg_pool2d(x_out, 8, stride=1)
    pred_uvd, pred_xyz_int = heatmap_integral(heat, nkpt, K, root_trans, ref)
    pred_root_uv = (pred_uvd[:, ref, :2] + 0.5) * 256.0
    # C. root translation, full_net.py:305
    pred_trans = uvz2xyz_singlepoint(pred_root_uv, pred_depth, K)
    # D. iterative regressors, full_net.py:308-331,365-378 (dropout is the identity in eval)
    xf = xf.view(B, -1)
    pose, rot = init_pose, init_rot
    for _ in range(n_iter):
        xc = torch.cat([xf, pose], 1)
        xc = F.linear(xc, sd["fc_pose_1.weight"], sd["fc_pose_1.bias"])
        xc = F.linear(xc, sd["fc_pose_2.weight"], sd["fc_pose_2.bias"])
        pose = F.linear(xc, sd["decpose.weight"], sd["decpose.bias"]) + pose
    for _ in range(n_iter):
        xc = torch.cat([xf, rot], 1)
        xc = F.linear(xc, sd["fc_rot_1.weight"], sd["fc_rot_1.bias"])
        xc = F.linear(xc, sd["fc_rot_2.weight"], sd["fc_rot_2.bias"])
        rot = F.linear(xc, sd["decrot.weight"], sd["decrot.bias"]) + rot
    # E. FK, full_net.py:380-383
    if ref == 0:
        xyz_fk = robot.get_keypoints(pose, rot, pred_trans)
    else:
        xyz_fk = robot.get_keypoints_root(pose, rot, pred_trans, root=ref)
    return pose, rot, pred_trans, pred_root_uv, pred_depth, pred_uvd, pred_xyz_int, xyz_fk


def full_forward(sd, robot: OracleRobot, x_reg, x_root, k_value, K, n_iter=4, calib=None, taps=None,
                 init_pose=None, init_rot=None):
    """RootNetwithRegInt.forward, full_net.py:239-397 (resnet50 + hrnet32, rotation_dim 6, fix_root True)."""
    feat, x_out, heat = full_features(sd, x_reg, x_root, calib, taps)
    return full_head(sd, robot, feat, x_out, heat, k_value, K, n_iter, init_pose, init_rot)
