"""URDF -> flat kinematic table for the CUDA forward-kinematics kernels.

Host-side mirror of the subset of the reference's vendored `urdfpytorch` that the hot path uses
(lib/utils/urdfpytorch/urdf.py:2399-2409 joint XML, :2746-2772 link graph, :3788-3813 actuated-joint order,
utils.py:22-51,142-167 origin parsing).  No meshes, visuals, inertials or XML export -- those are out of scope
(SURVEY.md section 2.1 row 8).  Arithmetic on the table happens only in libhrp_b200.so.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

JOINT_FIXED, JOINT_REVOLUTE, JOINT_PRISMATIC = 0, 1, 2


def _rpy_matrix(rpy) -> np.ndarray:
    r, p, y = (float(v) for v in rpy)
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - cr * sy, sy * sr + cy * cr * sp],
                     [cp * sy, cy * cr + sy * sp * sr, cr * sy * sp - cy * sr],
                     [-sp, cp * sr, cp * cr]], dtype=np.float64)


@dataclass
class Joint:
    name: str
    joint_type: str
    parent: str
    child: str
    origin: np.ndarray
    axis: np.ndarray
    mimic: tuple | None = None


@dataclass
class KinematicTree:
    """Links in parent-before-child order with, per link, the joint that attaches it to its parent."""
    link_names: list = field(default_factory=list)
    parent: list = field(default_factory=list)      # parent link index, -1 for the base
    jtype: list = field(default_factory=list)       # JOINT_*
    origin: list = field(default_factory=list)      # 4x4 float64
    axis: list = field(default_factory=list)        # unit 3-vector float64
    qcol: list = field(default_factory=list)        # column of q driving the joint, -1 if none
    qmul: list = field(default_factory=list)        # cfg = qmul*q[qcol] + qoff (mimic joints)
    qoff: list = field(default_factory=list)
    actuated_joint_names: list = field(default_factory=list)
    joints: dict = field(default_factory=dict)

    @property
    def n_links(self) -> int:
        return len(self.link_names)

    def link_index(self, name: str) -> int:
        return self.link_names.index(name)


def load_urdf(path) -> KinematicTree:
    root = ET.parse(str(path)).getroot()
    links = [l.attrib["name"] for l in root.findall("link")]
    joints = []
    for j in root.findall("joint"):
        origin = np.eye(4, dtype=np.float64)
        o = j.find("origin")
        if o is not None:
            if "xyz" in o.attrib:
                origin[:3, 3] = [float(v) for v in o.attrib["xyz"].split()]
            if "rpy" in o.attrib:
                origin[:3, :3] = _rpy_matrix(o.attrib["rpy"].split())
        a = j.find("axis")
        axis = np.array([float(v) for v in a.attrib["xyz"].split()], dtype=np.float64) if a is not None \
            else np.array([1.0, 0.0, 0.0])
        axis = axis / np.linalg.norm(axis)
        m = j.find("mimic")
        mimic = None if m is None else (m.attrib["joint"], float(m.attrib.get("multiplier", 1.0)),
                                        float(m.attrib.get("offset", 0.0)))
        jt = j.attrib["type"]
        if jt not in ("fixed", "revolute", "continuous", "prismatic"):
            raise NotImplementedError(f"joint type {jt!r} ({j.attrib['name']}) is not supported")
        joints.append(Joint(j.attrib["name"], jt, j.find("parent").attrib["link"], j.find("child").attrib["link"],
                            origin, axis, mimic))
    by_child = {}
    for j in joints:
        if j.child in by_child:
            raise ValueError(f"link {j.child} has two parent joints")
        if j.parent not in links or j.child not in links or j.parent == j.child:
            raise ValueError(f"joint {j.name} has an invalid parent/child")
        by_child[j.child] = j
    bases = [l for l in links if l not in by_child]
    if len(bases) != 1:
        raise ValueError(f"URDF must have exactly one base link, found {bases}")

    def depth(link):
        d = 1
        while link in by_child:
            link = by_child[link].parent
            d += 1
            if d > len(links) + 1:
                raise ValueError("URDF link graph has a cycle")
        return d

    actuated = [j for j in joints if j.mimic is None and j.joint_type != "fixed"]
    order = np.argsort([depth(j.child) for j in actuated], kind="stable")
    actuated = [actuated[i] for i in order]
    qcol = {j.name: i for i, j in enumerate(actuated)}

    tree = KinematicTree()
    tree.joints = {j.name: j for j in joints}
    tree.actuated_joint_names = [j.name for j in actuated]
    ordered = sorted(links, key=depth)
    index = {n: i for i, n in enumerate(ordered)}
    for n in ordered:
        tree.link_names.append(n)
        j = by_child.get(n)
        if j is None:
            tree.parent.append(-1)
            tree.jtype.append(JOINT_FIXED)
            tree.origin.append(np.eye(4))
            tree.axis.append(np.array([1.0, 0.0, 0.0]))
            tree.qcol.append(-1)
            tree.qmul.append(1.0)
            tree.qoff.append(0.0)
            continue
        tree.parent.append(index[j.parent])
        col, mul, off = -1, 1.0, 0.0
        if j.mimic is not None:
            if j.mimic[0] not in qcol:
                raise ValueError(f"joint {j.name} mimics unknown/unactuated joint {j.mimic[0]}")
            col, mul, off = qcol[j.mimic[0]], j.mimic[1], j.mimic[2]
        elif j.name in qcol:
            col = qcol[j.name]
        if j.joint_type == "fixed" or col < 0:
            jt = JOINT_FIXED
        elif j.joint_type in ("revolute", "continuous"):
            jt = JOINT_REVOLUTE
        else:
            jt = JOINT_PRISMATIC
        tree.jtype.append(jt)
        tree.origin.append(j.origin)
        tree.axis.append(j.axis)
        tree.qcol.append(col)
        tree.qmul.append(mul)
        tree.qoff.append(off)
    return tree
