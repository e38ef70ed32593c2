"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product package.

MuJoCo is absent from the image, yet the reference derives every kernel constant from a
compiled `MjModel` / `MjData` (`transformations/poses.py:14-23`, `core/simulate.py:98-156`).
This file restates, for exactly the MJCF subset the reference's models use, what MuJoCo
3.3.0 (pinned in the reference's `uv.lock:1131-1132`) and dm_control's `attach`
(`core/core.py:310-311`) would produce:

  * compiler defaults: angles in degrees, `eulerseq="xyz"` (intrinsic x-y-z => R = Rx Ry Rz)
  * `body` pos / euler / quat; one `joint` per body (type slide|hinge, axis, pos);
    `inertial` pos / quat / mass / diaginertia; `site` pos / euler / quat; `keyframe` qpos
  * attaching the target model at site "attachment" (`sequential.xml:33`): a new body
    "target/" (body 7, no mass, pose = the site's pose) holding body "target/object"
    (body 8, explicit inertial from the CAD CSV, `core/core.py:245-253,286`) and the site
    "target/ft_sensor" with euler "0 0 180" (`core/core.py:241-242`)  -- body numbering as in
    the reference's `README.md:14-26`
  * `mj_kinematics` for slide / hinge chains (xpos, xmat, xipos, ximat, site_xpos, site_xmat)

Only names the reference reads are produced.  Mesh geoms are ignored: the object body
carries an explicit <inertial>, which MuJoCo prefers over geom-derived inertia, and nothing
in the scene collides.
"""
from __future__ import annotations

import csv
import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

SLIDE, HINGE = 2, 3  # mjtJoint values tested at reference core/simulate.py:101,103


# ------------------------------------------------------------------ quaternion helpers (wxyz)
def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ]
    )


def quat_to_mat(q):
    w, x, y, z = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (w * y + x * z)],
            [2 * (w * z + x * y), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (w * x + y * z), 1 - 2 * (x * x + y * y)],
        ]
    )


def euler_deg_to_quat(euler, seq="xyz"):
    """MuJoCo euler -> quaternion: lower-case letters rotate about the moving frame (post-multiply)."""
    q = np.array([1.0, 0.0, 0.0, 0.0])
    for ang, ax in zip(euler, seq):
        half = 0.5 * math.radians(ang)
        r = np.array([math.cos(half), 0.0, 0.0, 0.0])
        r["xyz".index(ax.lower()) + 1] = math.sin(half)
        q = quat_mul(q, r) if ax.islower() else quat_mul(r, q)
    return q / np.linalg.norm(q)


def axis_angle_quat(axis, angle):
    axis = np.asarray(axis, float)
    n = np.linalg.norm(axis)
    half = 0.5 * angle
    return np.concatenate([[math.cos(half)], math.sin(half) * axis / n])


# ------------------------------------------------------------------ transforms3d restatements
def euler2mat_sxyz(ai, aj, ak):
    """transforms3d.euler.euler2mat(ai, aj, ak, 'sxyz') (pinned 0.4.2): static x, then y, then z
    => R = Rz(ak) Ry(aj) Rx(ai), radians.  Used at reference core/core.py:144."""
    ci, si = math.cos(ai), math.sin(ai)
    cj, sj = math.cos(aj), math.sin(aj)
    ck, sk = math.cos(ak), math.sin(ak)
    Rx = np.array([[1, 0, 0], [0, ci, -si], [0, si, ci]])
    Ry = np.array([[cj, 0, sj], [0, 1, 0], [-sj, 0, cj]])
    Rz = np.array([[ck, -sk, 0], [sk, ck, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def mat2quat(M):
    """transforms3d.quaternions.mat2quat: dominant eigenvector of Bar-Itzhack's K matrix, w >= 0.
    Used at reference core/core.py:145."""
    Qxx, Qyx, Qzx, Qxy, Qyy, Qzy, Qxz, Qyz, Qzz = np.asarray(M, float).flat
    K = (
        np.array(
            [
                [Qxx - Qyy - Qzz, 0, 0, 0],
                [Qyx + Qxy, Qyy - Qxx - Qzz, 0, 0],
                [Qzx + Qxz, Qzy + Qyz, Qzz - Qxx - Qyy, 0],
                [Qyz - Qzy, Qzx - Qxz, Qxy - Qyx, Qxx + Qyy + Qzz],
            ]
        )
        / 3.0
    )
    vals, vecs = np.linalg.eigh(K)
    q = vecs[[3, 0, 1, 2], np.argmax(vals)]
    if q[0] < 0:
        q = -q
    return q


# ------------------------------------------------------------------ CAD ground truth (core/core.py:133-192)
def read_cad_row(csv_path):
    """First data row of `object_cad_gt.csv` as {column: float} (reference core/core.py:135-138)."""
    with open(csv_path, newline="") as f:
        rd = csv.reader(f)
        header = next(rd)
        row = next(rd)
    return {k: (float(v) if k != "id" else v) for k, v in zip(header, row)}


def target_ground_truth(cad):
    """reference core/core.py:140-192 (get_target_object_ground_truth) from a parsed CSV row."""
    rot_obji_obj = euler2mat_sxyz(cad["rx"], cad["ry"], cad["rz"])  # :143-144
    iquat = mat2quat(rot_obji_obj.T)  # :145
    mass = cad["total_mass"]  # :148
    com = np.array([cad["cx"], cad["cy"], cad["cz"]])  # :149
    ixx, iyy, izz, ixy, iyz, izx = (cad[k] for k in ("ixx", "iyy", "izz", "ixy", "iyz", "izx"))  # :153
    full = np.array([[ixx, ixy, izx], [ixy, iyy, iyz], [izx, iyz, izz]])  # :155-161
    diag = np.diag(rot_obji_obj @ full @ rot_obji_obj.T).copy()  # :164-165 (off-diagonals dropped)
    # :168-173 coordinate_transfer_imat with R = I, t = com  (dynamics/dynamics.py:252-257)
    t = com.reshape(3, 1)
    glob = full + mass * (float((t.T @ t)[0, 0]) * np.eye(3) - t @ t.T)
    globalinertia = [glob[0, 0], glob[1, 1], glob[2, 2], glob[0, 1], glob[1, 2], glob[2, 0]]  # :175-182
    return dict(
        aabb_scale=cad["aabb_scale"],
        mass=mass,
        com=com,
        iquat=iquat,
        diaginertia=diag,
        fullinertia=[ixx, iyy, izz, ixy, iyz, izx],
        globalinertia=globalinertia,
    )


# ------------------------------------------------------------------ the compiled-model stand-in
@dataclass
class MjModelLike:
    body_names: list = field(default_factory=list)
    body_parentid: list = field(default_factory=list)
    body_pos: np.ndarray = None
    body_quat: np.ndarray = None
    body_ipos: np.ndarray = None
    body_iquat: np.ndarray = None
    body_mass: np.ndarray = None
    body_inertia: np.ndarray = None
    jnt_type: np.ndarray = None
    jnt_axis: np.ndarray = None
    jnt_pos: np.ndarray = None
    jnt_bodyid: np.ndarray = None
    site_names: list = field(default_factory=list)
    site_bodyid: list = field(default_factory=list)
    site_pos: np.ndarray = None
    site_quat: np.ndarray = None
    key_qpos: np.ndarray = None
    gravity: np.ndarray = None
    timestep: float = 0.002

    @property
    def njnt(self):
        return len(self.jnt_type)

    @property
    def nu(self):
        return len(self.jnt_type)

    def body_id(self, name):
        return self.body_names.index(name)

    def site_id(self, name):
        return self.site_names.index(name)


@dataclass
class MjDataLike:
    qpos: np.ndarray
    xpos: np.ndarray
    xmat: np.ndarray  # (nbody, 9) row-major, like MuJoCo
    xipos: np.ndarray
    ximat: np.ndarray
    site_xpos: np.ndarray
    site_xmat: np.ndarray


def _vec(s, n, default):
    if s is None:
        return np.array(default, float)
    v = np.array([float(x) for x in s.split()], float)
    assert len(v) == n
    return v


def _frame_quat(elem):
    if elem.get("quat") is not None:
        q = _vec(elem.get("quat"), 4, None)
        return q / np.linalg.norm(q)
    if elem.get("euler") is not None:
        return euler_deg_to_quat(_vec(elem.get("euler"), 3, None))
    return np.array([1.0, 0.0, 0.0, 0.0])


def compile_manipulator_with_target(manipulator_xml, ground_truth, attachment_site="attachment"):
    """Compile `sequential.xml` + the attached target into MjModel-like arrays (see module docstring)."""
    root = ET.parse(manipulator_xml).getroot()
    m = MjModelLike()
    bodies = [dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), ipos=np.zeros(3),
                   iquat=np.array([1.0, 0, 0, 0]), mass=0.0, inertia=np.zeros(3))]
    joints, sites = [], []

    def walk(elem, parent_id):
        for b in elem.findall("body"):
            bid = len(bodies)
            rec = dict(name=b.get("name"), parent=parent_id, pos=_vec(b.get("pos"), 3, [0, 0, 0]), quat=_frame_quat(b),
                       ipos=np.zeros(3), iquat=np.array([1.0, 0, 0, 0]), mass=0.0, inertia=np.zeros(3))
            inert = b.find("inertial")
            if inert is not None:
                rec["ipos"] = _vec(inert.get("pos"), 3, [0, 0, 0])
                rec["iquat"] = _frame_quat(inert)
                rec["mass"] = float(inert.get("mass"))
                rec["inertia"] = _vec(inert.get("diaginertia"), 3, None)
            bodies.append(rec)
            for j in b.findall("joint"):
                jt = {"slide": SLIDE, "hinge": HINGE}[j.get("type", "hinge")]
                ax = _vec(j.get("axis"), 3, [0, 0, 1])
                joints.append(dict(type=jt, axis=ax / np.linalg.norm(ax), pos=_vec(j.get("pos"), 3, [0, 0, 0]), body=bid))
            for s in b.findall("site"):
                sites.append(dict(name=s.get("name"), body=bid, pos=_vec(s.get("pos"), 3, [0, 0, 0]), quat=_frame_quat(s)))
            walk(b, bid)

    walk(root.find("worldbody"), 0)

    # dm_control attach at the site: body "target/" with the site's pose, then "target/object"
    att = next(s for s in sites if s["name"] == attachment_site)
    frame_id = len(bodies)
    bodies.append(dict(name="target/", parent=att["body"], pos=att["pos"].copy(), quat=att["quat"].copy(),
                       ipos=np.zeros(3), iquat=np.array([1.0, 0, 0, 0]), mass=0.0, inertia=np.zeros(3)))
    iq = np.asarray(ground_truth["iquat"], float)
    bodies.append(dict(name="target/object", parent=frame_id, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]),
                       ipos=np.asarray(ground_truth["com"], float), iquat=iq / np.linalg.norm(iq),
                       mass=float(ground_truth["mass"]), inertia=np.asarray(ground_truth["diaginertia"], float)))
    sites.append(dict(name="target/ft_sensor", body=frame_id, pos=np.zeros(3), quat=euler_deg_to_quat([0, 0, 180])))

    m.body_names = [b["name"] for b in bodies]
    m.body_parentid = [b["parent"] for b in bodies]
    for k in ("pos", "quat", "ipos", "iquat", "inertia"):
        setattr(m, "body_" + k, np.array([b[k] for b in bodies], float))
    m.body_mass = np.array([b["mass"] for b in bodies], float)
    m.jnt_type = np.array([j["type"] for j in joints])
    m.jnt_axis = np.array([j["axis"] for j in joints], float)
    m.jnt_pos = np.array([j["pos"] for j in joints], float)
    m.jnt_bodyid = np.array([j["body"] for j in joints])
    m.site_names = [s["name"] for s in sites]
    m.site_bodyid = [s["body"] for s in sites]
    m.site_pos = np.array([s["pos"] for s in sites], float)
    m.site_quat = np.array([s["quat"] for s in sites], float)
    key = root.find("keyframe/key")
    m.key_qpos = _vec(key.get("qpos"), len(joints), None) if key is not None else np.zeros(len(joints))
    m.gravity = np.array([0.0, 0.0, -9.81])  # MjOption default, read at core/simulate.py:149
    return m


def kinematics(m: MjModelLike, qpos) -> MjDataLike:
    """mj_kinematics for slide/hinge trees with at most one joint per body."""
    nb = len(m.body_names)
    xpos = np.zeros((nb, 3))
    xquat = np.zeros((nb, 4))
    xquat[0] = [1, 0, 0, 0]
    jnt_of_body = {int(b): j for j, b in enumerate(m.jnt_bodyid)}
    for b in range(1, nb):
        p = m.body_parentid[b]
        Rp = quat_to_mat(xquat[p])
        pos = xpos[p] + Rp @ m.body_pos[b]
        quat = quat_mul(xquat[p], m.body_quat[b])
        if b in jnt_of_body:
            j = jnt_of_body[b]
            R = quat_to_mat(quat)
            axis_w = R @ m.jnt_axis[j]
            if m.jnt_type[j] == SLIDE:
                pos = pos + axis_w * qpos[j]
            else:
                anchor = pos + R @ m.jnt_pos[j]
                quat = quat_mul(axis_angle_quat(axis_w, qpos[j]), quat)
                pos = anchor - quat_to_mat(quat) @ m.jnt_pos[j]
        xpos[b] = pos
        xquat[b] = quat / np.linalg.norm(quat)
    xmat = np.array([quat_to_mat(q) for q in xquat])
    xipos = np.array([xpos[b] + xmat[b] @ m.body_ipos[b] for b in range(nb)])
    ximat = np.array([xmat[b] @ quat_to_mat(m.body_iquat[b]) for b in range(nb)])
    sx = np.array([xpos[b] + xmat[b] @ p for b, p in zip(m.site_bodyid, m.site_pos)])
    sm = np.array([xmat[b] @ quat_to_mat(q) for b, q in zip(m.site_bodyid, m.site_quat)])
    return MjDataLike(
        qpos=np.asarray(qpos, float),
        xpos=xpos,
        xmat=xmat.reshape(nb, 9),
        xipos=xipos,
        ximat=ximat.reshape(nb, 9),
        site_xpos=sx,
        site_xmat=sm.reshape(len(m.site_names), 9),
    )
