"""MuJoCo-free model front-end: robot + target description -> the constants the kernels evaluate.

The reference derives these from a compiled MuJoCo model (reference transformations/poses.py:14-23 and the setup block
of core/simulate.py:98-156); MuJoCo, dm_control and the reference checkout are not available where the kernels run, so
the same constants are built here from plain numbers:

  * a robot description (JSON extracted from xml_models/manipulators/sequential.xml by
    tools/extract_reference_assets.py, or parsed straight from an MJCF file with `load_mjcf`)
  * the first data row of a target's `object_cad_gt.csv` (assets/targets.json, or `load_cad_csv`)

Recipe (all in homogeneous 4x4 / spatial 6x6 algebra, host numpy: a few dozen tiny matrices, once per model):
  hposes[k]  = (T_{k-1,(k-1)j}^-1  T_{k-1,k}  T_{k,kj})^-1                          core/simulate.py:140-146
  simats[k]  = inertia of link k about its joint frame                                core/simulate.py:115-123
  simats[6] += attachment frame body (massless) + object, moved to link 6's joint frame   core/simulate.py:129-137
  object inertial = CAD row -> (mass, CoM, principal frame R_e^T, diag(R_e I R_e^T))      core/core.py:143-165,245-253
  uscrews    = joint axis in the linear (slide) or angular (hinge) half                     core/simulate.py:98-110
  dtwist_0   = -[gravity, 0, 0, 0]                                                         core/simulate.py:149
  pose_sen_llj = static F/T sensor pose w.r.t. link 6's joint frame                         core/simulate.py:202
"""
from __future__ import annotations

import csv
import json
import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


# ----------------------------------------------------------------------------------------------------------------
# small rigid-transform helpers (4x4 homogeneous)
# ----------------------------------------------------------------------------------------------------------------
def _rot_axis(axis: int, angle: float) -> np.ndarray:
    c, s = math.cos(angle), math.sin(angle)
    i, j = (axis + 1) % 3, (axis + 2) % 3
    R = np.eye(3)
    R[i, i], R[i, j], R[j, i], R[j, j] = c, -s, s, c
    return R


def euler_to_R(euler, seq: str = "xyz", degrees: bool = True) -> np.ndarray:
    """MuJoCo `euler` attribute: lower-case axes rotate about the moving frame (R = R_a R_b R_c), upper-case about the fixed one."""
    R = np.eye(3)
    for ang, ax in zip(euler, seq):
        Ri = _rot_axis("xyz".index(ax.lower()), math.radians(ang) if degrees else ang)
        R = R @ Ri if ax.islower() else Ri @ R
    return R


def quat_to_R(q) -> np.ndarray:
    w, x, y, z = np.asarray(q, float) / np.linalg.norm(q)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (w * y + x * z)],
        [2 * (w * z + x * y), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (w * x + y * z), 1 - 2 * (x * x + y * y)],
    ])


def make_T(R=None, t=None) -> np.ndarray:
    T = np.eye(4)
    if R is not None:
        T[:3, :3] = R
    if t is not None:
        T[:3, 3] = t
    return T


def inv_T(T) -> np.ndarray:
    R, t = T[:3, :3], T[:3, 3]
    return make_T(R.T, -R.T @ t)


def hat(v) -> np.ndarray:
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


def adjoint_of(T) -> np.ndarray:
    R, t = T[:3, :3], T[:3, 3]
    Ad = np.zeros((6, 6))
    Ad[:3, :3] = Ad[3:, 3:] = R
    Ad[:3, 3:] = hat(t) @ R
    return Ad


def move_inertia(T_ab, G_b) -> np.ndarray:
    """Spatial inertia given in {b}, expressed in {a}, with T_ab the pose of {b} in {a}:  Ad(T_ba)^T G_b Ad(T_ba)
    (Modern Robotics Eq. 8.42; the call-site semantics of reference dynamics.transfer_simat, dynamics.py:102-104)."""
    Ad = adjoint_of(inv_T(T_ab))
    return Ad.T @ G_b @ Ad


def T_to_Rt(T) -> np.ndarray:
    return np.concatenate([T[:3, :3].reshape(9), T[:3, 3]])


# ----------------------------------------------------------------------------------------------------------------
# descriptions
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class Link:
    name: str
    pos: np.ndarray            # body frame in the parent body frame
    R: np.ndarray
    joint_type: str            # "slide" | "hinge"
    joint_axis: np.ndarray
    joint_pos: np.ndarray
    mass: float
    ipos: np.ndarray           # inertial frame in the body frame
    iR: np.ndarray
    diaginertia: np.ndarray


@dataclass
class Robot:
    links: list
    attachment_T: np.ndarray   # attachment site in the last link's body frame
    sensor_T_in_attachment: np.ndarray
    key_qpos: np.ndarray
    gravity: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -9.81]))
    timestep: float = 0.002


@dataclass
class Target:
    name: str
    aabb_scale: float
    mass: float
    com: np.ndarray
    inertia_com: np.ndarray    # 3x3 about the CoM, body (AABB) axes
    euler_sxyz: np.ndarray     # rx, ry, rz (radians, static x-y-z)

    # -- the reference's CAD -> MuJoCo inertial recipe, core/core.py:143-165 -------------------------------
    @property
    def R_principal_from_body(self) -> np.ndarray:
        """rot_obji_obj = euler2mat(rx, ry, rz, 'sxyz') = Rz(rz) Ry(ry) Rx(rx)   (core/core.py:144)."""
        rx, ry, rz = self.euler_sxyz
        return _rot_axis(2, rz) @ _rot_axis(1, ry) @ _rot_axis(0, rx)

    @property
    def diaginertia(self) -> np.ndarray:
        """diag(R_e I R_e^T): the off-diagonal residue is DROPPED by the reference (core/core.py:164-165)."""
        Re = self.R_principal_from_body
        return np.diag(Re @ self.inertia_com @ Re.T).copy()

    @property
    def global_inertia(self) -> np.ndarray:
        """Inertia about the body-frame origin (parallel axis, core/core.py:168-173); order ixx iyy izz ixy iyz izx."""
        c = self.com
        I0 = self.inertia_com + self.mass * (c @ c * np.eye(3) - np.outer(c, c))
        return np.array([I0[0, 0], I0[1, 1], I0[2, 2], I0[0, 1], I0[1, 2], I0[2, 0]])

    @property
    def ground_truth_params(self) -> np.ndarray:
        """[m, m c, globalinertia] -- what reference main.py:79-82 scores the identification against."""
        return np.concatenate([[self.mass], self.mass * self.com, self.global_inertia])


def robot_from_dict(d: dict) -> Robot:
    seq = d.get("eulerseq", "xyz")
    deg = d.get("angle", "degree") == "degree"
    links = []
    for rec in d["links"]:
        j, it = rec["joint"], rec["inertial"]
        ax = np.array(j["axis"], float)
        links.append(Link(
            name=rec["name"], pos=np.array(rec["pos"], float), R=euler_to_R(rec["euler_deg"], seq, deg),
            joint_type=j["type"], joint_axis=ax / np.linalg.norm(ax), joint_pos=np.array(j["pos"], float),
            mass=float(it["mass"]), ipos=np.array(it["pos"], float), iR=np.eye(3), diaginertia=np.array(it["diaginertia"], float)))
    site = next(s for s in d["sites"] if s["name"] == d["ft_sensor_site"]["parent_site"])
    if site["body"] != links[-1].name:
        raise ValueError("the attachment site must sit on the last link")
    att = make_T(euler_to_R(site["euler_deg"], seq, deg), np.array(site["pos"], float))
    sen = make_T(euler_to_R(d["ft_sensor_site"]["euler_deg"], seq, deg))
    return Robot(links=links, attachment_T=att, sensor_T_in_attachment=sen, key_qpos=np.array(d["keyframe"]["qpos"], float),
                 gravity=np.array(d.get("gravity", [0, 0, -9.81]), float), timestep=float(d.get("timestep", 0.002)))


def target_from_row(name: str, row: dict) -> Target:
    I = np.array([[row["ixx"], row["ixy"], row["izx"]], [row["ixy"], row["iyy"], row["iyz"]], [row["izx"], row["iyz"], row["izz"]]], float)
    return Target(name=name, aabb_scale=float(row["aabb_scale"]), mass=float(row["total_mass"]),
                  com=np.array([row["cx"], row["cy"], row["cz"]], float), inertia_com=I,
                  euler_sxyz=np.array([row["rx"], row["ry"], row["rz"]], float))


def load_cad_csv(path: str, name: str | None = None) -> Target:
    """First data row of an `object_cad_gt.csv` (reference core/core.py:135-138)."""
    with open(path, newline="") as f:
        rd = csv.reader(f)
        header, row = next(rd), next(rd)
    rec = {k: (v if k == "id" else float(v)) for k, v in zip(header, row)}
    return target_from_row(name or os.path.basename(os.path.dirname(path)), rec)


def load_mjcf(path: str) -> Robot:
    """Parse the MJCF subset of the reference's manipulators (serial chain, one joint per body, explicit inertials)."""
    root = ET.parse(path).getroot()

    def fl(s, default):
        return [float(x) for x in s.split()] if s is not None else list(default)

    links, sites = [], []

    def walk(elem):
        for b in elem.findall("body"):
            j, it = b.find("joint"), b.find("inertial")
            if j is None or it is None:
                raise ValueError(f"body {b.get('name')}: the supported subset needs one <joint> and an explicit <inertial>")
            if it.get("diaginertia") is None or it.get("quat") is not None or it.get("euler") is not None:
                raise ValueError(f"body {b.get('name')}: only axis-aligned `diaginertia` inertials are supported")
            if b.get("quat") is not None:
                raise ValueError("body `quat` is not supported by this subset parser (use euler)")
            links.append({"name": b.get("name"), "pos": fl(b.get("pos"), [0, 0, 0]), "euler_deg": fl(b.get("euler"), [0, 0, 0]),
                          "joint": {"type": j.get("type", "hinge"), "axis": fl(j.get("axis"), [0, 0, 1]), "pos": fl(j.get("pos"), [0, 0, 0])},
                          "inertial": {"pos": fl(it.get("pos"), [0, 0, 0]), "mass": float(it.get("mass")), "diaginertia": fl(it.get("diaginertia"), [])}})
            for s in b.findall("site"):
                sites.append({"name": s.get("name"), "body": b.get("name"), "pos": fl(s.get("pos"), [0, 0, 0]), "euler_deg": fl(s.get("euler"), [0, 0, 0])})
            walk(b)

    walk(root.find("worldbody"))
    key = root.find("keyframe/key")
    d = {"links": links, "sites": sites, "keyframe": {"qpos": fl(key.get("qpos"), []) if key is not None else [0.0] * len(links)},
         "ft_sensor_site": {"parent_site": "attachment", "euler_deg": [0.0, 0.0, 180.0]}}
    return robot_from_dict(d)


# ----------------------------------------------------------------------------------------------------------------
# constants
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class Constants:
    hposes_Rt: np.ndarray      # (nj+1, 12)
    simats: np.ndarray         # (nj+1, 6, 6)
    uscrews: np.ndarray        # (nj, 6)
    twist_0: np.ndarray
    dtwist_0: np.ndarray
    pose_sen_Rt: np.ndarray    # (12,) pose_sen_llj
    pose_sen_obj_Rt: np.ndarray
    pose_sen_obji_Rt: np.ndarray
    simat_object_llj: np.ndarray   # attachment + object inertia in link 6's joint frame ("simat_sen_obj" in the reference)
    key_qpos: np.ndarray
    timestep: float
    target: Target | None = None

    def hposes(self):
        from .lie import se3_from_Rt

        return [se3_from_Rt(r) for r in self.hposes_Rt]


def _diag_simat(mass, diag):
    return np.diag([mass, mass, mass, diag[0], diag[1], diag[2]]).astype(float)


def build_constants(robot: Robot, target: Target | None) -> Constants:
    n = len(robot.links)
    hposes = [np.eye(4)]
    simats = np.zeros((n + 1, 6, 6))
    uscrews = np.zeros((n, 6))
    T_prev_joint = np.eye(4)  # T_{k-1,(k-1)j}: the world has no joint -> identity (reference poses.py:20)
    for k, L in enumerate(robot.links, start=1):
        T_parent_body = make_T(L.R, L.pos)
        T_body_joint = make_T(None, L.joint_pos)
        hposes.append(inv_T(inv_T(T_prev_joint) @ T_parent_body @ T_body_joint))
        T_joint_inertial = inv_T(T_body_joint) @ make_T(L.iR, L.ipos)
        simats[k] = move_inertia(T_joint_inertial, _diag_simat(L.mass, L.diaginertia))
        if L.joint_type == "slide":
            uscrews[k - 1, :3] = L.joint_axis
        elif L.joint_type == "hinge":
            uscrews[k - 1, 3:] = L.joint_axis
        else:
            raise TypeError("Only slide or hinge joints are supported.")
        T_prev_joint = T_body_joint
    last = robot.links[-1]
    T_ll_llj = make_T(None, last.joint_pos)
    T_llj_att = inv_T(T_ll_llj) @ robot.attachment_T       # attachment body "target/" == object body frame (pos 0, no rotation)
    T_llj_sen = T_llj_att @ robot.sensor_T_in_attachment
    sim_obj = np.zeros((6, 6))
    T_sen_obj = inv_T(robot.sensor_T_in_attachment)
    T_sen_obji = T_sen_obj.copy()
    if target is not None:
        T_obj_obji = make_T(target.R_principal_from_body.T, target.com)   # iquat = mat2quat(R_e^T), core/core.py:145
        sim_obj = move_inertia(T_llj_att @ T_obj_obji, _diag_simat(target.mass, target.diaginertia))
        simats[n] = simats[n] + sim_obj
        T_sen_obji = T_sen_obj @ T_obj_obji
    return Constants(
        hposes_Rt=np.stack([T_to_Rt(T) for T in hposes]), simats=simats, uscrews=uscrews, twist_0=np.zeros(6),
        dtwist_0=-np.concatenate([robot.gravity, np.zeros(3)]), pose_sen_Rt=T_to_Rt(inv_T(T_llj_sen)),
        pose_sen_obj_Rt=T_to_Rt(T_sen_obj), pose_sen_obji_Rt=T_to_Rt(T_sen_obji), simat_object_llj=sim_obj,
        key_qpos=robot.key_qpos.copy(), timestep=robot.timestep, target=target)


# ----------------------------------------------------------------------------------------------------------------
# packaged assets
# ----------------------------------------------------------------------------------------------------------------
def packaged_robot(name: str = "sequential") -> Robot:
    with open(os.path.join(_ASSETS, f"{name}.json")) as f:
        return robot_from_dict(json.load(f))


def packaged_targets() -> dict:
    with open(os.path.join(_ASSETS, "targets.json")) as f:
        return json.load(f)


def packaged_target(name: str) -> Target:
    rows = packaged_targets()
    if name not in rows:
        raise ValueError(f"unknown target '{name}'; available: {', '.join(sorted(rows))}")
    return target_from_row(name, rows[name])


def load_packaged(robot: str = "sequential", target: str | None = "hammer") -> Constants:
    return build_constants(packaged_robot(robot), packaged_target(target) if target else None)
