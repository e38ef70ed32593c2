"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Restates, line by line, how the reference turns a compiled MuJoCo model into the
constants its inverse dynamics is evaluated with:

  * the pose registers of `transformations/poses.py:14-23`
  * the setup block of `core/simulate.py:74-156` (unit screws, per-link spatial inertias in
    joint frames, the attachment + object inertia folded into link 6, joint home poses,
    gravity twist) and the static sensor pose of `core/simulate.py:84-92,202`

MuJoCo itself is replaced by oracle/mjcf_subset.py.  Output: `ModelConstants`, the exact
argument set of `dynamics.inverse` as bound by `functools.partial` at `core/simulate.py:150-156`.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import mjcf_subset as mj
from . import rnea_oracle as ro
from .rnea_oracle import SE3


@dataclass
class ModelConstants:
    hposes: list  # list[7] SE3: hposes_lj_kj (index 0 = identity for the world)
    simats: np.ndarray  # (7, 6, 6): simats_lj_l with bodies 7, 8 folded into index 6
    uscrews: np.ndarray  # (6, 6)
    twist_0: np.ndarray  # (6,)
    dtwist_0: np.ndarray  # (6,) = -[gravity, 0, 0, 0]
    pose_sen_llj: SE3  # static sensor pose w.r.t. the last link's joint frame
    pose_sen_obj: SE3
    pose_sen_obji: SE3
    simat_sen_obj: np.ndarray  # (6,6) attachment+object inertia in {llj} (named as in the reference)
    ground_truth: dict
    key_qpos: np.ndarray

    def hposes_Rt(self):
        """(7, 12): row-major R (9) then t (3) per pose."""
        return np.array([np.concatenate([np.asarray(h.rot.as_matrix()).reshape(9), np.asarray(h.trans, float)]) for h in self.hposes])


class PosesRegister:
    """reference transformations/poses.py:8-23."""

    def __init__(self, m: mj.MjModelLike, d: mj.MjDataLike):
        self.m = m
        self.a_b = ro.compose(m.body_pos, m.body_quat)  # :14
        self.b_bi = ro.compose(m.body_ipos, m.body_iquat)  # :15
        self.x_b = ro.compose(d.xpos, d.xmat)  # :16
        self.x_bi = ro.compose(d.xipos, d.ximat)  # :17
        self.x_site = ro.compose(d.site_xpos, d.site_xmat)  # :19
        self.l_lj = [SE3.identity()] + ro.compose(m.jnt_pos)  # :20
        self.lj_li = [l_lj.inv().dot(l_li) for l_lj, l_li in zip(self.l_lj, self.b_bi)]  # :21-23


def build_constants(manipulator_xml, target_csv, last_link="link6") -> ModelConstants:
    gt = mj.target_ground_truth(mj.read_cad_row(target_csv))  # core/core.py:202
    m = mj.compile_manipulator_with_target(manipulator_xml, gt)  # core/core.py:303-322
    d = mj.kinematics(m, m.key_qpos)  # core/core.py:327 + first forward pass
    return constants_from_model(m, d, gt, last_link)


def constants_from_model(m, d, gt, last_link="link6") -> ModelConstants:
    poses = PosesRegister(m, d)  # simulate.py:74
    id_ll = m.body_id(last_link)  # :78
    sl = slice(0, id_ll + 1)  # :79

    pose_x_obj = poses.x_b[m.body_id("target/object")]  # :84
    pose_obj_obji = poses.b_bi[m.body_id("target/object")]  # :85
    pose_x_obji = pose_x_obj.dot(pose_obj_obji)  # :86
    pose_x_sen = poses.x_site[m.site_id("target/ft_sensor")]  # :88
    pose_sen_obj = pose_x_sen.inv().dot(pose_x_obj)  # :89
    pose_sen_obji = pose_x_sen.inv().dot(pose_x_obji)  # :90
    pose_x_ll = poses.x_b[id_ll]  # :91
    pose_ll_llj = poses.l_lj[id_ll]  # :92

    uscrews = []  # :98-110
    for t, ax in zip(m.jnt_type, m.jnt_axis):
        us = np.zeros(6)
        if t == mj.SLIDE:
            us[:3] += ax
        elif t == mj.HINGE:
            us[3:] += ax
        else:
            raise TypeError("Only slide or hinge joints, represented as 2 or 3 for an element of m.jnt_type, are supported.")
        uscrews.append(us)
    uscrews = np.array(uscrews)

    simats_bi_b = ro.spatial_inertia_stack(m.body_mass, m.body_inertia)  # :115-117
    simats_lj_l = np.array([ro.transfer_simat(p, g) for p, g in zip(poses.lj_li, simats_bi_b[sl])])  # :119-123

    simat_sen_obj = np.zeros((6, 6))  # :129
    for pose_x_bi, simat_bi_b in zip(poses.x_bi[id_ll + 1 :], simats_bi_b[id_ll + 1 :]):  # :131-137
        pose_x_llj = pose_x_ll.dot(pose_ll_llj)
        pose_bi_llj = pose_x_bi.inv().dot(pose_x_llj)
        simat_llj_b = ro.transfer_simat(pose_bi_llj.inv(), simat_bi_b)
        simat_sen_obj += simat_llj_b
        simats_lj_l[id_ll] += simat_llj_b

    hposes = [SE3.identity()]  # :140-146
    for k in range(m.njnt):
        hpose_kj_lj = poses.l_lj[k].inv().dot(poses.a_b[k + 1].dot(poses.l_lj[k + 1]))
        hposes.append(hpose_kj_lj.inv())

    gacc_x = -1 * np.array([*m.gravity, 0, 0, 0])  # :149
    pose_sen_llj = pose_x_sen.inv().dot(pose_x_ll.dot(pose_ll_llj))  # :202 (configuration independent)
    return ModelConstants(
        hposes=hposes,
        simats=simats_lj_l,
        uscrews=uscrews,
        twist_0=np.zeros(6),
        dtwist_0=gacc_x,
        pose_sen_llj=pose_sen_llj,
        pose_sen_obj=pose_sen_obj,
        pose_sen_obji=pose_sen_obji,
        simat_sen_obj=simat_sen_obj,
        ground_truth=gt,
        key_qpos=np.array(m.key_qpos, float),
    )
