"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product package.

A FUNCTIONAL stand-in for the few MuJoCo entry points the reference's closed loop drives (`mj_step`, `mj_forward`,
`mj_differentiatePos`, `mjd_transitionFD`, `mj_resetDataKeyframe`, `mj_name2id`) on the reference's own model, so that the reference's
`core/simulate.py::simulate`, `controllers/lqr.py::LinearQuadraticRegulator`, `sensors/sensors.py`, `transformations/poses.py` can be
EXECUTED UNMODIFIED in a container without MuJoCo (oracle/reference_loader.load_simulation).  What is pinned that way is the
reference's control flow and bookkeeping -- which quantities it reads when, the control law, the frame schedule, the sensor-frame
transformation, the noise model; the physics below is a restatement of MuJoCo's published behaviour for this model, not MuJoCo:

  * model: oracle/mjcf_subset.py (bodies, joints, sites of sequential.xml + the attached target, core/core.py:296-322) plus the
    name tables and the sensor layout of sequential.xml:51-78 (51 sensordata values; `force` at 45:48, `torque` at 48:51)
  * mj_forward: forward kinematics into d.xpos / xmat / xipos / ximat / site_x* IN PLACE (the reference's pose registers alias these
    arrays), qacc = M(q)^-1 (ctrl - bias) (oracle/lqr_oracle.py), force / torque sensors = the Newton-Euler wrench of the body behind the
    site in the site frame (MuJoCo: cfrc_int of the site's body); every other sensor reads 0 (the reference does not consume them)
  * mj_step: mj_forward, then semi-implicit Euler (qvel += dt qacc; qpos += dt qvel), time += dt
  * mj_differentiatePos(m, out, dt, qpos1, qpos2): out = (qpos2 - qpos1) / dt (slide and hinge joints)
  * mjd_transitionFD: oracle/lqr_oracle.transition_fd at (qpos, qvel, ctrl); leaves `d` as after a forward pass at the nominal state
    (MuJoCo leaves the state of its last perturbed evaluation: identical up to the O(eps) perturbation)
"""
from __future__ import annotations

import numpy as np

from . import lqr_oracle as lo
from . import mjcf_subset as mj
from . import model_oracle as mo
from . import replay_oracle as ro
from . import rnea_vec as rv

SENSORS = [("x", 1), ("y", 1), ("z", 1), ("roll", 1), ("pitch", 1), ("yaw", 1)] + [
    (f"{kind}_x_{who}", 3) for kind in ("linvel", "angvel", "linacc", "angacc") for who in ("obj", "obji", "sen")
] + [("linacc_sen", 3), ("force", 3), ("torque", 3)]

# mjtObj values of oracle/shims/mujoco/_enums.py
OBJ_BODY, OBJ_JOINT, OBJ_SITE, OBJ_CAMERA, OBJ_SENSOR, OBJ_NUMERIC, OBJ_KEY = 1, 3, 6, 7, 18, 19, 23


class StandinModel:
    def __init__(self, manipulator_xml, target_csv):
        self.gt = mj.target_ground_truth(mj.read_cad_row(target_csv))
        self._m = mj.compile_manipulator_with_target(manipulator_xml, self.gt)
        for k in ("body_names", "body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia", "jnt_type", "jnt_axis", "jnt_pos",
                  "site_names", "key_qpos", "gravity", "timestep"):
            setattr(self, k, getattr(self._m, k))
        self.nq = self.nv = self.nu = self.njnt = self._m.njnt
        self.nbody = len(self.body_names)
        self.na = 0
        self.sensor_names = [s for s, _ in SENSORS]
        self.sensor_dim = np.array([n for _, n in SENSORS])
        self.nsensordata = int(self.sensor_dim.sum())
        self.key_names = ["initial_state"]
        self.cam_names = ["tracking"]
        self.numeric_names = ["target/aabb_scale"]
        self.numeric_data = np.array([self.gt["aabb_scale"]])
        d0 = mj.kinematics(self._m, self._m.key_qpos)
        c = mo.constants_from_model(self._m, d0, self.gt)
        self.consts = dict(hposes_Rt=c.hposes_Rt(), simats=c.simats, uscrews=c.uscrews, twist_0=c.twist_0, dtwist_0=c.dtwist_0)
        self.pose_sen_Rt = np.concatenate([np.asarray(c.pose_sen_llj.rot.as_matrix()).reshape(9), np.asarray(c.pose_sen_llj.trans, float)])
        self.G_sensed = ro.sensor_inertia(c.simat_sen_obj, self.pose_sen_Rt)
        self.model_constants = c

    def forward_kinematics(self, d):
        """look-alike protocol of rigid_body_manipulation_b200.mujoco_bridge.refresh_kinematics (what mujoco.mj_forward does for a real model)"""
        mj_forward(self, d)

    @property
    def opt(self):
        class _Opt:
            timestep = self.timestep
            gravity = self.gravity
        return _Opt()


class StandinData:
    def __init__(self, m: StandinModel):
        nb, ns = len(m.body_names), len(m.site_names)
        self.qpos, self.qvel, self.qacc, self.ctrl = (np.zeros(m.nv) for _ in range(4))
        self.time = 0.0
        self.sensordata = np.zeros(m.nsensordata)
        self.xpos, self.xmat, self.xipos, self.ximat = np.zeros((nb, 3)), np.zeros((nb, 9)), np.zeros((nb, 3)), np.zeros((nb, 9))
        self.site_xpos, self.site_xmat = np.zeros((ns, 3)), np.zeros((ns, 9))
        self.cam_xpos, self.cam_xmat = np.zeros((1, 3)), np.eye(3).reshape(1, 9).copy()


def mj_name2id(m, objtype, name):
    table = {OBJ_BODY: m.body_names, OBJ_SITE: m.site_names, OBJ_SENSOR: m.sensor_names, OBJ_KEY: m.key_names, OBJ_CAMERA: m.cam_names,
             OBJ_NUMERIC: m.numeric_names, OBJ_JOINT: [f"joint{k}" for k in range(m.nv)]}[int(objtype)]
    return table.index(name) if name in table else -1


def mj_resetDataKeyframe(m, d, key):
    d.qpos[:] = m.key_qpos
    d.qvel[:] = 0.0
    d.qacc[:] = 0.0
    d.ctrl = np.zeros(m.nu)
    d.time = 0.0


def mj_forward(m, d):
    k = mj.kinematics(m._m, d.qpos)
    for name in ("xpos", "xmat", "xipos", "ximat", "site_xpos", "site_xmat"):
        getattr(d, name)[:] = getattr(k, name)
    ctrl = np.asarray(d.ctrl, float)
    d.qacc[:] = lo.forward_dynamics(m.consts, d.qpos[None], d.qvel[None], ctrl[None])[0]
    wrench = ro._ft_reading(m.G_sensed, *ro._sensor_state(m.consts, m.pose_sen_Rt, d.qpos, d.qvel, d.qacc))
    d.sensordata[:] = 0.0
    d.sensordata[45:51] = wrench


def mj_step(m, d):
    mj_forward(m, d)
    d.qvel[:] = d.qvel + m.timestep * d.qacc
    d.qpos[:] = d.qpos + m.timestep * d.qvel
    d.time = d.time + m.timestep


def mj_differentiatePos(m, qvel, dt, qpos1, qpos2):
    qvel[:] = (np.asarray(qpos2, float) - np.asarray(qpos1, float)) / dt


def mjd_transitionFD(m, d, eps, flg_centered, A, B, C, D):
    a, b = lo.transition_fd(m.consts, d.qpos[None], d.qvel[None], np.asarray(d.ctrl, float)[None], dt=m.timestep, eps=eps, centered=bool(flg_centered))
    if A is not None:
        A[:] = a[0]
    if B is not None:
        B[:] = b[0]
    mj_forward(m, d)
