"""Drop-in for reference dynamics/dynamics.py -- same public names, signatures, defaults and error behaviour; every
function evaluates on the GPU through librbm_b200.so (there is no CPU fallback: without a CUDA device the calls raise).

  reference symbol (dynamics/dynamics.py)            -> C ABI entry point
  inverse :109-157                                    rbm_model_create + rbm_rnea_full_host_f64   (batched: rbm_rnea_*)
  transfer_simat :72-106                              rbm_transfer_simat_f64
  get_spatial_inertia_matrix :62-69                   rbm_spatial_inertia_f64
  get_regressor_matrix :215-249                       rbm_regressor_rows_f64
  extract_linvel/linacc_frame_transferred :160-212    rbm_point_motion_f64
  coordinate_transfer_imat :252-257                   rbm_coordinate_transfer_imat_f64
  coordinate_transfer_simat :260-263                  rbm_coordinate_transfer_simat_f64
  StateSpace / StateSpaceConfig :14-46                rbm_linearize_f64 (RNEA-based transition linearisation)

Additive batched entry points: inverse_batched, regressor_batched, make_model (they take / return torch CUDA tensors).
"""
from __future__ import annotations

from collections.abc import Sequence
from dataclasses import dataclass
from typing import Union

import numpy as np
import torch
from numpy.typing import NDArray

from rigid_body_manipulation_b200 import engine as _engine
from rigid_body_manipulation_b200.lie import SE3, is_se3, se3_from_Rt

__all__ = [
    "StateSpaceConfig", "StateSpace", "get_spatial_inertia_matrix", "transfer_simat", "inverse",
    "extract_linvel_frame_transferred", "extract_linacc_frame_transferred", "get_regressor_matrix",
    "coordinate_transfer_imat", "coordinate_transfer_simat",
    "make_model", "inverse_batched", "regressor_batched",
]


def _np(t):
    return t.detach().cpu().numpy()


# ---------------------------------------------------------------------------------------------------------------
# model cache: the reference passes the constants on every call (bound once with functools.partial,
# core/simulate.py:150-156); the device-resident handle is looked up by the constants' content.
# ---------------------------------------------------------------------------------------------------------------
_MODELS: dict = {}
_MODELS_MAX = 32


def make_model(hposes_body_parent, simats_body, uscrews_body, twist_0, dtwist_0, wrench_tip=None, pose_tip_ee=None,
               pose_sen_llj=None, device=None) -> _engine.Model:
    """Device-resident constants for the batched API (same arguments as `inverse` binds)."""
    return _engine.Model(hposes_body_parent, simats_body, uscrews_body, twist_0, dtwist_0, wrench_tip, pose_tip_ee, pose_sen_llj, device)


def _pose_key(p):
    if hasattr(p, "rot"):
        return np.asarray(p.rot.as_matrix(), dtype=np.float64).tobytes() + np.asarray(p.trans, dtype=np.float64).tobytes()
    return np.asarray(p, dtype=np.float64).tobytes()


def _cached_model(hposes_body_parent, simats_body, uscrews_body, twist_0, dtwist_0, wrench_tip, pose_tip_ee):
    """Content-keyed lookup (a few microseconds of hashing per call; the arrays are mutable, so identity is not enough)."""
    nj = len(uscrews_body)
    key = (_engine.current_device(), np.asarray(uscrews_body, dtype=np.float64).tobytes(), np.asarray(simats_body, dtype=np.float64).tobytes(),
           np.asarray(twist_0, dtype=np.float64).tobytes(), np.asarray(dtwist_0, dtype=np.float64).tobytes(),
           np.asarray(wrench_tip, dtype=np.float64).tobytes(), _pose_key(pose_tip_ee)) + tuple(_pose_key(h) for h in hposes_body_parent[: nj + 1])
    m = _MODELS.get(key)
    if m is None:
        if len(_MODELS) >= _MODELS_MAX:
            _MODELS.pop(next(iter(_MODELS)))
        m = _MODELS[key] = _engine.Model(_engine.poses_to_Rt(hposes_body_parent[: nj + 1]), np.asarray(simats_body, dtype=np.float64)[: nj + 1],
                                         uscrews_body, twist_0, dtwist_0, wrench_tip, _engine.pose_to_Rt(pose_tip_ee))
    return m


# ---------------------------------------------------------------------------------------------------------------
# state space (LQR linearisation)
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class StateSpaceConfig:
    epsilon: float = 1e-8
    centered: bool = True


class StateSpace:
    """A, B of the discrete-time transition x+ = f(x, u), x = [q; qd] (reference dynamics.py:19-46).

    `m` / `d` may be MuJoCo's MjModel / MjData (constants are then derived exactly as reference core/simulate.py:74-156
    does, through this package's own `transfer_simat` / `Poses`) or a prepared pair
    (rigid_body_manipulation_b200.engine.Model or model.Constants, state) where state has qpos, qvel[, ctrl].
    C and D (sensor Jacobians) are allocated with the reference's shapes but left zero: no caller reads them
    (controllers/lqr.py:48-49 uses A and B only) and they depend on MuJoCo's sensor pipeline.
    """

    def __init__(self, cfg: StateSpaceConfig, m, d) -> None:
        self.epsilon = cfg.epsilon
        self.centered = cfg.centered
        nv = int(getattr(m, "nv", getattr(m, "nj", 0)) or len(np.atleast_1d(_state_field(d, "qpos"))))
        na = int(getattr(m, "na", 0))
        nu = int(getattr(m, "nu", nv))
        nsens = int(getattr(m, "nsensordata", 0))
        self.ns = 2 * nv + na
        self.nsensordata = nsens
        self.A = np.zeros((self.ns, self.ns))
        self.B = np.zeros((self.ns, nu))
        self.C = np.zeros((nsens, self.ns))
        self.D = np.zeros((nsens, nu))
        self._model = None
        self.update_matrices(m, d)

    def _engine_model(self, m, d):
        if self._model is None:
            if isinstance(m, _engine.Model):
                self._model = m
            elif hasattr(m, "hposes_Rt") and hasattr(m, "simats"):  # model.Constants
                self._model = _engine.Model(m.hposes_Rt, m.simats, m.uscrews, m.twist_0, m.dtwist_0)
            else:
                from rigid_body_manipulation_b200.mujoco_bridge import constants_from_mujoco

                c = constants_from_mujoco(m, d)
                self._model = _engine.Model(c["hposes"], c["simats"], c["uscrews"], c["twist_0"], c["dtwist_0"])
        return self._model

    def update_matrices(self, m, d) -> None:
        if not isinstance(m, _engine.Model) and not hasattr(m, "hposes_Rt"):
            from rigid_body_manipulation_b200.mujoco_bridge import refresh_kinematics

            refresh_kinematics(m, d)  # mjd_transitionFD leaves d evaluated at the current state; callers rely on it
        model = self._engine_model(m, d)
        dt = float(getattr(getattr(m, "opt", None), "timestep", getattr(m, "timestep", 0.002)))
        dev = model.device
        q = torch.as_tensor(np.asarray(_state_field(d, "qpos"), dtype=np.float64).reshape(-1, 1), device=dev).contiguous()
        qd = torch.as_tensor(np.asarray(_state_field(d, "qvel"), dtype=np.float64).reshape(-1, 1), device=dev).contiguous()
        ctrl = _state_field(d, "ctrl", None)
        u = None if ctrl is None else torch.as_tensor(np.asarray(ctrl, dtype=np.float64).reshape(-1, 1), device=dev).contiguous()
        A, B = model.linearize(q, qd, u, dt=dt, eps=self.epsilon, centered=self.centered)
        self.A[...] = _np(A[0])
        self.B[...] = _np(B[0])


def _state_field(d, name, default=...):
    if isinstance(d, dict):
        v = d.get(name, default)
    else:
        v = getattr(d, name, default)
    if v is ...:
        raise ValueError(f"state object has no '{name}'")
    return v


# ---------------------------------------------------------------------------------------------------------------
# spatial inertias
# ---------------------------------------------------------------------------------------------------------------
def get_spatial_inertia_matrix(mass, diagonal_inertia):
    assert len(mass) == len(diagonal_inertia), (
        "Lenght of 'mass' of the bodies and that of 'diagonal_inertia' vectors must match."
    )
    if len(mass) == 0:
        return np.zeros((0, 6, 6))
    return _np(_engine.spatial_inertia(np.asarray(mass, dtype=np.float64), np.asarray(diagonal_inertia, dtype=np.float64).reshape(-1, 3)))


def transfer_simat(pose: Union[SE3, Sequence], simat: NDArray) -> NDArray:
    """Spatial inertia `simat` expressed in {b} -> expressed in {a}, with `pose` = T_ab (the reference's call-site
    semantics: Ad(T_ab^-1)^T G Ad(T_ab^-1), dynamics.py:102-104).  Single pose or single matrix -> one (6,6) result."""
    single_pose = is_se3(pose)
    poses = [pose] if single_pose else list(pose)
    simat = np.asarray(simat, dtype=np.float64)
    single_simat = 2 == simat.ndim
    if single_simat:
        simat = np.expand_dims(simat, 0)
    assert len(poses) == len(simat), ValueError("The numbers of spatial inertia tensors and SE3 instances do not match.")
    out = _np(_engine.transfer_simat(_engine.poses_to_Rt(poses), simat))
    return out[0] if single_pose or single_simat else out


def coordinate_transfer_imat(pose_target_current, imat_current, mass):
    Rt = _engine.pose_to_Rt(pose_target_current)[None]
    return _np(_engine.coordinate_transfer_imat(Rt, np.asarray(imat_current, dtype=np.float64)[None], np.array([mass], dtype=np.float64)))[0]


def coordinate_transfer_simat(pose_target_current, simat_current):
    Rt = _engine.pose_to_Rt(pose_target_current)[None]
    return _np(_engine.transfer_simat(Rt, np.asarray(simat_current, dtype=np.float64)[None], adjoint_form=True))[0]


# ---------------------------------------------------------------------------------------------------------------
# inverse dynamics
# ---------------------------------------------------------------------------------------------------------------
def inverse(
    traj: np.ndarray,
    hposes_body_parent,
    simats_body: np.ndarray,
    uscrews_body: np.ndarray,
    twist_0: np.ndarray,
    dtwist_0: np.ndarray,
    wrench_tip: np.ndarray = np.zeros(6),
    pose_tip_ee=SE3.identity(),
):
    """Recursive Newton-Euler inverse dynamics of one trajectory sample, `traj` = rows (q, qd, qdd).

    Returns, like the reference, (tau (n,), poses list[n+1] of SE3 (T_{i,i-1}, then the tip pose), twists list[n+1],
    dtwists list[n+1]) with index 0 of the twist lists being the `twist_0` / `dtwist_0` objects that were passed in.
    """
    uscrews_body = np.asarray(uscrews_body)
    nj = len(uscrews_body)
    traj = np.asarray(traj, dtype=np.float64)
    if traj.shape != (3, nj):
        raise ValueError(f"traj must have shape (3, {nj})")
    model = _cached_model(hposes_body_parent, simats_body, uscrews_body, twist_0, dtwist_0, wrench_tip, pose_tip_ee)
    tau, poses, tw, dtw = model.rnea_full_host(traj[None])  # one H2D, one launch, one D2H (rbm_rnea_full_host_f64)
    tau_h, poses_h, tw_h, dtw_h = tau[0], poses[0], tw[0], dtw[0]
    pose_list = [se3_from_Rt(p) for p in poses_h] + [pose_tip_ee]
    twists = [twist_0] + [tw_h[i + 1] for i in range(nj)]
    dtwists = [dtwist_0] + [dtw_h[i + 1] for i in range(nj)]
    return tau_h, pose_list, twists, dtwists


def inverse_batched(model: _engine.Model, q, qd, qdd, want_twists=False):
    """Batched, device-resident: q, qd, qdd CUDA tensors (nj, n) -> tau (nj, n) [, V_last (6, n), dV_last (6, n)]."""
    return model.rnea(q, qd, qdd, want_twists=want_twists)


# ---------------------------------------------------------------------------------------------------------------
# frame-transferred point motion (reference dynamics.py:160-212)
# ---------------------------------------------------------------------------------------------------------------
def extract_linvel_frame_transferred(twist: NDArray, pose, homogeneous: bool = False) -> NDArray:
    lv, _ = _engine.point_motion(np.asarray(twist, dtype=np.float64)[None], None, np.asarray(pose.trans, dtype=np.float64)[None], want_acc=False)
    out = _np(lv)[0]
    return np.append(out, 0.0) if homogeneous else out


def extract_linacc_frame_transferred(twist: NDArray, dtwist: NDArray, pose, homogeneous: bool = False) -> NDArray:
    _, la = _engine.point_motion(np.asarray(twist, dtype=np.float64)[None], np.asarray(dtwist, dtype=np.float64)[None],
                                 np.asarray(pose.trans, dtype=np.float64)[None])
    out = _np(la)[0]
    return np.append(out, 0.0) if homogeneous else out


# ---------------------------------------------------------------------------------------------------------------
# regressor (reference dynamics.py:215-249)
# ---------------------------------------------------------------------------------------------------------------
def get_regressor_matrix(twist: NDArray, dtwist: NDArray) -> NDArray:
    twist, dtwist = np.asarray(twist, dtype=np.float64), np.asarray(dtwist, dtype=np.float64)
    if twist.shape != (6,) or dtwist.shape != (6,):
        raise ValueError("twist and dtwist must have 6 elements")
    return _np(_engine.regressor_rows(twist[None], dtwist[None]))[0]


def regressor_batched(twists, dtwists):
    """(n, 6), (n, 6) host or device -> (n, 6, 10) device tensor."""
    return _engine.regressor_rows(twists, dtwists)
