"""Batched B200 engine: a model handle plus the batched entry points, on torch CUDA tensors.

PyTorch is plumbing here (device memory, streams); all arithmetic happens in librbm_b200.so through
the C ABI of include/rbm_b200.h.  The constants are exactly the arguments the reference binds onto
`dynamics.inverse` with functools.partial (reference core/simulate.py:150-156).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _f64(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
    return a


def pose_to_Rt(pose) -> np.ndarray:
    """SE3-like (anything with .rot.as_matrix() and .trans), a 4x4 / 3x4 matrix, or 12 scalars -> [R row-major | t]."""
    if hasattr(pose, "rot") and hasattr(pose, "trans"):
        R = np.asarray(pose.rot.as_matrix(), dtype=np.float64)
        t = np.asarray(pose.trans, dtype=np.float64)
        return np.concatenate([R.reshape(9), t.reshape(3)])
    a = np.asarray(pose, dtype=np.float64)
    if a.shape == (12,):
        return a.copy()
    if a.shape in ((4, 4), (3, 4)):
        return np.concatenate([a[:3, :3].reshape(9), a[:3, 3]])
    raise ValueError("pose must be an SE3-like object, a 4x4 matrix or 12 scalars [R | t]")


def poses_to_Rt(poses) -> np.ndarray:
    if isinstance(poses, np.ndarray) and poses.ndim == 2 and poses.shape[1] == 12:
        return _f64(poses)
    return np.stack([pose_to_Rt(p) for p in poses])


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())


class Model:
    """Device-resident model constants (opaque `rbm_model*`) for one CUDA device."""

    def __init__(self, hposes_body_parent, simats_body, uscrews_body, twist_0, dtwist_0, wrench_tip=None,
                 pose_tip_ee=None, pose_sen_llj=None, device=None, force_generic=False):
        lib = _lib.load()
        self.uscrews = _f64(uscrews_body)
        if self.uscrews.ndim != 2 or self.uscrews.shape[1] != 6:
            raise ValueError("uscrews_body must have shape (nj, 6)")
        self.nj = nj = self.uscrews.shape[0]
        self.hposes_Rt = poses_to_Rt(hposes_body_parent)
        if self.hposes_Rt.shape != (nj + 1, 12):
            raise ValueError(f"hposes_body_parent must hold nj+1 = {nj + 1} poses (entry 0 is the unused world pose)")
        self.simats = _f64(simats_body, (nj + 1, 6, 6))
        self.twist_0 = _f64(twist_0, (6,))
        self.dtwist_0 = _f64(dtwist_0, (6,))
        self.wrench_tip = None if wrench_tip is None else _f64(wrench_tip, (6,))
        self.pose_tip = None if pose_tip_ee is None else pose_to_Rt(pose_tip_ee)
        self.pose_sen = None if pose_sen_llj is None else pose_to_Rt(pose_sen_llj)
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device = torch.device("cuda", int(device) if not isinstance(device, torch.device) else (device.index or 0))
        handle = C.c_void_p()
        rc = lib.rbm_model_create(nj, _ptr(self.hposes_Rt), _ptr(self.simats), _ptr(self.uscrews), _ptr(self.twist_0),
                                  _ptr(self.dtwist_0), _ptr(self.wrench_tip), _ptr(self.pose_tip), _ptr(self.pose_sen),
                                  _lib.FLAG_FORCE_GENERIC if force_generic else 0, self.device.index, C.byref(handle))
        _lib.check(rc, "rbm_model_create")
        self._h = handle
        self._lib = lib
        self.kernel_path = _lib.PATH_NAMES[lib.rbm_model_kernel_path(handle)]

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rbm_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers ----------------------------------------------------------------------------
    def _check_dev(self, *tensors):
        for t in tensors:
            if t is None:
                continue
            if not t.is_cuda or t.device != self.device:
                raise ValueError(f"tensor must live on {self.device} (got {t.device}); there is no CPU path")
            if not t.is_contiguous():
                raise ValueError("tensor must be contiguous")

    @staticmethod
    def _stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # ---- batched inverse dynamics: SoA [nj][n] ----------------------------------------------
    def rnea(self, q, qd, qdd, tau=None, want_twists=False):
        """q, qd, qdd: CUDA tensors (nj, n) float64 or float32.  Returns tau (nj, n) [, V (6, n), dV (6, n)]."""
        self._check_dev(q, qd, qdd, tau)
        if q.shape != qd.shape or q.shape != qdd.shape or q.dim() != 2 or q.shape[0] != self.nj:
            raise ValueError(f"q, qd, qdd must all have shape ({self.nj}, n)")
        if q.dtype not in (torch.float64, torch.float32) or qd.dtype != q.dtype or qdd.dtype != q.dtype:
            raise ValueError("q, qd, qdd must share dtype float64 or float32")
        n = q.shape[1]
        if tau is None:
            tau = torch.empty_like(q)
        elif tau.shape != q.shape or tau.dtype != q.dtype:
            raise ValueError("tau must match q in shape and dtype")
        V = dV = None
        if want_twists:
            V = torch.empty((6, n), dtype=q.dtype, device=q.device)
            dV = torch.empty((6, n), dtype=q.dtype, device=q.device)
        fn = self._lib.rbm_rnea_f64 if q.dtype == torch.float64 else self._lib.rbm_rnea_f32
        with torch.cuda.device(self.device):
            rc = fn(self._h, _ptr(q), _ptr(qd), _ptr(qdd), _ptr(tau), _ptr(V), _ptr(dV), n, n, self._stream())
        _lib.check(rc, "rbm_rnea")
        return (tau, V, dV) if want_twists else tau

    # ---- batched inverse dynamics: AoS [n][3][nj], the reference's own `traj` layout ------------
    def rnea_aos(self, traj, tau=None):
        self._check_dev(traj, tau)
        if traj.dim() != 3 or traj.shape[1] != 3 or traj.shape[2] != self.nj:
            raise ValueError(f"traj must have shape (n, 3, {self.nj})")
        if traj.dtype not in (torch.float64, torch.float32):
            raise ValueError("traj must be float64 or float32")
        n = traj.shape[0]
        if tau is None:
            tau = torch.empty((n, self.nj), dtype=traj.dtype, device=traj.device)
        fn = self._lib.rbm_rnea_aos_f64 if traj.dtype == torch.float64 else self._lib.rbm_rnea_aos_f32
        with torch.cuda.device(self.device):
            rc = fn(self._h, _ptr(traj), _ptr(tau), n, self._stream())
        _lib.check(rc, "rbm_rnea_aos")
        return tau

    def rnea_full(self, traj):
        """Everything `dynamics.inverse` returns, per sample: tau (n,nj), poses (n,nj,12), twists, dtwists (n,nj+1,6)."""
        self._check_dev(traj)
        if traj.dim() != 3 or traj.shape[1] != 3 or traj.shape[2] != self.nj or traj.dtype != torch.float64:
            raise ValueError(f"traj must be float64 with shape (n, 3, {self.nj})")
        n, nj = traj.shape[0], self.nj
        kw = dict(dtype=torch.float64, device=traj.device)
        tau = torch.empty((n, nj), **kw)
        poses = torch.empty((n, nj, 12), **kw)
        tw = torch.empty((n, nj + 1, 6), **kw)
        dtw = torch.empty((n, nj + 1, 6), **kw)
        with torch.cuda.device(self.device):
            rc = self._lib.rbm_rnea_full_f64(self._h, _ptr(traj), _ptr(tau), _ptr(poses), _ptr(tw), _ptr(dtw), n, self._stream())
        _lib.check(rc, "rbm_rnea_full")
        return tau, poses, tw, dtw

    # ---- host end-to-end ---------------------------------------------------------------------------
    def rnea_host(self, traj, tau=None, chunk=0):
        """traj: HOST array / pinned CPU tensor (n, 3, nj); returns tau on the host (same kind).  H2D, kernel and D2H are
        pipelined inside the library (rbm_rnea_host_*)."""
        is_t = isinstance(traj, torch.Tensor)
        if is_t and traj.is_cuda:
            raise ValueError("rnea_host takes host memory; use rnea_aos for device tensors")
        arr = traj if is_t else np.ascontiguousarray(traj)
        shape, dt = tuple(arr.shape), (arr.dtype if not is_t else {torch.float64: np.float64, torch.float32: np.float32}[arr.dtype])
        if len(shape) != 3 or shape[1] != 3 or shape[2] != self.nj:
            raise ValueError(f"traj must have shape (n, 3, {self.nj})")
        n = shape[0]
        if tau is None:
            tau = torch.empty((n, self.nj), dtype=arr.dtype, pin_memory=True) if is_t else np.empty((n, self.nj), dtype=dt)
        fn = self._lib.rbm_rnea_host_f64 if np.dtype(dt) == np.float64 else self._lib.rbm_rnea_host_f32
        rc = fn(self._h, _ptr(arr), _ptr(tau), n, int(chunk))
        _lib.check(rc, "rbm_rnea_host")
        return tau
