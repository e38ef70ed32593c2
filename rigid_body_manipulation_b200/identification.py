"""Inertial-parameter identification from the fused Gram pack (replaces reference loggers/loggers.py:127-129).

The reference stacks every frame's 6x10 regressor into a (6F, 10) matrix and calls np.linalg.lstsq.  At 1e8 samples
that matrix is 48 GB, so the device accumulates the normal equations instead (rbm_regressor_gram_*), shards all-reduce
the 112-double pack, and the 10x10 system is solved here in float64 on the host (once per solve, O(1e3) flops).

pack layout: [Y^T Y (100, row-major) | Y^T f (10) | f^T f | n_samples].
Parameter order: [m, m cx, m cy, m cz, Ixx, Iyy, Izz, Ixy, Iyz, Izx] about the sensor-frame origin
(reference dynamics.py:225-230, loggers.py:133).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

PARAM_LABELS = ["total_mass", "mx", "my", "mz", "ixx", "iyy", "izz", "ixy", "iyz", "izx"]


@dataclass
class Identification:
    phi: np.ndarray          # (10,)
    n_samples: float
    residual_ss: float       # ||Y phi - f||^2 from the normal equations (cancellation-limited: ~1e-16 * f_ss)
    f_ss: float              # ||f||^2
    rms_residual: float      # per scalar wrench component
    cond: float              # condition number of the column-scaled Gram
    rank: int


def unpack(pack):
    p = np.asarray(pack.detach().cpu().numpy() if hasattr(pack, "detach") else pack, dtype=np.float64)
    if p.shape != (112,):
        raise ValueError("gram pack must have 112 elements")
    return p[:100].reshape(10, 10), p[100:110], float(p[110]), float(p[111])


def solve(pack, rcond: float = 1e-13) -> Identification:
    """Minimum-norm least-squares solution of the normal equations, column-equilibrated (the columns of Y mix m/s^2,
    rad/s^2 and (rad/s)^2 scales) and solved by a symmetric eigen-decomposition with a relative cut-off, so rank-deficient
    excitations (e.g. a trajectory that never rotates) behave like np.linalg.lstsq's minimum-norm answer."""
    G, b, ff, n = unpack(pack)
    G = 0.5 * (G + G.T)
    d = np.sqrt(np.clip(np.diag(G), 0.0, None))
    d[d == 0.0] = 1.0
    Gs = G / np.outer(d, d)
    w, V = np.linalg.eigh(Gs)
    keep = w > rcond * max(w.max(), 0.0)
    z = V[:, keep] @ ((V[:, keep].T @ (b / d)) / w[keep])
    phi = z / d
    rss = max(ff - 2.0 * phi @ b + phi @ G @ phi, 0.0)
    cond = float(w.max() / w[keep].min()) if keep.any() else float("inf")
    return Identification(phi=phi, n_samples=n, residual_ss=rss, f_ss=ff, rms_residual=float(np.sqrt(rss / max(6.0 * n, 1.0))), cond=cond, rank=int(keep.sum()))


def score(estimate, gt_params, aabb_scale: float) -> float:
    """The reference's normalised squared error (main.py:21-38): mass / first moments / inertias made dimensionless with
    m, m*L, m*L^2 (L = aabb_scale), averaged over the 10 parameters."""
    estimate, gt = np.asarray(estimate, dtype=float), np.asarray(gt_params, dtype=float)
    m = gt[0]
    s = ((estimate[0] - gt[0]) ** 2) / (m * aabb_scale**0) ** 2
    s += ((estimate[1:4] - gt[1:4]) ** 2).sum() / (m * aabb_scale**1) ** 2
    s += ((estimate[4:10] - gt[4:10]) ** 2).sum() / (m * aabb_scale**2) ** 2
    return float(s / 10.0)


def sensor_frame_params(target, pose_sen_obj_Rt) -> np.ndarray:
    """Ground-truth phi of a target expressed in the SENSOR frame (what the regressor identifies), from the body-frame CAD
    values: c_s = R c + t, I_s(origin) = R I_com R^T + m (|c_s|^2 1 - c_s c_s^T)."""
    Rt = np.asarray(pose_sen_obj_Rt, dtype=float)
    R, t = Rt[:9].reshape(3, 3), Rt[9:]
    c = R @ target.com + t
    I0 = R @ target.inertia_com @ R.T + target.mass * (c @ c * np.eye(3) - np.outer(c, c))
    return np.array([target.mass, *(target.mass * c), I0[0, 0], I0[1, 1], I0[2, 2], I0[0, 1], I0[1, 2], I0[2, 0]])
