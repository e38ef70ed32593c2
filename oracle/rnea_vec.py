"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Vectorised (batch-over-samples) fp64 numpy restatement of the same path as
oracle/rnea_oracle.py -- reference `dynamics/dynamics.py:109-157` (inverse),
`core/simulate.py:202-209` (sensor-frame twists), `dynamics/dynamics.py:215-249` (regressor),
`loggers/loggers.py:127-129` (least-squares identification) -- so that million-sample
parity checks finish in seconds.  It is validated against the per-sample oracle and the
reference-generated golden vectors in tests/test_oracle_*.py; it is NOT the reference's
cost profile (the per-sample oracle is) and is never reported as the reference arm.

Layout: everything batch-first, trajectories (B, 3, nj), twists (B, 6) in [v; w] order.
"""
from __future__ import annotations

import numpy as np


def skew(v):
    """(..., 3) -> (..., 3, 3) cross-product matrices (liegroups SO3.wedge)."""
    v = np.asarray(v, float)
    S = np.zeros(v.shape[:-1] + (3, 3))
    S[..., 0, 1] = -v[..., 2]
    S[..., 1, 0] = v[..., 2]
    S[..., 0, 2] = v[..., 1]
    S[..., 2, 0] = -v[..., 1]
    S[..., 1, 2] = -v[..., 0]
    S[..., 2, 1] = v[..., 0]
    return S


def so3_exp_and_jacobian(phi):
    """Batched liegroups SO3.exp / SO3.left_jacobian incl. the isclose(angle, 0) first-order branch."""
    angle = np.linalg.norm(phi, axis=-1)
    small = np.isclose(angle, 0.0)
    safe = np.where(small, 1.0, angle)
    axis = phi / safe[..., None]
    s, c = np.sin(angle), np.cos(angle)
    I = np.eye(3)
    aa = axis[..., :, None] * axis[..., None, :]
    K = skew(axis)
    R = c[..., None, None] * I + (1 - c)[..., None, None] * aa + s[..., None, None] * K
    sa = s / safe
    J = sa[..., None, None] * I + (1 - sa)[..., None, None] * aa + ((1 - c) / safe)[..., None, None] * K
    W = skew(phi)
    R = np.where(small[..., None, None], I + W, R)
    J = np.where(small[..., None, None], I + 0.5 * W, J)
    return R, J


def adjoint(R, t):
    """(B,3,3),(B,3) -> (B,6,6)  [[R, [t]x R],[0, R]]."""
    Ad = np.zeros(R.shape[:-2] + (6, 6))
    Ad[..., :3, :3] = R
    Ad[..., 3:, 3:] = R
    Ad[..., :3, 3:] = skew(t) @ R
    return Ad


def curlywedge(V):
    """(B,6) -> (B,6,6)  [[ [w]x, [v]x ],[0, [w]x ]]."""
    ad = np.zeros(V.shape[:-1] + (6, 6))
    Ww = skew(V[..., 3:])
    ad[..., :3, :3] = Ww
    ad[..., 3:, 3:] = Ww
    ad[..., :3, 3:] = skew(V[..., :3])
    return ad


def inverse_batched(traj, hposes_Rt, simats, uscrews, twist_0, dtwist_0, wrench_tip=None, pose_tip_Rt=None):
    """Batched reference dynamics.py:109-157.

    traj (B,3,nj); hposes_Rt (nj+1,12) rows [R row-major | t] (entry 0 unused); simats (nj+1,6,6);
    uscrews (nj,6).  Returns dict(tau (B,nj), poses (B,nj,12), twists (B,nj+1,6), dtwists (B,nj+1,6)).
    """
    traj = np.asarray(traj, float)
    B, _, nj = traj.shape
    hposes_Rt = np.asarray(hposes_Rt, float)
    V = np.broadcast_to(np.asarray(twist_0, float), (B, 6))
    dV = np.broadcast_to(np.asarray(dtwist_0, float), (B, 6))
    twists, dtwists, Ads, poses = [V], [dV], [], np.zeros((B, nj, 12))
    for i in range(nj):
        S = np.asarray(uscrews[i], float)
        q, qd, qdd = traj[:, 0, i], traj[:, 1, i], traj[:, 2, i]
        xi = -S[None, :] * q[:, None]
        Re, Je = so3_exp_and_jacobian(xi[:, 3:])
        te = np.einsum("bij,bj->bi", Je, xi[:, :3])
        Rh, th = hposes_Rt[i + 1, :9].reshape(3, 3), hposes_Rt[i + 1, 9:]
        R = Re @ Rh
        t = np.einsum("bij,j->bi", Re, th) + te
        poses[:, i, :9] = R.reshape(B, 9)
        poses[:, i, 9:] = t
        Ad = adjoint(R, t)
        V = np.einsum("bij,bj->bi", Ad, twists[-1]) + S * qd[:, None]
        dV = np.einsum("bij,bj->bi", Ad, dtwists[-1]) + np.einsum("bij,j->bi", curlywedge(V), S) * qd[:, None] + S * qdd[:, None]
        Ads.append(Ad)
        twists.append(V)
        dtwists.append(dV)
    if pose_tip_Rt is None:
        Ad_tip = np.broadcast_to(np.eye(6), (B, 6, 6))
    else:
        p = np.asarray(pose_tip_Rt, float)
        Ad_tip = np.broadcast_to(adjoint(p[:9].reshape(3, 3), p[9:]), (B, 6, 6))
    Ads.append(Ad_tip)
    F = np.broadcast_to(np.zeros(6) if wrench_tip is None else np.asarray(wrench_tip, float), (B, 6))
    tau = np.zeros((B, nj))
    for i in range(nj, 0, -1):
        G = np.asarray(simats[i], float)
        GV = twists[i] @ G.T
        F = (
            np.einsum("bji,bj->bi", Ads[i], F)
            + dtwists[i] @ G.T
            - np.einsum("bji,bj->bi", curlywedge(twists[i]), GV)
        )
        tau[:, i - 1] = F @ np.asarray(uscrews[i - 1], float)
    return dict(tau=tau, poses=poses, twists=np.stack(twists, 1), dtwists=np.stack(dtwists, 1))


def sensor_frame_twists_batched(pose_sen_Rt, V, dV):
    """Batched reference core/simulate.py:202-209 (the ad(V_s) Ad V term is evaluated, as there)."""
    p = np.asarray(pose_sen_Rt, float)
    Ad = adjoint(p[:9].reshape(3, 3), p[9:])
    Vs = V @ Ad.T
    dVs = np.einsum("bij,bj->bi", curlywedge(Vs) @ Ad, V) + dV @ Ad.T
    return Vs, dVs


def regressor_batched(V, dV):
    """Batched reference dynamics.py:215-249 -> (B, 6, 10)."""
    V, dV = np.asarray(V, float), np.asarray(dV, float)
    B = V.shape[0]
    v, w, dv, dw = V[:, :3], V[:, 3:], dV[:, :3], dV[:, 3:]
    Ww, Wdw = skew(w), skew(dw)

    def bullet(a):
        M = np.zeros((B, 3, 6))
        M[:, 0, 0], M[:, 0, 3], M[:, 0, 5] = a[:, 0], a[:, 1], a[:, 2]
        M[:, 1, 1], M[:, 1, 3], M[:, 1, 4] = a[:, 1], a[:, 0], a[:, 2]
        M[:, 2, 2], M[:, 2, 4], M[:, 2, 5] = a[:, 2], a[:, 1], a[:, 0]
        return M

    x = dv + np.einsum("bij,bj->bi", Ww, v)
    Y = np.zeros((B, 6, 10))
    Y[:, :3, 0] = x
    Y[:, :3, 1:4] = Wdw + Ww @ Ww
    Y[:, 3:, 1:4] = -skew(x)
    Y[:, 3:, 4:] = bullet(dw) + Ww @ bullet(w)
    return Y


def gram_pack(Y, f):
    """[Y^T Y (100, row-major) | Y^T f (10) | f^T f | n_samples] -- the 112-double pack of SURVEY.md 8(b/e)."""
    Yf = Y.reshape(-1, 10)
    ff = np.asarray(f, float).reshape(-1)
    return np.concatenate([(Yf.T @ Yf).reshape(100), Yf.T @ ff, [ff @ ff], [float(Y.shape[0])]])


def identify_lstsq(Y, f):
    """reference loggers/loggers.py:127-129: lstsq over the stacked (6F, 10) regressor."""
    phi, *_ = np.linalg.lstsq(Y.reshape(-1, 10), np.asarray(f, float).reshape(-1), rcond=None)
    return phi
