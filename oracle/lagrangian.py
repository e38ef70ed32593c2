"""ORACLE / TEST INFRASTRUCTURE ONLY -- an INDEPENDENT derivation of the joint forces, used to cross-check the
Newton-Euler restatements (and, through them, the CUDA kernels) in place of the MuJoCo `mj_inverse` comparison that the
north star asks for but that cannot run here (MuJoCo is absent from the image).

Euler-Lagrange in WORLD coordinates, sharing no recursion with reference dynamics/dynamics.py:109-157:

    tau = M(q) qdd + Mdot(q, qd) qd - 1/2 d(qd^T M qd)/dq + dU/dq

  * forward kinematics by chaining 4x4 matrices, joint motion through scipy.linalg.expm of the se(3) matrix
    (not the closed-form Rodrigues of the liegroups shim)
  * M(q) = sum_i m_i Jv_i^T Jv_i + Jw_i^T (R_i Ic_i R_i^T) Jw_i with geometric Jacobians of each link's centre of mass
  * U(q) = -sum_i m_i g . p_ci ; derivatives of M and U by central differences (step 1e-6 -> ~1e-9 relative accuracy)

Link mass, centre of mass and central inertia are read back from the rigid-body form of the spatial inertias
G = [[m 1, -m[c]x], [m[c]x, Ic + m(|c|^2 1 - c c^T)]], so this only applies to physical inertias (the reference's models).
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import expm


def _hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


def _se3_hat(xi):
    X = np.zeros((4, 4))
    X[:3, :3] = _hat(xi[3:])
    X[:3, 3] = xi[:3]
    return X


def _T_of(Rt):
    T = np.eye(4)
    T[:3, :3] = np.asarray(Rt[:9]).reshape(3, 3)
    T[:3, 3] = Rt[9:]
    return T


def _link_inertials(G):
    m = G[0, 0]
    h = np.array([G[5, 1], G[3, 2], G[4, 0]])
    c = h / m
    Ic = G[3:, 3:] - m * (c @ c * np.eye(3) - np.outer(c, c))
    return m, c, Ic


def _kinematics(consts, q):
    """World poses T_{0,i} of every joint frame: T_{0,i} = T_{0,i-1} M_i^-1 exp([S_i] q_i)."""
    nj = len(consts["uscrews"])
    T = np.eye(4)
    out = []
    for i in range(nj):
        T = T @ np.linalg.inv(_T_of(consts["hposes_Rt"][i + 1])) @ expm(_se3_hat(consts["uscrews"][i]) * q[i])
        out.append(T.copy())
    return out


def mass_matrix_and_potential(consts, q):
    nj = len(consts["uscrews"])
    Ts = _kinematics(consts, q)
    g_w = -np.asarray(consts["dtwist_0"][:3], float)  # dtwist_0 = -[gravity, 0, 0, 0]  (core/simulate.py:149)
    # world-frame screw of joint j: Ad(T_{0,j}) S_j  -> (v_s, w_s)
    axes = []
    for j in range(nj):
        R, p = Ts[j][:3, :3], Ts[j][:3, 3]
        Sv, Sw = consts["uscrews"][j][:3], consts["uscrews"][j][3:]
        w = R @ Sw
        v = R @ Sv + np.cross(p, w)
        axes.append((v, w))
    M = np.zeros((nj, nj))
    U = 0.0
    for i in range(nj):
        m, c, Ic = _link_inertials(np.asarray(consts["simats"][i + 1], float))
        R, p = Ts[i][:3, :3], Ts[i][:3, 3]
        pc = R @ c + p
        Jv, Jw = np.zeros((3, nj)), np.zeros((3, nj))
        for j in range(i + 1):
            v, w = axes[j]
            Jv[:, j] = v + np.cross(w, pc)
            Jw[:, j] = w
        M += m * Jv.T @ Jv + Jw.T @ (R @ Ic @ R.T) @ Jw
        U -= m * g_w @ pc
    return M, U


def lagrangian_tau(consts, q, qd, qdd, h=1e-6):
    q, qd, qdd = (np.asarray(a, float) for a in (q, qd, qdd))
    nj = len(q)
    M, _ = mass_matrix_and_potential(consts, q)
    dM = np.zeros((nj, nj, nj))  # dM[k] = dM/dq_k
    dU = np.zeros(nj)
    for k in range(nj):
        e = np.zeros(nj)
        e[k] = h
        Mp, Up = mass_matrix_and_potential(consts, q + e)
        Mm, Um = mass_matrix_and_potential(consts, q - e)
        dM[k] = (Mp - Mm) / (2 * h)
        dU[k] = (Up - Um) / (2 * h)
    Mdot = np.einsum("kij,k->ij", dM, qd)
    quad = 0.5 * np.einsum("kij,i,j->k", dM, qd, qd)
    return M @ qdd + Mdot @ qd - quad + dU
