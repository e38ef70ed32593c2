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
    """`consts` is the dict of Newton-Euler constants, or any callable q -> (M(q), U(q)) (e.g. `mjcf_mass_matrix_and_potential`)."""
    q, qd, qdd = (np.asarray(a, float) for a in (q, qd, qdd))
    nj = len(q)
    if callable(consts):
        mass_matrix_and_potential = consts  # noqa: F811 -- same contract, other source of M and U
        consts = None
    else:
        mass_matrix_and_potential = globals()["mass_matrix_and_potential"]
    M, _ = mass_matrix_and_potential(consts, q) if consts is not None else mass_matrix_and_potential(q)
    dM = np.zeros((nj, nj, nj))  # dM[k] = dM/dq_k
    dU = np.zeros(nj)
    for k in range(nj):
        e = np.zeros(nj)
        e[k] = h
        Mp, Up = mass_matrix_and_potential(consts, q + e) if consts is not None else mass_matrix_and_potential(q + e)
        Mm, Um = mass_matrix_and_potential(consts, q - e) if consts is not None else mass_matrix_and_potential(q - e)
        dM[k] = (Mp - Mm) / (2 * h)
        dU[k] = (Up - Um) / (2 * h)
    Mdot = np.einsum("kij,k->ij", dM, qd)
    quad = 0.5 * np.einsum("kij,i,j->k", dM, qd, qd)
    return M @ qdd + Mdot @ qd - quad + dU


# ---------------------------------------------------------------------------------------------------------------------
# M(q), U(q) straight from an MJCF text (MuJoCo's documented body / joint / inertial semantics), bypassing every
# Newton-Euler constant: used to check `rigid_body_manipulation_b200.mjcf_export.to_mjcf` without MuJoCo.
# ---------------------------------------------------------------------------------------------------------------------
def _quat_R(q):
    w, x, y, z = np.asarray(q, float) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def mjcf_mass_matrix_and_potential(xml_text):
    """Returns f(q) -> (M, U) for a tree given as MJCF text with quat-only frames (what `to_mjcf` writes).

    MuJoCo semantics: a body frame is (pos, quat) in its parent; its joint moves the body relative to the parent about / along
    `axis` (body coordinates) through the anchor `pos` (body coordinates); `inertial` is (pos, quat, mass, diaginertia) in the body."""
    import xml.etree.ElementTree as ET

    root = ET.fromstring(xml_text)
    fl = lambda s, d: np.array([float(x) for x in s.split()]) if s is not None else np.array(d, float)  # noqa: E731
    g = fl(root.find("option").get("gravity"), [0, 0, -9.81])
    bodies = []  # (parent index, pos, R, joint or None, inertial or None)

    def walk(elem, parent):
        for b in elem.findall("body"):
            j, it = b.find("joint"), b.find("inertial")
            jt = None if j is None else (j.get("type", "hinge"), fl(j.get("axis"), [0, 0, 1]), fl(j.get("pos"), [0, 0, 0]))
            inert = None if it is None else (float(it.get("mass")), fl(it.get("pos"), [0, 0, 0]), _quat_R(fl(it.get("quat"), [1, 0, 0, 0])),
                                             fl(it.get("diaginertia"), [0, 0, 0]))
            bodies.append((parent, fl(b.get("pos"), [0, 0, 0]), _quat_R(fl(b.get("quat"), [1, 0, 0, 0])), jt, inert))
            walk(b, len(bodies) - 1)

    walk(root.find("worldbody"), -1)
    nq = sum(1 for b in bodies if b[3] is not None)

    def f(q):
        frames, joints, k = [], [], 0  # joints: (kind, world axis, world anchor, body index)
        for parent, pos, R, jt, _ in bodies:
            Rp, pp = (np.eye(3), np.zeros(3)) if parent < 0 else frames[parent]
            Rb, pb = Rp @ R, Rp @ pos + pp
            if jt is not None:
                kind, axis, anchor = jt
                a_w, anc_w = Rb @ axis, Rb @ anchor + pb
                if kind == "slide":
                    pb = pb + a_w * q[k]
                    anc_w = anc_w + a_w * q[k]
                else:
                    Rj = expm(_hat(a_w) * q[k])
                    Rb, pb = Rj @ Rb, anc_w + Rj @ (pb - anc_w)
                joints.append((kind, a_w, anc_w, len(frames)))
                k += 1
            frames.append((Rb, pb))
        # ancestors
        M, U = np.zeros((nq, nq)), 0.0
        for bi, (parent, _, _, _, inert) in enumerate(bodies):
            if inert is None:
                continue
            m, ipos, iR, diag = inert
            Rb, pb = frames[bi]
            pc, Rc = Rb @ ipos + pb, Rb @ iR
            anc, a = set(), bi
            while a >= 0:
                anc.add(a)
                a = bodies[a][0]
            Jv, Jw = np.zeros((3, nq)), np.zeros((3, nq))
            for col, (kind, a_w, anc_w, jb) in enumerate(joints):
                if jb not in anc:
                    continue
                if kind == "slide":
                    Jv[:, col] = a_w
                else:
                    Jw[:, col] = a_w
                    Jv[:, col] = np.cross(a_w, pc - anc_w)
            M += m * Jv.T @ Jv + Jw.T @ (Rc @ np.diag(diag) @ Rc.T) @ Jw
            U -= m * g @ pc
        return M, U

    return f
