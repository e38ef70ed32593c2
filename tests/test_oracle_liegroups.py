"""Pins oracle/shims/liegroups (the restated third-party dependency, absent from /root/reference)
against INDEPENDENT implementations: scipy.linalg.expm, a power series of the SO(3) left Jacobian,
scipy.spatial.transform.Rotation, and the reference's own scratch identity check
(reference test_adjoint_inv_transpose.py:8-27)."""
import math

import numpy as np
import pytest
from scipy.linalg import expm
from scipy.spatial.transform import Rotation

from oracle.shims.liegroups import SE3, SO3
from oracle.shims.liegroups.numpy import SE3 as SE3n


def test_import_paths_agree():
    # reference uses both `from liegroups import SE3` and `from liegroups.numpy import SE3`
    assert SE3 is SE3n


@pytest.mark.parametrize("seed", range(5))
def test_so3_exp_matches_expm(seed):
    rng = np.random.default_rng(seed)
    phi = rng.standard_normal(3) * rng.choice([1e-3, 1.0, 10.0])
    assert np.allclose(SO3.exp(phi).as_matrix(), expm(SO3.wedge(phi)), atol=1e-12)


@pytest.mark.parametrize("seed", range(5))
def test_se3_exp_matches_expm(seed):
    rng = np.random.default_rng(100 + seed)
    xi = rng.standard_normal(6) * rng.choice([1e-2, 1.0, 5.0])
    assert np.allclose(SE3.exp(xi).as_matrix(), expm(SE3.wedge(xi)), atol=1e-11)


def test_small_angle_branch_is_first_order():
    phi = np.array([3e-9, -4e-9, 1e-9])  # norm < isclose atol 1e-8
    assert np.array_equal(SO3.exp(phi).as_matrix(), np.eye(3) + SO3.wedge(phi))
    assert np.array_equal(SO3.left_jacobian(phi), np.eye(3) + 0.5 * SO3.wedge(phi))
    phi = np.zeros(3)
    assert np.array_equal(SE3.exp(np.array([1.0, 2, 3, 0, 0, 0])).trans, np.array([1.0, 2, 3]))


def test_left_jacobian_series():
    phi = np.array([0.3, -0.5, 0.7])
    W = SO3.wedge(phi)
    J = sum(np.linalg.matrix_power(W, n) / math.factorial(n + 1) for n in range(30))
    assert np.allclose(SO3.left_jacobian(phi), J, atol=1e-14)


def test_quaternion_and_rpy_against_scipy():
    rng = np.random.default_rng(1)
    for _ in range(10):
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        R = Rotation.from_quat([q[1], q[2], q[3], q[0]]).as_matrix()  # scipy is xyzw
        assert np.allclose(SO3.from_quaternion(q).as_matrix(), R, atol=1e-14)
    r, p, y = 0.1, -0.7, 2.0
    assert np.allclose(SO3.from_rpy(r, p, y).as_matrix(), Rotation.from_euler("xyz", [r, p, y]).as_matrix(), atol=1e-14)
    with pytest.raises(ValueError):
        SO3.from_quaternion(np.array([1.0, 1.0, 0, 0]))
    with pytest.raises(ValueError):
        SO3.from_matrix(np.diag([1.0, 1.0, -1.0]))


def test_reference_adjoint_identity_check():
    # reference test_adjoint_inv_transpose.py: prints allclose(Ad^T, Ad(T^-1)) [False] and allclose(Ad(T^-1), pinv(Ad(T))) [True]
    pose = SE3(SO3.from_rpy(10 / 180 * math.pi, 20 / 180 * math.pi, 40 / 180 * math.pi), np.array([1, 2, 3]))
    Ad = pose.adjoint()
    assert not np.allclose(Ad.T, pose.inv().adjoint())
    assert np.allclose(pose.inv().adjoint(), np.linalg.pinv(Ad))


def test_group_identities():
    rng = np.random.default_rng(2)
    T1 = SE3.exp(rng.standard_normal(6))
    T2 = SE3.exp(rng.standard_normal(6))
    assert np.allclose(T1.dot(T2).adjoint(), T1.adjoint() @ T2.adjoint(), atol=1e-12)
    assert np.allclose(T1.dot(T1.inv()).as_matrix(), np.eye(4), atol=1e-12)
    V = rng.standard_normal(6)
    assert np.allclose(SE3.curlywedge(V) @ V, 0, atol=1e-14)
    xi = rng.standard_normal(6)
    assert np.allclose(SE3.exp(xi).dot(SE3.exp(-xi)).as_matrix(), np.eye(4), atol=1e-12)
    # adjoint acts on twists like conjugation acts on se(3) matrices
    lhs = SE3.wedge(T1.adjoint() @ V)
    rhs = T1.as_matrix() @ SE3.wedge(V) @ T1.inv().as_matrix()
    assert np.allclose(lhs, rhs, atol=1e-12)
    # points and homogeneous points
    p = rng.standard_normal(3)
    assert np.allclose(T1.dot(p), T1.as_matrix()[:3, :3] @ p + T1.trans)
    assert np.allclose(T1.dot(np.append(p, 1.0))[:3], T1.dot(p))


def test_constructors_keep_references():
    t = np.zeros(3)
    T = SE3(SO3.identity(), t)
    t[0] = 5.0
    assert T.trans[0] == 5.0  # reference transformations/poses.py:16-19 relies on this aliasing
