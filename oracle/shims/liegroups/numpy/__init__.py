"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product package.

numpy restatement of the slice of `liegroups` 1.1.0 that the reference's hot path
uses.  The reference pins `liegroups @ git+https://github.com/utiasSTARS/liegroups.git
@refs/pull/10/head` (reference `pyproject.toml:15`, `requirements.txt:2`,
`uv.lock:864-866`, commit dad56d8ff21f553d92a7343d5ac2ce0f45cb007b).  That source is
NOT under /root/reference and cannot be fetched (no network), so the published
algorithm of `liegroups.numpy.{SO3,SE3}` is restated here:

  * twist / se(3) vector order is [rho (translation, 3); phi (rotation, 3)]
  * SO3.exp      : Rodrigues  c*I + (1-c)*a a^T + s*[a]x ; `np.isclose(angle, 0.)`
                   branch returns the first-order  I + [phi]x
  * SO3.left_jacobian : (s/t) I + (1 - s/t) a a^T + ((1-c)/t) [a]x ; small-angle
                   branch  I + 0.5 [phi]x
  * SE3.exp      : (SO3.exp(phi), J_l(phi) rho)
  * SE3.adjoint  : [[R, [t]x R], [0, R]]
  * SE3.curlywedge : [[ [phi]x, [rho]x ], [0, [phi]x ]]
  * SE3.wedge    : 4x4 [[ [phi]x, rho ], [0, 0]]
  * constructors keep REFERENCES to the arrays they are given (no copy) -- the
    reference relies on this for its "dynamic" poses (`transformations/poses.py:16-19`).

Call sites in the reference that fix the required surface:
`dynamics/dynamics.py:88,102,117,126,128,130,143,145,180,210,237-244,253,261`,
`transformations/transformations.py:11-12,18-19`, `transformations/poses.py:20,23`,
`core/simulate.py:140-146,202-209,232-237`, `core/core.py:168`,
`test_adjoint_inv_transpose.py:8-22`.

The shim is pinned (tests/test_oracle_liegroups.py) against independent
implementations: `scipy.linalg.expm` for both exponentials, a series expansion of the
left Jacobian, `scipy.spatial.transform.Rotation` for quaternion / rpy conversions and
the reference's own adjoint identity check (`test_adjoint_inv_transpose.py`).
"""
import numpy as _np


class SO3:
    """Rotation matrix in SO(3) (liegroups.numpy.so3.SO3Matrix)."""

    dim = 3
    dof = 3

    def __init__(self, mat):
        self.mat = mat  # reference kept, not copied

    # -- construction ---------------------------------------------------------
    @classmethod
    def identity(cls):
        return cls(_np.identity(cls.dim))

    @classmethod
    def is_valid_matrix(cls, mat):
        return (
            mat.shape == (cls.dim, cls.dim)
            and _np.isclose(_np.linalg.det(mat), 1.0)
            and _np.allclose(mat.T.dot(mat), _np.identity(cls.dim))
        )

    @classmethod
    def from_matrix(cls, mat, normalize=False):
        mat_is_valid = cls.is_valid_matrix(mat)
        if mat_is_valid or normalize:
            result = cls(mat)
            if not mat_is_valid and normalize:
                result.normalize()
        else:
            raise ValueError("Invalid rotation matrix. Use normalize=True to handle rounding errors.")
        return result

    def normalize(self):
        U, _, V = _np.linalg.svd(self.mat, full_matrices=False)
        S = _np.identity(self.dim)
        S[self.dim - 1, self.dim - 1] = _np.linalg.det(U) * _np.linalg.det(V)
        self.mat = U.dot(S).dot(V)

    @classmethod
    def from_quaternion(cls, quat, ordering="wxyz"):
        if not _np.isclose(_np.linalg.norm(quat), 1.0):
            raise ValueError("Quaternion must be unit length")
        if ordering == "xyzw":
            qx, qy, qz, qw = quat
        elif ordering == "wxyz":
            qw, qx, qy, qz = quat
        else:
            raise ValueError("Valid orderings are 'xyzw' and 'wxyz'. Got '{}'.".format(ordering))
        qw2 = qw * qw
        qx2 = qx * qx
        qy2 = qy * qy
        qz2 = qz * qz
        R = _np.array(
            [
                [1.0 - 2.0 * (qy2 + qz2), 2.0 * (qx * qy - qw * qz), 2.0 * (qw * qy + qx * qz)],
                [2.0 * (qw * qz + qx * qy), 1.0 - 2.0 * (qx2 + qz2), 2.0 * (qy * qz - qw * qx)],
                [2.0 * (qx * qz - qw * qy), 2.0 * (qw * qx + qy * qz), 1.0 - 2.0 * (qx2 + qy2)],
            ]
        )
        del qw2
        return cls(R)

    @classmethod
    def rotx(cls, angle_in_radians):
        c = _np.cos(angle_in_radians)
        s = _np.sin(angle_in_radians)
        return cls(_np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]]))

    @classmethod
    def roty(cls, angle_in_radians):
        c = _np.cos(angle_in_radians)
        s = _np.sin(angle_in_radians)
        return cls(_np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]]))

    @classmethod
    def rotz(cls, angle_in_radians):
        c = _np.cos(angle_in_radians)
        s = _np.sin(angle_in_radians)
        return cls(_np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]]))

    @classmethod
    def from_rpy(cls, roll, pitch, yaw):
        return cls.rotz(yaw).dot(cls.roty(pitch).dot(cls.rotx(roll)))

    # -- lie algebra ----------------------------------------------------------
    @classmethod
    def wedge(cls, phi):
        phi = _np.atleast_2d(phi)
        if phi.shape[1] != cls.dof:
            raise ValueError("phi must have shape ({},) or (N,{})".format(cls.dof, cls.dof))
        Phi = _np.zeros([phi.shape[0], cls.dim, cls.dim])
        Phi[:, 0, 1] = -phi[:, 2]
        Phi[:, 1, 0] = phi[:, 2]
        Phi[:, 0, 2] = phi[:, 1]
        Phi[:, 2, 0] = -phi[:, 1]
        Phi[:, 1, 2] = -phi[:, 0]
        Phi[:, 2, 1] = phi[:, 0]
        return _np.squeeze(Phi)

    @classmethod
    def exp(cls, phi):
        if len(phi) != cls.dof:
            raise ValueError("phi must have length 3")
        angle = _np.linalg.norm(phi)
        if _np.isclose(angle, 0.0):
            return cls(_np.identity(cls.dim) + cls.wedge(phi))
        axis = phi / angle
        s = _np.sin(angle)
        c = _np.cos(angle)
        return cls(c * _np.identity(cls.dim) + (1 - c) * _np.outer(axis, axis) + s * cls.wedge(axis))

    @classmethod
    def left_jacobian(cls, phi):
        if len(phi) != cls.dof:
            raise ValueError("phi must have length 3")
        angle = _np.linalg.norm(phi)
        if _np.isclose(angle, 0.0):
            return _np.identity(cls.dof) + 0.5 * cls.wedge(phi)
        axis = phi / angle
        s = _np.sin(angle)
        c = _np.cos(angle)
        return (
            (s / angle) * _np.identity(cls.dof)
            + (1 - s / angle) * _np.outer(axis, axis)
            + ((1 - c) / angle) * cls.wedge(axis)
        )

    # -- group ops ------------------------------------------------------------
    def as_matrix(self):
        return self.mat

    def inv(self):
        return self.__class__(self.mat.T)

    def dot(self, other):
        if isinstance(other, self.__class__):
            return self.__class__(_np.dot(self.mat, other.mat))
        other = _np.atleast_2d(other)
        if other.shape[1] == self.dim:
            return _np.squeeze(_np.dot(self.mat, other.T).T)
        raise ValueError("Vector must have shape ({},) or (N,{})".format(self.dim, self.dim))

    def __repr__(self):
        return "<{}.{}>\n{}".format(self.__class__.__module__, self.__class__.__name__, self.as_matrix())


class SE3:
    """Homogeneous transform in SE(3) (liegroups.numpy.se3.SE3Matrix)."""

    dim = 4
    dof = 6
    RotationType = SO3

    def __init__(self, rot, trans):
        self.rot = rot  # references kept, not copied
        self.trans = trans

    @classmethod
    def identity(cls):
        return cls.from_matrix(_np.identity(cls.dim))

    @classmethod
    def is_valid_matrix(cls, mat):
        bottom_row = _np.append(_np.zeros(cls.dim - 1), 1.0)
        return (
            mat.shape == (cls.dim, cls.dim)
            and _np.array_equal(mat[cls.dim - 1, :], bottom_row)
            and cls.RotationType.is_valid_matrix(mat[0 : cls.dim - 1, 0 : cls.dim - 1])
        )

    @classmethod
    def from_matrix(cls, mat, normalize=False):
        mat_is_valid = cls.is_valid_matrix(mat)
        if mat_is_valid or normalize:
            result = cls(cls.RotationType(mat[0 : cls.dim - 1, 0 : cls.dim - 1]), mat[0 : cls.dim - 1, cls.dim - 1])
            if not mat_is_valid and normalize:
                result.rot.normalize()
        else:
            raise ValueError("Invalid transformation matrix. Use normalize=True to handle rounding errors.")
        return result

    # -- lie algebra ----------------------------------------------------------
    @classmethod
    def wedge(cls, xi):
        xi = _np.atleast_2d(xi)
        if xi.shape[1] != cls.dof:
            raise ValueError("xi must have shape ({},) or (N,{})".format(cls.dof, cls.dof))
        Xi = _np.zeros([xi.shape[0], cls.dim, cls.dim])
        Xi[:, 0:3, 0:3] = SO3.wedge(xi[:, 3:6])
        Xi[:, 0:3, 3] = xi[:, 0:3]
        return _np.squeeze(Xi)

    @classmethod
    def curlywedge(cls, xi):
        xi = _np.atleast_2d(xi)
        if xi.shape[1] != cls.dof:
            raise ValueError("xi must have shape ({},) or (N,{})".format(cls.dof, cls.dof))
        Psi = _np.zeros([xi.shape[0], cls.dof, cls.dof])
        Psi[:, 0:3, 0:3] = SO3.wedge(xi[:, 3:6])
        Psi[:, 0:3, 3:6] = SO3.wedge(xi[:, 0:3])
        Psi[:, 3:6, 3:6] = Psi[:, 0:3, 0:3]
        return _np.squeeze(Psi)

    @classmethod
    def exp(cls, xi):
        if len(xi) != cls.dof:
            raise ValueError("xi must have length 6")
        rho = xi[0:3]
        phi = xi[3:6]
        return cls(SO3.exp(phi), SO3.left_jacobian(phi).dot(rho))

    # -- group ops ------------------------------------------------------------
    def as_matrix(self):
        R = self.rot.as_matrix()
        t = _np.reshape(self.trans, (self.dim - 1, 1))
        bottom_row = _np.append(_np.zeros(self.dim - 1), 1.0)
        return _np.vstack([_np.hstack([R, t]), bottom_row])

    def adjoint(self):
        rotmat = self.rot.as_matrix()
        return _np.vstack(
            [
                _np.hstack([rotmat, SO3.wedge(self.trans).dot(rotmat)]),
                _np.hstack([_np.zeros((3, 3)), rotmat]),
            ]
        )

    def inv(self):
        inv_rot = self.rot.inv()
        inv_trans = -(inv_rot.dot(self.trans))
        return self.__class__(inv_rot, inv_trans)

    def dot(self, other):
        if isinstance(other, self.__class__):
            return self.__class__(self.rot.dot(other.rot), self.rot.dot(other.trans) + self.trans)
        other = _np.atleast_2d(other)
        if other.shape[1] == self.dim - 1:
            return _np.squeeze(self.rot.dot(other) + self.trans)
        if other.shape[1] == self.dim:
            return _np.squeeze(self.as_matrix().dot(other.T)).T
        raise ValueError(
            "Vector must have shape ({},), ({},), (N,{}) or (N,{})".format(
                self.dim - 1, self.dim, self.dim - 1, self.dim
            )
        )

    def __repr__(self):
        return "<{}.{}>\n{}".format(self.__class__.__module__, self.__class__.__name__, self.as_matrix())


SO3Matrix = SO3
SE3Matrix = SE3
