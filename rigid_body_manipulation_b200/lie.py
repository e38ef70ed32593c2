"""SE(3) / SO(3) value types for the drop-in API surface.

The reference's public functions take and return `liegroups` objects (reference dynamics/dynamics.py:6,
transformations/transformations.py:4): `hposes_body_parent` is a list of SE3, `inverse()` returns a list of SE3,
`compose()` builds them.  When the real `liegroups` package is importable its classes are used unchanged, so objects
flow between the caller's code and this package.  Where it is absent (this image), the minimal host-side stand-ins
below provide the same attribute / method surface (`rot`, `trans`, `as_matrix`, `inv`, `dot`, `adjoint`, `exp`,
`wedge`, `curlywedge`, `identity`, `from_quaternion`, `from_matrix`, `from_rpy`) with liegroups' conventions:
twist order [translation; rotation], wxyz quaternions, constructors keep references to their arrays.

These are containers plus 3x3 / 4x4 setup algebra for a handful of constant poses; the batched hot path never goes
through them (it runs in librbm_b200.so).
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - liegroups is absent from the build image
    from liegroups.numpy import SE3, SO3  # type: ignore

    HAVE_LIEGROUPS = True
except ImportError:
    HAVE_LIEGROUPS = False

    def _hat(phi):
        """(3,) -> (3,3) or (N,3) -> (N,3,3) cross-product matrix."""
        phi = np.atleast_2d(phi)
        if phi.shape[1] != 3:
            raise ValueError("phi must have shape (3,) or (N,3)")
        out = np.zeros((phi.shape[0], 3, 3))
        out[:, 2, 1], out[:, 0, 2], out[:, 1, 0] = phi[:, 0], phi[:, 1], phi[:, 2]
        out[:, 1, 2], out[:, 2, 0], out[:, 0, 1] = -phi[:, 0], -phi[:, 1], -phi[:, 2]
        return np.squeeze(out)

    class SO3:
        dim, dof = 3, 3

        def __init__(self, mat):
            self.mat = mat

        @classmethod
        def identity(cls):
            return cls(np.eye(3))

        @classmethod
        def is_valid_matrix(cls, mat):
            return mat.shape == (3, 3) and np.isclose(np.linalg.det(mat), 1.0) and np.allclose(mat.T @ mat, np.eye(3))

        @classmethod
        def from_matrix(cls, mat, normalize=False):
            ok = cls.is_valid_matrix(mat)
            if not ok and not normalize:
                raise ValueError("Invalid rotation matrix. Use normalize=True to handle rounding errors.")
            out = cls(mat)
            if not ok:
                U, _, Vt = np.linalg.svd(mat)
                out.mat = U @ np.diag([1.0, 1.0, np.linalg.det(U) * np.linalg.det(Vt)]) @ Vt
            return out

        @classmethod
        def from_quaternion(cls, quat, ordering="wxyz"):
            quat = np.asarray(quat, dtype=float)
            if not np.isclose(np.linalg.norm(quat), 1.0):
                raise ValueError("Quaternion must be unit length")
            if ordering == "wxyz":
                w, x, y, z = quat
            elif ordering == "xyzw":
                x, y, z, w = quat
            else:
                raise ValueError(f"Valid orderings are 'xyzw' and 'wxyz'. Got '{ordering}'.")
            return cls(np.array([
                [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (w * y + x * z)],
                [2 * (w * z + x * y), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                [2 * (x * z - w * y), 2 * (w * x + y * z), 1 - 2 * (x * x + y * y)],
            ]))

        @classmethod
        def from_rpy(cls, roll, pitch, yaw):
            cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
            Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
            Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
            Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
            return cls(Rz @ Ry @ Rx)

        wedge = staticmethod(_hat)

        @classmethod
        def exp(cls, phi):
            phi = np.asarray(phi, dtype=float)
            if phi.shape != (3,):
                raise ValueError("phi must have length 3")
            th = np.linalg.norm(phi)
            if np.isclose(th, 0.0):
                return cls(np.eye(3) + _hat(phi))
            a = phi / th
            return cls(np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(a, a) + np.sin(th) * _hat(a))

        @classmethod
        def left_jacobian(cls, phi):
            phi = np.asarray(phi, dtype=float)
            if phi.shape != (3,):
                raise ValueError("phi must have length 3")
            th = np.linalg.norm(phi)
            if np.isclose(th, 0.0):
                return np.eye(3) + 0.5 * _hat(phi)
            a = phi / th
            s = np.sin(th) / th
            return s * np.eye(3) + (1 - s) * np.outer(a, a) + ((1 - np.cos(th)) / th) * _hat(a)

        def as_matrix(self):
            return self.mat

        def inv(self):
            return SO3(self.mat.T)

        def dot(self, other):
            if isinstance(other, SO3):
                return SO3(self.mat @ other.mat)
            other = np.atleast_2d(other)
            if other.shape[1] != 3:
                raise ValueError("Vector must have shape (3,) or (N,3)")
            return np.squeeze((self.mat @ other.T).T)

        def __repr__(self):
            return f"<SO3>\n{self.mat}"

    class SE3:
        dim, dof = 4, 6
        RotationType = SO3

        def __init__(self, rot, trans):
            self.rot = rot
            self.trans = trans

        @classmethod
        def identity(cls):
            return cls(SO3.identity(), np.zeros(3))

        @classmethod
        def from_matrix(cls, mat, normalize=False):
            mat = np.asarray(mat)
            if mat.shape != (4, 4) or not np.array_equal(mat[3], [0, 0, 0, 1]):
                if not normalize:
                    raise ValueError("Invalid transformation matrix. Use normalize=True to handle rounding errors.")
            return cls(SO3.from_matrix(mat[:3, :3], normalize), mat[:3, 3])

        @classmethod
        def wedge(cls, xi):
            xi = np.atleast_2d(xi)
            if xi.shape[1] != 6:
                raise ValueError("xi must have shape (6,) or (N,6)")
            out = np.zeros((xi.shape[0], 4, 4))
            out[:, :3, :3] = _hat(xi[:, 3:]).reshape(-1, 3, 3)
            out[:, :3, 3] = xi[:, :3]
            return np.squeeze(out)

        @classmethod
        def curlywedge(cls, xi):
            xi = np.atleast_2d(xi)
            if xi.shape[1] != 6:
                raise ValueError("xi must have shape (6,) or (N,6)")
            out = np.zeros((xi.shape[0], 6, 6))
            out[:, :3, :3] = out[:, 3:, 3:] = _hat(xi[:, 3:]).reshape(-1, 3, 3)
            out[:, :3, 3:] = _hat(xi[:, :3]).reshape(-1, 3, 3)
            return np.squeeze(out)

        @classmethod
        def exp(cls, xi):
            xi = np.asarray(xi, dtype=float)
            if xi.shape != (6,):
                raise ValueError("xi must have length 6")
            return cls(SO3.exp(xi[3:]), SO3.left_jacobian(xi[3:]) @ xi[:3])

        def as_matrix(self):
            out = np.eye(4)
            out[:3, :3] = self.rot.as_matrix()
            out[:3, 3] = self.trans
            return out

        def adjoint(self):
            R = self.rot.as_matrix()
            out = np.zeros((6, 6))
            out[:3, :3] = out[3:, 3:] = R
            out[:3, 3:] = _hat(self.trans) @ R
            return out

        def inv(self):
            Rt = self.rot.inv()
            return SE3(Rt, -(Rt.dot(self.trans)))

        def dot(self, other):
            if isinstance(other, SE3):
                return SE3(self.rot.dot(other.rot), self.rot.dot(other.trans) + self.trans)
            other = np.atleast_2d(other)
            if other.shape[1] == 3:
                return np.squeeze(self.rot.dot(other) + self.trans)
            if other.shape[1] == 4:
                return np.squeeze(self.as_matrix() @ other.T).T
            raise ValueError("Vector must have shape (3,), (4,), (N,3) or (N,4)")

        def __repr__(self):
            return f"<SE3>\n{self.as_matrix()}"


def is_se3(obj) -> bool:
    """Duck-typed SE3 test (accepts liegroups objects, ours, and the oracle shim's)."""
    return hasattr(obj, "rot") and hasattr(obj, "trans") and hasattr(obj, "adjoint")


def se3_from_Rt(Rt) -> "SE3":
    Rt = np.asarray(Rt, dtype=float)
    return SE3(SO3(Rt[:9].reshape(3, 3).copy()), Rt[9:12].copy())
