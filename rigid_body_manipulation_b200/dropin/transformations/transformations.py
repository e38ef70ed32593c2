"""Drop-in for reference transformations/transformations.py:8-56 (same names, arguments, return kinds, errors).

`compose` converts and validates every row on the GPU in one launch (rbm_compose_f64); the returned objects are
liegroups-style SE3 values (rigid_body_manipulation_b200.lie).  As in the reference, translation rows -- and rotation
rows given as matrices -- are kept as VIEWS of the caller's arrays, which is what makes the reference's world poses
"dynamic" (they alias live MuJoCo buffers, reference transformations/poses.py:16-19)."""
from typing import Optional, Union

import numpy as np
from numpy.typing import NDArray

from rigid_body_manipulation_b200 import engine as _engine
from rigid_body_manipulation_b200.lie import SE3, SO3

__all__ = ["tq2se3", "tr2se3", "compose", "homogenize"]

_ERR = {1: "Quaternion must be unit length", 2: "Invalid rotation matrix. Use normalize=True to handle rounding errors."}


def _compose_rows(trans, rot):
    """rows of one kind (all quaternions or all matrices) -> list[SE3]; raises ValueError like liegroups does."""
    n, width = rot.shape
    Rt, status = _engine.compose_poses(np.ascontiguousarray(trans, dtype=np.float64), np.ascontiguousarray(rot, dtype=np.float64))
    status = status.cpu().numpy()
    if status.any():
        raise ValueError(_ERR[int(status[np.nonzero(status)[0][0]])])
    Rt = Rt.cpu().numpy()
    out = []
    for k in range(n):
        R = rot[k].reshape(3, 3) if width == 9 else Rt[k, :9].reshape(3, 3)
        out.append(SE3(SO3(R), trans[k]))
    return out


def tq2se3(t, q) -> SE3:
    return _compose_rows(np.asarray(t)[None], np.asarray(q, dtype=float)[None])[0]


def tr2se3(t, r) -> SE3:
    r = np.asarray(r)
    if r.shape != (3, 3):
        raise ValueError("Invalid rotation matrix. Use normalize=True to handle rounding errors.")
    return _compose_rows(np.asarray(t)[None], r.reshape(1, 9))[0]


def compose(trans: NDArray, rot: Optional[NDArray] = None) -> Union[SE3, list]:
    """Translations (3,) / (n, 3) with rotations given as wxyz quaternions (4,) / (n, 4) or flattened matrices (9,) / (n, 9) -> SE3.

    Same contract as the reference: no `rot` means identity rotations; one translation with one rotation gives a single SE3,
    anything else a list; mismatched counts raise AssertionError; invalid quaternions / matrices raise ValueError.
    """
    trans = np.asarray(trans)
    rot = np.tile(np.array([1, 0, 0, 0]), (len(trans), 1)) if rot is None else np.asarray(rot)
    scalar_result = trans.ndim == 1 and rot.ndim == 1
    trans2d = trans[None] if trans.ndim == 1 else trans
    rot2d = rot[None] if rot.ndim == 1 else rot
    assert len(trans2d) == len(rot2d), "Numbers of vectors in 'trans' and 'rot' must match."
    if len(trans2d) and rot2d.ndim == 2 and rot2d.shape[1] in (4, 9):
        poses = _compose_rows(trans2d, rot2d)  # one launch for the whole register
    else:
        # ragged rotation rows: per-row dispatch on the row length; rows that are neither quaternions nor matrices are dropped,
        # as the reference does
        poses = []
        for t, r in zip(trans2d, rot2d):
            r = np.asarray(r)
            if r.size == 4:
                poses.append(tq2se3(t, r))
            elif r.size == 9:
                poses.append(tr2se3(t, r.reshape(3, 3)))
    return poses[0] if scalar_result else poses


def homogenize(coord, forth_val=1):
    homog = forth_val * np.ones(4)
    homog[:3] = coord
    return homog
