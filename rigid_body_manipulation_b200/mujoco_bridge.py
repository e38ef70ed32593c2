"""Kernel constants from a compiled MuJoCo model -- the setup block of reference core/simulate.py:74-156.

For callers that DO have MuJoCo (the reference application itself): given `MjModel` / `MjData` (or look-alikes with the
same array attributes and `<type>_names` lists) this reproduces what `simulate()` binds onto `dynamics.inverse`,
using this package's own GPU-backed `transfer_simat` / `get_spatial_inertia_matrix` / `Poses`.  The MuJoCo-free route
(MJCF / CAD numbers -> constants) is rigid_body_manipulation_b200/model.py.
"""
from __future__ import annotations

import numpy as np

from .lie import SE3

SLIDE, HINGE = 2, 3  # mjtJoint


def refresh_kinematics(m, d) -> bool:
    """Run the forward kinematics so that d.xpos / xmat / xipos / site_x* describe the current qpos.  The reference relies on
    mjd_transitionFD (inside the LQR constructor, controllers/lqr.py:34-36) for this side effect before simulate() reads the world
    poses; the replacement StateSpace must therefore provide it.  Real MuJoCo objects get mujoco.mj_forward; MuJoCo-free look-alike
    models may offer the same service as a `forward_kinematics(d)` method.  Returns False when neither applies (the caller's arrays
    are then taken as they are)."""
    fk = getattr(m, "forward_kinematics", None)
    if callable(fk):
        fk(d)
        return True
    try:
        import mujoco
    except ImportError:
        return False
    if isinstance(m, mujoco.MjModel) and isinstance(d, mujoco.MjData):
        mujoco.mj_forward(m, d)
        return True
    return False


def constants_from_mujoco(m, d, last_link="link6", object_body="target/object", sensor_site="target/ft_sensor") -> dict:
    from .dropin.dynamics import get_spatial_inertia_matrix, transfer_simat
    from .dropin.transformations.poses import Poses, _element_id

    refresh_kinematics(m, d)

    poses = Poses(m, d)
    id_ll = _element_id(m, "body", last_link)
    n_chain = id_ll + 1

    # unit screws in the joint frames (simulate.py:98-110)
    uscrews = np.zeros((len(m.jnt_type), 6))
    for k, (t, ax) in enumerate(zip(m.jnt_type, m.jnt_axis)):
        if t == SLIDE:
            uscrews[k, :3] = ax
        elif t == HINGE:
            uscrews[k, 3:] = ax
        else:
            raise TypeError("Only slide or hinge joints, represented as 2 or 3 for an element of m.jnt_type, are supported.")

    # per-body inertias about the principal frames, moved to the joint frames (simulate.py:115-123)
    simats_bi = get_spatial_inertia_matrix(m.body_mass, m.body_inertia)
    simats = transfer_simat(poses.lj_li[:n_chain], simats_bi[:n_chain])

    # bodies behind the last link (attachment frame, object) folded into it (simulate.py:129-137)
    pose_x_llj = poses.x_b[id_ll].dot(poses.l_lj[id_ll])
    extra = np.zeros((6, 6))
    for pose_x_bi, simat_bi in zip(poses.x_bi[n_chain:], simats_bi[n_chain:]):
        extra += transfer_simat(pose_x_llj.inv().dot(pose_x_bi), simat_bi)
    simats[id_ll] += extra

    # joint home poses w.r.t. the parent joint frame (simulate.py:140-146)
    hposes = [SE3.identity()]
    for k in range(len(m.jnt_type)):
        hposes.append(poses.l_lj[k].inv().dot(poses.a_b[k + 1].dot(poses.l_lj[k + 1])).inv())

    gravity = np.asarray(getattr(getattr(m, "opt", None), "gravity", getattr(m, "gravity", [0.0, 0.0, -9.81])), dtype=float)
    out = dict(hposes=hposes, simats=simats, uscrews=uscrews, twist_0=np.zeros(6), dtwist_0=-np.concatenate([gravity, np.zeros(3)]),
               simat_sen_obj=extra)
    try:
        pose_x_sen = poses.get_x_("site", sensor_site)
        out["pose_sen_llj"] = pose_x_sen.inv().dot(pose_x_llj)  # simulate.py:202
        pose_x_obj = poses.get_x_("body", object_body)
        out["pose_sen_obj"] = pose_x_sen.inv().dot(pose_x_obj)
        out["pose_sen_obji"] = pose_x_sen.inv().dot(pose_x_obj.dot(poses.get_b_biof(object_body)))
    except ValueError:
        pass  # model without the target attached
    return out
