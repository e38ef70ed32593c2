"""MJCF text of the manipulator + attached target from the packaged descriptions (model.py), for cross-checks against MuJoCo
where MuJoCo is available (`tests/test_mujoco_crosscheck.py`, skipped otherwise -- MuJoCo is absent from the build image and
from the GPU boxes, so this path has NOT been executed there).

It mirrors what the reference assembles with dm_control (reference core/core.py:195-329): the serial chain of
xml_models/manipulators/sequential.xml, a massless frame body at the attachment site and the object body with the explicit
inertial derived from the CAD row (pos = CoM, quat = principal frame, diaginertia), plus the F/T site with euler "0 0 180"."""
from __future__ import annotations

import numpy as np

from .model import Robot, Target


def _fmt(v):
    return " ".join(repr(float(x)) for x in np.asarray(v).reshape(-1))


def _R_to_quat(R):
    """Rotation matrix -> wxyz quaternion (Shepperd's method)."""
    R = np.asarray(R, float)
    tr = np.trace(R)
    if tr > 0:
        s = 2.0 * np.sqrt(1.0 + tr)
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = 2.0 * np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k])
        q = [0.0, 0.0, 0.0, 0.0]
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    q = np.array(q)
    return q / np.linalg.norm(q) * (1.0 if q[0] >= 0 else -1.0)


def to_mjcf(robot: Robot, target: Target | None) -> str:
    lines = ['<mujoco model="manipulator">', f'  <option gravity="{_fmt(robot.gravity)}" timestep="{robot.timestep!r}"/>', "  <worldbody>"]
    indent = "    "
    for L in robot.links:
        lines.append(f'{indent}<body name="{L.name}" pos="{_fmt(L.pos)}" quat="{_fmt(_R_to_quat(L.R))}">')
        indent += "  "
        lines.append(f'{indent}<joint name="j_{L.name}" type="{L.joint_type}" axis="{_fmt(L.joint_axis)}" pos="{_fmt(L.joint_pos)}"/>')
        lines.append(f'{indent}<inertial pos="{_fmt(L.ipos)}" quat="{_fmt(_R_to_quat(L.iR))}" mass="{L.mass!r}" diaginertia="{_fmt(L.diaginertia)}"/>')
    att = robot.attachment_T
    lines.append(f'{indent}<body name="target" pos="{_fmt(att[:3, 3])}" quat="{_fmt(_R_to_quat(att[:3, :3]))}">')
    sen = robot.sensor_T_in_attachment
    lines.append(f'{indent}  <site name="ft_sensor" pos="{_fmt(sen[:3, 3])}" quat="{_fmt(_R_to_quat(sen[:3, :3]))}"/>')
    if target is not None:
        iq = _R_to_quat(target.R_principal_from_body.T)
        lines.append(f'{indent}  <body name="object">')
        lines.append(f'{indent}    <inertial pos="{_fmt(target.com)}" quat="{_fmt(iq)}" mass="{target.mass!r}" diaginertia="{_fmt(target.diaginertia)}"/>')
        lines.append(f"{indent}  </body>")
    lines.append(f"{indent}</body>")
    for _ in robot.links:
        indent = indent[:-2]
        lines.append(f"{indent}</body>")
    lines += ["  </worldbody>", "</mujoco>"]
    return "\n".join(lines)
