"""Drop-in for reference transformations/poses.py:8-49: the register of model poses.

Works with a real MuJoCo `MjModel` / `MjData` (when `mujoco` is installed) and with any object exposing the same
array attributes (`body_pos`, `body_quat`, `body_ipos`, `body_iquat`, `jnt_pos`; `xpos`, `xmat`, `xipos`, `ximat`,
`cam_xpos`, `cam_xmat`, `site_xpos`, `site_xmat`)."""
from rigid_body_manipulation_b200.lie import SE3

from .transformations import compose

__all__ = ["Poses"]

_OBJ = {"body": 1, "joint": 3, "site": 6, "camera": 7, "sensor": 18, "numeric": 19, "keyframe": 23}  # mjtObj


def _element_id(m, elem_type, name):
    """reference utilities.py:22-47 (mj_name2id) with a fallback for MuJoCo-free model stand-ins."""
    if elem_type not in _OBJ:
        raise ValueError(f"'{elem_type}' is not supported for now. Use mj_name2id and check the value of an ID instead.")
    idx = -1
    if hasattr(m, f"{elem_type}_names"):
        names = list(getattr(m, f"{elem_type}_names"))
        idx = names.index(name) if name in names else -1
    else:
        try:
            import mujoco

            idx = mujoco.mj_name2id(m, _OBJ[elem_type], name)
        except ImportError as e:
            raise ValueError(f"cannot resolve '{name}': mujoco is not installed and the model has no {elem_type}_names") from e
    if -1 == idx:
        raise ValueError(f"ID for '{name}' not found. Check the manipulator .xml or the object .xml")
    return idx


class Poses:
    def __init__(self, m, d) -> None:
        self.m = m
        self.a_b = compose(m.body_pos, m.body_quat)
        self.b_bi = compose(m.body_ipos, m.body_iquat)
        self.x_b = compose(d.xpos, d.xmat)
        self.x_bi = compose(d.xipos, d.ximat)
        cam_pos, cam_mat = getattr(d, "cam_xpos", None), getattr(d, "cam_xmat", None)
        self.x_cam = compose(cam_pos, cam_mat) if cam_pos is not None and len(cam_pos) else []
        self.x_site = compose(d.site_xpos, d.site_xmat)
        self.l_lj = [SE3.identity()] + compose(m.jnt_pos)
        self.lj_li = [l_lj.inv().dot(l_li) for l_lj, l_li in zip(self.l_lj, self.b_bi)]

    # -- name-based getters (reference poses.py:25-49); "pricipal" [sic] is the key the reference uses ----------
    _REGISTERS = {"body": ("x_b", "body"), "pricipal": ("x_bi", "body"), "camera": ("x_cam", "camera"), "site": ("x_site", "site")}

    def get_a_(self, name) -> SE3:
        """Pose of body `name` in its parent's frame."""
        return self.a_b[_element_id(self.m, "body", name)]

    def get_b_biof(self, name) -> SE3:
        """Pose of the inertial (principal) frame of body `name` in the body frame."""
        return self.b_bi[_element_id(self.m, "body", name)]

    def get_x_(self, elem_type, name) -> SE3:
        """World pose of a body / principal frame / camera / site."""
        try:
            register, id_type = self._REGISTERS[elem_type]
        except KeyError:
            raise ValueError(f"Pose retrieval for element type {elem_type} is not supported for now") from None
        return getattr(self, register)[_element_id(self.m, id_type, name)]
