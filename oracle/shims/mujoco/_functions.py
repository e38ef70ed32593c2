"""Stub (oracle only): MuJoCo C entry points named by the reference; all raise."""


def _absent(name):
    def f(*a, **k):
        raise RuntimeError(f"mujoco.{name} is not available: MuJoCo is absent from this image (oracle stub)")

    f.__name__ = name
    return f


mjd_transitionFD = _absent("mjd_transitionFD")
mj_name2id = _absent("mj_name2id")
mj_step = _absent("mj_step")
mj_differentiatePos = _absent("mj_differentiatePos")
mj_resetDataKeyframe = _absent("mj_resetDataKeyframe")
