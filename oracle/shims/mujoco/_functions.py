"""Stub (oracle only): MuJoCo C entry points named by the reference.  Called on a oracle.mujoco_standin.StandinModel they run the
functional stand-in documented there; on anything else they raise (MuJoCo itself is absent from this image)."""


def _entry(name):
    def f(m, *a, **k):
        from oracle import mujoco_standin as ms

        if isinstance(m, ms.StandinModel):
            return getattr(ms, name)(m, *a, **k)
        raise RuntimeError(f"mujoco.{name} is not available: MuJoCo is absent from this image (oracle stub)")

    f.__name__ = name
    return f


mjd_transitionFD = _entry("mjd_transitionFD")
mj_name2id = _entry("mj_name2id")
mj_step = _entry("mj_step")
mj_forward = _entry("mj_forward")
mj_differentiatePos = _entry("mj_differentiatePos")
mj_resetDataKeyframe = _entry("mj_resetDataKeyframe")
