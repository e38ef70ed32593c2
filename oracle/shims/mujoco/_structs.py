"""Stub (oracle only): type names used in annotations by the reference."""


class MjModel:  # noqa: D101
    pass


class MjData:  # noqa: D101
    pass


class MjOption:
    """Default option block: MuJoCo's documented defaults (timestep 0.002 s, gravity 0 0 -9.81)
    -- the only fields the reference reads (`core/simulate.py:149`, `joint_position_planner.py:41`)."""

    def __init__(self):
        self.timestep = 0.002
        self.gravity = [0.0, 0.0, -9.81]
