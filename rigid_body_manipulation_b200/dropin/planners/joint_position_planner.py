"""`planners.joint_position_planner` of the reference (planners/joint_position_planner.py:15-131) over this package's QuinticPlan:
the producer of the hot path's inputs.  Same names, configuration fields, attributes and `planner.plan(step)` contract; in addition
`planner.plan` is a QuinticPlan, so the whole trajectory (`plan.trajectory()`) or the planner-driven kernels
(`Model.rnea_planned(planner.plan)`, `Model.closed_loop(planner.plan, ...)`) are one call."""
from __future__ import annotations

import ast
import numbers
import operator
import re
from dataclasses import dataclass, field
from math import pi
from typing import Any, Union

from rigid_body_manipulation_b200.planner import QuinticPlan, traj_5th_spline  # noqa: F401

MUJOCO_DEFAULT_TIMESTEP = 0.002  # MjOption().timestep, used when cfg.timestep <= 0 (reference :42)


@dataclass
class JointPositionPlannerConfig:
    target_class: str = "JointPositionPlanner"
    duration: Any = None          # MISSING in the reference (omegaconf); must be set
    timestep: float = -1
    pos_offset: Any = None        # MISSING in the reference: filled from d.qpos
    displacements: list = field(default_factory=lambda: [0.2, 0.4, 0.6, 1.0 * pi, 0.3 * pi, 1.5 * pi])


_OPS = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv, ast.USub: operator.neg,
        ast.UAdd: operator.pos}


def _arith(node):
    if isinstance(node, ast.Expression):
        return _arith(node.body)
    if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
        return node.value
    if isinstance(node, ast.BinOp) and type(node.op) in _OPS:
        return _OPS[type(node.op)](_arith(node.left), _arith(node.right))
    if isinstance(node, ast.UnaryOp) and type(node.op) in _OPS:
        return _OPS[type(node.op)](_arith(node.operand))
    raise ValueError("Invalid characters in expression.")


class JointPositionPlanner:
    def __init__(self, cfg: JointPositionPlannerConfig, m=None, d=None) -> None:
        pos_offset = getattr(cfg, "pos_offset", None)
        if pos_offset is None or (isinstance(pos_offset, str) and pos_offset == "???"):   # reference :34-38
            if d is None:
                raise ValueError("pos_offset is missing and no MjData was given to read qpos from")
            pos_offset = [float(x) for x in d.qpos]
            try:
                cfg.pos_offset = list(pos_offset)
            except Exception:  # frozen / structured configs keep their own copy
                pass
        if cfg.duration is None:
            raise ValueError("duration is mandatory")
        self.duration = cfg.duration
        if cfg.timestep <= 0:
            try:
                from mujoco._structs import MjOption

                self.timestep = MjOption().timestep
            except Exception:
                self.timestep = MUJOCO_DEFAULT_TIMESTEP
        else:
            self.timestep = cfg.timestep
        self.n_steps = int(self.duration / self.timestep)
        self.displacements = [self._number(x) for x in cfg.displacements]
        self.plan = traj_5th_spline(self.displacements, pos_offset, self.timestep, self.n_steps)

    def _number(self, value) -> float:
        if isinstance(value, numbers.Real):
            return float(value)
        text = repr(value).strip("'")                      # reference :46-52: numbers or strings such as "6 * pi"
        try:
            return float(text)
        except ValueError:
            return self.safe_eval(self.replace_pi(text))

    def safe_eval(self, expr):
        """Arithmetic on literals only (digits . + * / - ( ) and blanks); anything else raises ValueError like the reference's
        character filter, without handing the string to eval()."""
        allowed_chars = "0123456789.+*/-() "
        if not all(ch in allowed_chars for ch in expr):
            raise ValueError("Invalid characters in expression.")
        try:
            return _arith(ast.parse(expr, mode="eval"))
        except SyntaxError as e:
            raise ValueError("Invalid characters in expression.") from e

    def replace_pi(self, text):
        return re.sub(r"\bpi\b", str(pi), text)
