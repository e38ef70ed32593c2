"""`controllers.lqr` of the reference (controllers/lqr.py:13-51): the caller of `dynamics.StateSpace`.  Same names, fields and
`gain_matrix` contract; the linearisation underneath is the GPU kernel (rbm_linearize_f64) through the drop-in StateSpace, the
12x12 Riccati solve stays scipy on the host as in the reference."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

import numpy as np
from scipy import linalg

from dynamics import StateSpace, StateSpaceConfig   # the drop-in package (this directory is on sys.path next to it)


@dataclass
class LinearQuadraticRegulatorConfig:
    target_class: str = "LinearQuadraticRegulator"
    state_space: StateSpaceConfig = field(default_factory=StateSpaceConfig)
    input_gain: Any = None        # MISSING in the reference: ones(m.nu)


class LinearQuadraticRegulator:
    def __init__(self, cfg: LinearQuadraticRegulatorConfig, m, d) -> None:
        self.ss = StateSpace(cfg.state_space, m, d)
        gain = getattr(cfg, "input_gain", None)
        if gain is None or (isinstance(gain, str) and gain == "???"):     # reference :28-32: ones(m.nu)
            gain = np.ones(self.ss.B.shape[1]).tolist()
            try:
                cfg.input_gain = gain
            except Exception:
                pass
        self.input_gain = gain
        self.gain_matrix = self.update_control_gain(m, d)

    def update_control_gain(self, m, d):
        self.ss.update_matrices(m, d)
        Q = np.eye(self.ss.ns)
        R = np.diag(self.input_gain)
        P = linalg.solve_discrete_are(self.ss.A, self.ss.B, Q, R)
        return linalg.pinv(R + self.ss.B.T @ P @ self.ss.B) @ self.ss.B.T @ P @ self.ss.A
