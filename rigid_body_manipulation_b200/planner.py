"""Quintic rest-to-rest joint trajectories -- the producer of the hot path's inputs.

Semantics of reference planners/joint_position_planner.py:86-131 (`traj_5th_spline`): a normalised 5th-order polynomial
s(k) in the integer step variable k in [init_step, init_step + n_steps] with s = 0 -> 1 and zero end velocities /
accelerations; pos = disp * s + offset, vel = disp * s' / dt, acc = disp * s'' / dt^2.  The six coefficients come
from the same (badly scaled: entries up to n_steps^5) 6x6 linear system the reference solves, so they agree with the
reference's to round-off; evaluation is vectorised over all steps at once instead of one Python call per step.
"""
from __future__ import annotations

import numpy as np


class QuinticPlan:
    def __init__(self, displacement, pos_offset, timestep: float, n_steps: int, init_step: int = 0):
        self.displacement = np.asarray(displacement, dtype=np.float64)
        self.pos_offset = np.asarray(pos_offset, dtype=np.float64)
        self.timestep = float(timestep)
        self.n_steps = int(n_steps)
        self.init_step = int(init_step)
        k0, k1 = float(init_step), float(init_step + n_steps)

        def rows(k):
            return [[k**5, k**4, k**3, k**2, k, 1.0], [5 * k**4, 4 * k**3, 3 * k**2, 2 * k, 1.0, 0.0], [20 * k**3, 12 * k**2, 6 * k, 2.0, 0.0, 0.0]]

        (p0, v0, a0), (p1, v1, a1) = rows(k0), rows(k1)
        system = np.array([p0, p1, v0, v1, a0, a1], dtype=float)  # row order of the reference: pos, pos, vel, vel, acc, acc
        self.coeffs = np.linalg.solve(system, np.array([0.0, 1.0, 0.0, 0.0, 0.0, 0.0]))

    # ---- scalar API of the reference: plan(step) -> (3, n_joints) ---------------------------------------------
    def __call__(self, step):
        return self.trajectory(np.array([step]))[0]

    # ---- vectorised ---------------------------------------------------------------------------------------------
    def profile(self, steps):
        """s, ds/dk, d2s/dk2 at the given (possibly fractional) steps."""
        k = np.asarray(steps, dtype=np.float64)
        c = self.coeffs
        P = np.stack([k**5, k**4, k**3, k**2, k, np.ones_like(k)], axis=-1)
        s = P @ c
        ds = P[..., 1:] @ (c[:5] * np.array([5.0, 4.0, 3.0, 2.0, 1.0]))
        dds = P[..., 2:] @ (c[:4] * np.array([20.0, 12.0, 6.0, 2.0]))
        return s, ds, dds

    def trajectory(self, steps=None):
        """(len(steps), 3, n_joints) array of [pos; vel; acc] rows -- the `traj` argument of dynamics.inverse."""
        if steps is None:
            steps = np.arange(self.init_step, self.init_step + self.n_steps)
        s, ds, dds = self.profile(steps)
        d = self.displacement
        pos = s[:, None] * d + self.pos_offset
        vel = ds[:, None] * d / self.timestep
        acc = dds[:, None] * d / self.timestep**2
        return np.stack([pos, vel, acc], axis=1)


def traj_5th_spline(displacement, pos_offset, timestep: float, n_steps: int, init_step: int = 0):
    """Same call signature and return kind as the reference: a callable plan(step) -> ndarray (3, n_joints)."""
    return QuinticPlan(displacement, pos_offset, timestep, n_steps, init_step)
