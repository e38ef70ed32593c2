"""Closed-loop replay of the reference's `simulate()` main loop (reference core/simulate.py:185-290) without MuJoCo: planner ->
feed-forward inverse dynamics -> LQR state feedback -> plant step -> F/T + regressor logging -> identification, for MANY
environments per launch (one environment per GPU thread, `rbm_closed_loop_f64`).

What stands in for MuJoCo (absent here, so this path is not pinned against a reference run -- see DESIGN.md):
  * mj_step            : qacc = M(q)^-1 (ctrl - h(q, qd)), semi-implicit Euler (the plant of sequential.xml has no damping, armature,
                         friction or contacts), time += timestep;
  * force/torque sensor: the Newton-Euler wrench Y(V_s, dV_s) phi of the body hanging off the sensor site, in the site frame
                         (MuJoCo's cfrc_int of the site's body), as of the forward pass of the PREVIOUS step -- which is what
                         `d.sensordata` holds when the reference reads it (:218-221), like `d.qacc` (:191);
  * mjd_transitionFD   : `Model.linearize` (controllers/lqr.py:43 -> dynamics.py:41-46) for the LQR gain.
The reference's control-law quirks are kept: the position residual is divided by m.nu (mj_differentiatePos receives m.nu in its dt
slot, :257-263) and the feedback enters as `tgt_ctrl - K res` with res = target - actual (:265-268)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import identification as idn


def lqr_gain(model, q_key, input_gain, dt: float = 0.002, eps: float = 1e-8, centered: bool = True, qd_key=None) -> np.ndarray:
    """controllers/lqr.py:38-51: K = (R + B^T P B)^+ B^T P A with Q = I, R = diag(input_gain), P from the discrete Riccati
    equation, (A, B) the linearisation at the keyframe."""
    from scipy import linalg

    nj = model.nj
    q = torch.as_tensor(np.asarray(q_key, dtype=np.float64).reshape(nj, 1), device=model.device)
    qd = torch.zeros_like(q) if qd_key is None else torch.as_tensor(np.asarray(qd_key, dtype=np.float64).reshape(nj, 1), device=model.device)
    A, B = model.linearize(q, qd, None, dt=dt, eps=eps, centered=centered)
    A, B = A[0].cpu().numpy(), B[0].cpu().numpy()
    Q, R = np.eye(2 * nj), np.diag(np.asarray(input_gain, dtype=np.float64))
    P = linalg.solve_discrete_are(A, B, Q, R)
    return linalg.pinv(R + B.T @ P @ B) @ B.T @ P @ A


@dataclass
class ReplayLog:
    """Device tensors, frame-major; `n` environments in the last axis."""
    frame_steps: torch.Tensor     # (F,) int32: the planner step each frame was logged at
    trajectory: torch.Tensor      # (F, 3, nj, n): qpos, qvel and the previous forward pass' qacc (core/simulate.py:191-194)
    twists_sen: torch.Tensor      # (F, 6, n)
    dtwists_sen: torch.Tensor     # (F, 6, n)
    fts_sen: torch.Tensor         # (F, 6, n): F/T readings [force; torque]
    final: torch.Tensor           # (3, nj, n): state after the last step
    timestep: float

    @property
    def time(self) -> np.ndarray:
        """d.time at each logged frame: the timestep accumulated step by step, as mj_step does."""
        acc, t = [], 0.0
        steps = set(self.frame_steps.cpu().tolist())
        last = max(steps) if steps else -1
        for k in range(last + 1):
            if k in steps:
                acc.append(t)
            t += self.timestep
        return np.asarray(acc)

    def diverged(self, limit: float = 1e3) -> torch.Tensor:
        """(n,) bool: environments whose log is not finite or whose joint positions / velocities left +-limit -- an explicit 2 ms Euler
        step on a fast, heavy wrist (or the reference's feedback sign on a perturbed start) can run away."""
        tr = self.trajectory[:, :2]
        bad = ~torch.isfinite(self.fts_sen).all(dim=0).all(dim=0)
        return bad | ~(tr.abs() < limit).all(dim=0).all(dim=0).all(dim=0)

    def env(self, e: int) -> dict:
        """Host copies of one environment, shaped like the reference's post-processed arrays (core/simulate.py:273-277)."""
        return dict(trajectory=self.trajectory[..., e].cpu().numpy(), twists_sen=self.twists_sen[..., e].cpu().numpy(),
                    dtwists_sen=self.dtwists_sen[..., e].cpu().numpy(), fts_sen=self.fts_sen[..., e].cpu().numpy())


def closed_loop_replay(model, plan, gain, phi_sensed, q0=None, qd0=None, n_envs: int = 1, fps: float = 50.0, pos_residual_divisor=None,
                       max_frames=None) -> ReplayLog:
    """Runs the loop for `n_envs` environments (all from plan.pos_offset at rest unless q0 / qd0 (nj, n) are given)."""
    nj = model.nj
    if q0 is None:
        q0 = torch.as_tensor(np.asarray(plan.pos_offset, dtype=np.float64), device=model.device).reshape(nj, 1).repeat(1, n_envs).contiguous()
    out = model.closed_loop(plan, gain, phi_sensed, q0, qd0, fps=fps, pos_residual_divisor=pos_residual_divisor, max_frames=max_frames)
    fr, n = out["frames"], q0.shape[1]
    F = fr.shape[0]
    return ReplayLog(frame_steps=out["frame_steps"], trajectory=fr[:, : 3 * nj].reshape(F, 3, nj, n), twists_sen=fr[:, 3 * nj : 3 * nj + 6],
                     dtwists_sen=fr[:, 3 * nj + 6 : 3 * nj + 12], fts_sen=fr[:, 3 * nj + 12 :], final=out["final"].reshape(3, nj, n),
                     timestep=float(plan.timestep))


def identify(model, log: ReplayLog, env: int = 0, perturb: bool = True, error_rate: float = 0.05, seed: int = 0):
    """Post-processing of core/simulate.py:279-290 + loggers.py:127-129 for one environment: optional measurement noise on the
    logged wrenches, then the fused regressor + Gram kernel over the logged (qpos, qvel, qacc) and the 10x10 solve."""
    f = log.fts_sen[..., env].cpu().numpy()
    if not np.isfinite(f).all():
        raise ValueError(f"environment {env}: the rollout diverged (non-finite log)")
    if perturb:
        f = idn.perturb_wrench(f, error_rate, seed)
    tr = log.trajectory[..., env]
    q, qd, qdd = (tr[:, k].t().contiguous() for k in range(3))
    ft = torch.as_tensor(np.ascontiguousarray(f.T), device=model.device)
    return idn.solve(model.regressor_gram(q, qd, qdd, ft))


def perturb_wrench_device(fts: torch.Tensor, error_rate: float = 0.05, seed: int = 0) -> torch.Tensor:
    """core/simulate.py:281-290 for every environment at once: fts (F, 6, n) on the device; sigma per environment = error_rate x the
    largest force (resp. torque) norm over its frames.  Same noise MODEL as identification.perturb_wrench, torch's generator instead of
    numpy's (so the draws differ from the reference's; use `identify` for the reference's exact stream)."""
    g = torch.Generator(device=fts.device).manual_seed(seed)
    out = fts.clone()
    fs_std = error_rate * fts[:, :3].norm(dim=1).amax(dim=0)   # (n,)
    ts_std = error_rate * fts[:, 3:].norm(dim=1).amax(dim=0)
    noise = torch.randn(fts.shape, generator=g, device=fts.device, dtype=fts.dtype)
    out[:, :3] += fs_std * noise[:, :3]
    out[:, 3:] += ts_std * noise[:, 3:]
    return out


def identify_all(model, log: ReplayLog, perturb: bool = True, error_rate: float = 0.05, seed: int = 0) -> list:
    """Identification of EVERY environment of a log: one grouped Gram launch (one environment per thread), then the 10x10 solves.
    Returns one identification.Identification per environment, None for environments whose rollout diverged."""
    f = perturb_wrench_device(log.fts_sen, error_rate, seed) if perturb else log.fts_sen
    tr = log.trajectory
    packs = model.regressor_gram_grouped(tr[:, 0], tr[:, 1], tr[:, 2], f.contiguous() if perturb else f).cpu().numpy()
    # a rollout can diverge (ReplayLog.diverged): it gets no estimate
    bad = log.diverged().cpu().numpy()
    return [None if (bad[e] or not np.isfinite(p).all()) else idn.solve(p) for e, p in enumerate(packs)]
