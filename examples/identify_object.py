"""The reference's `main.py` experiment (configs[0]) without MuJoCo, on the GPU path end to end:

    model (sequential.xml + a target's CAD row)  ->  quintic joint plan  ->  LQR gain from the keyframe linearisation
    ->  closed-loop replay (feed-forward inverse dynamics + state feedback + plant step, F/T and twist log at 50 fps)
    ->  5 % wrench noise  ->  inertial parameters by least squares  ->  the reference's score

for `--envs` environments in one launch (each gets its own measurement-noise seed and, with --perturb, its own initial state).

    python examples/identify_object.py --target hammer --envs 1024
    python examples/identify_object.py --target hammer --envs 1024 --perturb 0.01 --fix-feedback

The reference's control law is replayed literally by default: `ctrl = tgt_ctrl - K [ (tgt_q - qpos) / nu ; tgt_qd - qvel ]`
(core/simulate.py:257-268).  With res = target - actual that is POSITIVE feedback -- harmless in the reference's own run, which
starts exactly on the plan so only integration error gets amplified (a few 1e-2 rad in 3 s), but perturbed starts drift away.
--fix-feedback uses the stabilising sign and the plain position residual instead (gain -K, divisor 1).
Heavy targets (e.g. kill_la_kill, 16 kg) spun at the plan's 6 pi wrist rotation can blow up under the explicit 2 ms Euler step whatever
the feedback; such environments are reported as diverged and skipped.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rigid_body_manipulation_b200 import identification as idn  # noqa: E402
from rigid_body_manipulation_b200 import model as pm, planner, replay  # noqa: E402
from rigid_body_manipulation_b200.engine import Model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--target", default="hammer", help="one of the packaged targets (rigid_body_manipulation_b200/assets/targets.json)")
    ap.add_argument("--envs", type=int, default=64)
    ap.add_argument("--duration", type=float, default=3.0)          # configurations/base.yaml planner.duration
    ap.add_argument("--show", type=int, default=3, help="environments whose estimates are printed")
    ap.add_argument("--perturb", type=float, default=0.0, help="std of the initial joint-position offsets of environments 1..")
    ap.add_argument("--fix-feedback", action="store_true", help="stabilising feedback sign and plain residual instead of the reference's")
    a = ap.parse_args()

    c = pm.load_packaged("sequential", a.target)
    m = Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0, pose_sen_llj=c.pose_sen_Rt)
    plan = planner.QuinticPlan([0.2, 1.4, 0.6, np.pi, 0.0, 18.8495559215], c.key_qpos, c.timestep, int(a.duration / c.timestep))
    K = replay.lqr_gain(m, c.key_qpos, [10.0, 10.0, 10.0, 1e4, 1e4, 1e4])       # base.yaml controller.input_gain
    phi_true = idn.sensor_frame_params(c.target, c.pose_sen_obj_Rt)
    g = torch.Generator(device="cuda").manual_seed(0)
    q0 = torch.as_tensor(c.key_qpos, device="cuda").reshape(6, 1) + a.perturb * torch.randn((6, a.envs), generator=g, device="cuda", dtype=torch.float64)
    q0[:, 0] = torch.as_tensor(c.key_qpos, device="cuda")                       # environment 0 is the reference's own run
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    log = replay.closed_loop_replay(m, plan, -K if a.fix_feedback else K, phi_true, q0.contiguous(), pos_residual_divisor=1.0 if a.fix_feedback else None)
    torch.cuda.synchronize()
    t_roll = time.perf_counter() - t0
    print(f"kernel path {m.kernel_path}; {a.envs} environments x {plan.n_steps} steps in {t_roll * 1e3:.1f} ms, {log.frame_steps.shape[0]} frames each")
    tgt = plan.trajectory()[log.frame_steps.cpu().numpy()]
    err = (log.trajectory[:, 0] - torch.as_tensor(tgt[:, 0], device="cuda")[..., None]).abs().amax(dim=0)      # (nj, envs)
    finite = ~log.diverged()
    print(f"max tracking error per joint over the {int(finite.sum())} environments that did not diverge:",
          np.array2string(err[:, finite].amax(dim=1).cpu().numpy(), precision=4))
    if not bool(finite[0]):
        raise SystemExit("environment 0 diverged: this target / plan is not integrable with the explicit step (try --duration or another target)")
    print(f"\n{'parameter':>12} {'truth (sensor frame)':>22}" + "".join(f" {'env ' + str(e):>14}" for e in range(min(a.show, a.envs))))
    ests = [replay.identify(m, log, e, perturb=True, seed=e) for e in range(min(a.show, a.envs))]
    for k, name in enumerate(idn.PARAM_LABELS):
        print(f"{name:>12} {phi_true[k]:22.6e}" + "".join(f" {est.phi[k]:14.6e}" for est in ests))
    print(f"{'score':>12} {'':>22}" + "".join(f" {idn.score(est.phi, phi_true, c.target.aabb_scale):14.3e}" for est in ests))
    if a.envs > 1:  # every environment at once: one grouped Gram launch + the host solves
        t0 = time.perf_counter()
        all_ests = replay.identify_all(m, log, perturb=True, seed=0)
        t_id = time.perf_counter() - t0
        good = [e for e in all_ests if e is not None]
        masses = np.array([e.phi[0] for e in good])
        scores = np.array([idn.score(e.phi, phi_true, c.target.aabb_scale) for e in good])
        print(f"\n{len(good)} of {a.envs} environments identified in {t_id * 1e3:.1f} ms ({a.envs - len(good)} rollouts diverged): "
              f"mass median {np.median(masses):.5f} kg, quartiles {np.percentile(masses, 25):.5f} .. {np.percentile(masses, 75):.5f} "
              f"(truth {phi_true[0]:.5f}), median score {np.median(scores):.3e}")
    clean = replay.identify(m, log, 0, perturb=False)
    print(f"\nnoise-free estimate, env 0: score {idn.score(clean.phi, phi_true, c.target.aabb_scale):.3e}, rms residual {clean.rms_residual:.3e} N")
    print("(the reference scores against the object-frame CAD numbers, main.py:79-82; in that frame the score would be "
          f"{idn.score(ests[0].phi, c.target.ground_truth_params, c.target.aabb_scale):.3e})")


if __name__ == "__main__":
    main()
