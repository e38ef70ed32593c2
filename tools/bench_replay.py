"""Throughput of the one-launch closed-loop rollout (rbm_closed_loop_f64): environments x steps per second.

    python tools/bench_replay.py [--envs 1 1024 9472 37888 151552] [--steps 1500] > profiles/rX_replay.jsonl
Each environment is the configs[0] experiment (base.yaml plan and gains, hammer) from a slightly different initial state."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rigid_body_manipulation_b200 import identification as idn  # noqa: E402
from rigid_body_manipulation_b200 import model as pm, planner, replay  # noqa: E402
from rigid_body_manipulation_b200.engine import Model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, nargs="+", default=[1, 1024, 9472, 37888, 151552])
ap.add_argument("--steps", type=int, default=1500)
ap.add_argument("--generic", action="store_true")
a = ap.parse_args()

c = pm.load_packaged("sequential", "hammer")
m = Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0, pose_sen_llj=c.pose_sen_Rt, force_generic=a.generic)
plan = planner.QuinticPlan([0.2, 1.4, 0.6, np.pi, 0.0, 18.8495559215], c.key_qpos, 0.002, a.steps)
K = replay.lqr_gain(m, c.key_qpos, [10.0, 10.0, 10.0, 1e4, 1e4, 1e4])
phi = idn.sensor_frame_params(c.target, c.pose_sen_obj_Rt)
for n in a.envs:
    g = torch.Generator(device="cuda").manual_seed(n)
    q0 = torch.as_tensor(c.key_qpos, device="cuda").reshape(6, 1) + 0.01 * torch.randn((6, n), generator=g, device="cuda", dtype=torch.float64)
    q0 = q0.contiguous()
    for _ in range(2):
        out = m.closed_loop(plan, K, phi, q0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        out = m.closed_loop(plan, K, phi, q0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"tool": "bench_replay", "kernel_path": "generic" if a.generic else "fast", "n_envs": n, "n_steps": a.steps, "ms": ms,
                      "env_steps_per_s": n * a.steps / ms * 1e3, "frames_logged": int(out["frames"].shape[0]),
                      "note": "includes the host-side allocation of the frame log (torch.zeros) in the timed region"}), flush=True)
