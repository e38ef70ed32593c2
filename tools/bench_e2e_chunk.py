"""e2e throughput of the host entry points (pinned host in / out) against the chunk size of their copy/compute pipelines."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rigid_body_manipulation_b200 import model as rbm_model  # noqa: E402
from rigid_body_manipulation_b200 import planner  # noqa: E402
from rigid_body_manipulation_b200.engine import Model  # noqa: E402

c = rbm_model.load_packaged("sequential", "hammer")
m = Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0)
plan = planner.QuinticPlan([0.2, 1.4, 0.6, np.pi, 0.0, 6 * np.pi], [1, 1, 1, 0, 0, 0], 0.002, 1500)
for n in (1 << 20, 1 << 23):
    traj = torch.randn((n, 3, 6), dtype=torch.float64).pin_memory()
    tau = torch.empty((n, 6), dtype=torch.float64).pin_memory()
    soa = [torch.randn((6, n), dtype=torch.float64).pin_memory() for _ in range(3)]
    tau_soa = torch.empty((6, n), dtype=torch.float64).pin_memory()
    calls = {
        "aos": (lambda ch: m.rnea_host(traj, tau=tau, chunk=ch), 192),
        "soa": (lambda ch: m.rnea_host_soa(*soa, tau=tau_soa, chunk=ch), 168),
        "planned": (lambda ch: m.rnea_planned_host(plan, n=n, step0=0.0, stride=1500.0 / n, tau=tau_soa, chunk=ch), 48),
    }
    for name, (fn, bytes_per_sample) in calls.items():
        for chunk in (0, 1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19):
            for _ in range(3):
                fn(chunk)
            reps = 10 if n <= 1 << 20 else 3
            t0 = time.perf_counter()
            for _ in range(reps):
                fn(chunk)
            dt = (time.perf_counter() - t0) / reps
            print(json.dumps({"entry": name, "n": n, "chunk": chunk, "ms": dt * 1e3, "samples_per_s": n / dt, "pcie_GBps": bytes_per_sample * n / dt / 1e9}), flush=True)
