"""e2e throughput of rbm_rnea_host_f64 (pinned host in / out) against the chunk size of its copy/compute pipeline."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rigid_body_manipulation_b200 import model as rbm_model  # noqa: E402
from rigid_body_manipulation_b200.engine import Model  # noqa: E402

c = rbm_model.load_packaged("sequential", "hammer")
m = Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0)
for n in (1 << 20, 1 << 24):
    traj = torch.randn((n, 3, 6), dtype=torch.float64).pin_memory()
    tau = torch.empty((n, 6), dtype=torch.float64).pin_memory()
    for chunk in (0, 1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19):
        for _ in range(3):
            m.rnea_host(traj, tau=tau, chunk=chunk)
        reps = 10 if n <= 1 << 20 else 3
        t0 = time.perf_counter()
        for _ in range(reps):
            m.rnea_host(traj, tau=tau, chunk=chunk)
        dt = (time.perf_counter() - t0) / reps
        print(json.dumps({"n": n, "chunk": chunk, "ms": dt * 1e3, "samples_per_s": n / dt, "pcie_GBps": 192 * n / dt / 1e9}))
