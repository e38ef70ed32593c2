"""Device-resident throughput of the AoS entry point (rbm_rnea_aos_*: the reference's own (n, 3, 6) trajectory layout)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rigid_body_manipulation_b200 import model as rbm_model  # noqa: E402
from rigid_body_manipulation_b200.engine import Model  # noqa: E402


def main():
    c = rbm_model.load_packaged("sequential", "hammer")
    for no_tma in (False, True):
        m = Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0, no_tma=no_tma)
        for dt, es in ((torch.float64, 8), (torch.float32, 4)):
            for n in (1 << 18, 1 << 20, 1 << 24):
                nset = max(2, int(np.ceil(3 * 126e6 / (24 * es * n))) + 1)
                sets = [(torch.randn((n, 3, 6), dtype=dt, device="cuda"), torch.empty((n, 6), dtype=dt, device="cuda")) for _ in range(nset)]
                for t, o in sets[:2]:
                    m.rnea_aos(t, tau=o)
                torch.cuda.synchronize()
                reps = max(8, int(2e8 / n))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for k in range(reps):
                    t, o = sets[k % nset]
                    m.rnea_aos(t, tau=o)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                print(json.dumps({"layout": "AoS", "tma": not no_tma, "dtype": str(dt).split(".")[1], "n": n, "ms": ms,
                                  "samples_per_s": n / ms * 1e3, "GBps_algorithmic": 24 * es * n / ms / 1e6, "frac": 24 * es * n / ms / 1e6 / 6550.4}))
                del sets


if __name__ == "__main__":
    main()
