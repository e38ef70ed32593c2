"""BASELINE.json configs[4]: throughput sweep of the batched inverse dynamics over batch sizes, fp32 and fp64, at 1/2/4/8 GPUs.

    python tools/sweep_rnea.py [--max-exp 9] [--out profiles/r1_sweep_rnea.jsonl]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep_rnea.py --out profiles/r2_sweep_rnea_8gpu.jsonl

Under torchrun the B samples of every size are sharded contiguously over the ranks (distributed.shard_range; no data-path collective),
each launch loop is bracketed by barrier + synchronize, timed with CUDA events, and the MAX over ranks is reported by rank 0.

Inputs are generated on the device (seeded); every size is timed over enough back-to-back launches to last >= ~50 ms,
with rotating buffer sets when one batch is smaller than 3x the L2.  One JSON line per (dtype, B).  The planner-driven
kernel (trajectory generated in-kernel, no input traffic) is swept alongside.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rigid_body_manipulation_b200 import model as rbm_model  # noqa: E402
from rigid_body_manipulation_b200.engine import Model  # noqa: E402
from rigid_body_manipulation_b200.planner import traj_5th_spline  # noqa: E402

L2 = 126e6
PEAK = 6550.4


def gen(B, dt, gen_):
    q = torch.empty((6, B), dtype=dt, device="cuda")
    q[:3] = torch.rand((3, B), generator=gen_, device="cuda", dtype=dt) * 4 - 1.5
    q[3:] = (torch.rand((3, B), generator=gen_, device="cuda", dtype=dt) * 2 - 1) * (6 * np.pi)
    qd = torch.randn((6, B), generator=gen_, device="cuda", dtype=dt)
    qdd = torch.randn((6, B), generator=gen_, device="cuda", dtype=dt) * 3
    return q, qd, qdd


WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))


def barrier():
    if WORLD > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def timed(fn, min_ms=50.0):
    """ms per call; the repetition count is decided from the max-reduced time, so it is identical on every rank"""
    for _ in range(3):
        fn()
    reps = 4
    while True:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if WORLD > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        ms = ms.item()
        if ms >= min_ms or reps >= 1 << 16:
            return ms / reps
        reps *= 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-exp", type=int, default=9)
    ap.add_argument("--out", default="")
    ap.add_argument("--generic", action="store_true", help="force the generic (any-model) kernel")
    ap.add_argument("--min-exp", type=int, default=4)
    ap.add_argument("--dtypes", default="f32,f64")
    ap.add_argument("--model", default="hammer", help="hammer (the reference robot) or nj6: the structure-free 6-joint model of "
                    "tests/golden/ref_inverse_generic_nj6.npz (dense inertias, general screws, moving base, tip wrench)")
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if WORLD > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    from rigid_body_manipulation_b200 import distributed as rbm_dist

    if args.model == "nj6":
        g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_inverse_generic_nj6.npz"))
        m = Model(g["hposes_Rt"], g["simats"], g["uscrews"], g["twist_0"], g["dtwist_0"], wrench_tip=g["wrench_tip"], pose_tip_ee=g["pose_tip"], device=local)
    else:
        c = rbm_model.load_packaged("sequential", "hammer")
        m = Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0, force_generic=args.generic, device=local)
    free_b = torch.cuda.mem_get_info()[0]
    lines = []
    for name, dt, es in (("f32", torch.float32, 4), ("f64", torch.float64, 8)):
        if name not in args.dtypes.split(","):
            continue
        for e in range(args.min_exp, args.max_exp + 1):
            B_total = 10**e
            a_, b_ = rbm_dist.shard_range(B_total, RANK, WORLD)
            B = max(b_ - a_, 2)  # this rank's shard (>= 2 keeps the degenerate sizes launchable on every rank)
            alg = 24 * es * B
            nset = 1 if alg > 8 * L2 else int(np.ceil(3 * L2 / alg)) + 1
            nset = min(nset, 64)
            if nset * alg > 0.9 * free_b:
                lines.append({"dtype": name, "B": B, "skipped": f"{nset * alg / 1e9:.0f} GB does not fit in HBM on one GPU; chunk or shard"})
                continue
            g = torch.Generator(device="cuda").manual_seed(1234 + RANK)
            sets = [gen(B, dt, g) + (torch.empty((6, B), dtype=dt, device="cuda"),) for _ in range(nset)]
            k = [0]

            def step():
                q, qd, qdd, tau = sets[k[0] % nset]
                k[0] += 1
                m.rnea(q, qd, qdd, tau=tau)

            ms = timed(step)
            plan = traj_5th_spline([0.2, 1.4, 0.6, np.pi, 0.0, 6 * np.pi], [1, 1, 1, 0, 0, 0], 0.002, 1500)
            taus = [s[3] for s in sets]

            def pstep():
                k[0] += 1
                m.rnea_planned(plan, n=B, step0=0.0, stride=1500.0 / B, dtype=dt, tau=taus[k[0] % nset])

            pms = timed(pstep)
            gbs = 24 * es * B_total / (ms * 1e-3) / 1e9  # aggregate over all ranks
            lines.append({"model": args.model, "kernel_path": m.kernel_path, "dtype": name, "n_gpus": WORLD, "B": B_total, "B_per_gpu": B, "ms": ms, "samples_per_s": B_total / (ms * 1e-3),
                          "GBps_algorithmic": gbs, "frac_of_measured_hbm": gbs / PEAK / WORLD,
                          "buffer_sets": nset, "planned_ms": pms, "planned_samples_per_s": B_total / (pms * 1e-3),
                          "planned_GBps_algorithmic": 6 * es * B_total / (pms * 1e-3) / 1e9})
            if RANK == 0:
                print(json.dumps(lines[-1]), flush=True)
            del sets, taus
            torch.cuda.empty_cache()
    if args.out and RANK == 0:
        with open(args.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
    if WORLD > 1:
        torch.distributed.destroy_process_group()
