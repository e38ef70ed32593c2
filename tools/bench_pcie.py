"""Bare host<->device copy bandwidth per rank (no kernel): the ceiling of the end-to-end (`e2e`) legs of bench.py.

    python tools/bench_pcie.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tools/bench_pcie.py   # all ranks at once: what the ranks get when they share the host

For each rank: pinned host buffers of 128 MiB, cudaMemcpyAsync H2D alone, D2H alone, and both directions concurrently on two
streams (what the pipelined host entry points do), timed on the device with CUDA events after a barrier, max over ranks.
Rank 0 prints one JSON line with per-rank numbers and the aggregate.  `--numa` binds each rank to its GPU's NUMA node first.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=128)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--numa", action="store_true")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    numa = None
    if args.numa:
        from rigid_body_manipulation_b200 import numa as rbm_numa

        numa = rbm_numa.bind_to_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.mib << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.ones(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(mode):
        def once():
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)

        for _ in range(3):
            once()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for _ in range(args.iters):
            once()
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        nbytes = n * args.iters * (2 if mode == "both" else 1)
        return nbytes / (ms.item() * 1e-3) / 1e9  # GB/s per rank at the pace of the slowest rank

    res = {m: run(m) for m in ("h2d", "d2h", "both")}
    topo = None
    try:
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
        topo = {"nodes": nodes, "cpus_allowed": len(os.sched_getaffinity(0))}
    except OSError:
        pass
    if rank == 0:
        print(json.dumps({"n_gpus": world, "mib": args.mib, "numa_bind": bool(args.numa), "numa_rank0": numa and {k: numa.get(k) for k in ("pci", "numa_node", "bound", "why")},
                          "host_topology": topo,
                          "per_rank_GBps": res, "aggregate_GBps": {k: v * world for k, v in res.items()}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
