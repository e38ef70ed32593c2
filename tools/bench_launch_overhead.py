"""Host enqueue cost of one RNEA launch and eager-loop vs CUDA-graph timing of K back-to-back 2^20-sample launches (why bench.py replays a graph)."""
import time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rigid_body_manipulation_b200 import model as rbm_model
from rigid_body_manipulation_b200.engine import Model
c = rbm_model.load_packaged("sequential", "hammer")
m = Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0, device=0)
B = 1 << 20
sets = []
for i in range(3):
    q = torch.randn(6, B, dtype=torch.float64, device="cuda"); qd = torch.randn_like(q); qdd = torch.randn_like(q)
    sets.append((q, qd, qdd, torch.empty_like(q)))
k = [0]
def step():
    q, qd, qdd, tau = sets[k[0] % 3]; k[0] += 1
    m.rnea(q, qd, qdd, tau=tau)
for _ in range(200): step()
torch.cuda.synchronize()
# host enqueue cost: tiny batch so the GPU never back-pressures
qs = [t[:, :128].contiguous() for t in sets[0]]
t0 = time.perf_counter()
for _ in range(2000): m.rnea(qs[0], qs[1], qs[2], tau=qs[3])
t1 = time.perf_counter(); torch.cuda.synchronize()
print("host enqueue us/call", (t1 - t0) / 2000 * 1e6)
def gpu_time(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for K in (20, 200):
    for trial in range(3):
        print("eager K", K, "us/step", gpu_time(lambda: [step() for _ in range(K)], K))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(K): step()
    g.replay(); torch.cuda.synchronize()
    for trial in range(3):
        print("graph K", K, "us/step", gpu_time(g.replay, K))
