"""(Lives under tests/ because it times the CPU oracle next to the GPU path; nothing outside tests/, smoke() and bench.py's CPU arm may
touch oracle/.)

BASELINE.json configs[0] (the reference's own run: base.yaml, 1500 control steps, two `dynamics.inverse` calls per step,
a 6x10 regressor at the ~151 frame steps, lstsq at the end, one LQR linearisation) -- the open-loop part that does not
need MuJoCo, timed three ways on this machine:

  reference_cpu : the per-sample oracle port of the reference's algorithm, called like core/simulate.py:187-224 does
  dropin_scalar : the same loop through the drop-in `dynamics` package (one tiny GPU launch per call)
  batched       : the whole planned trajectory in one launch per kernel (what the B200 path is for)

Prints one JSON line.  (The closed loop itself -- state feedback through the plant step -- is `rigid_body_manipulation_b200/replay.py`
/ `rbm_closed_loop_f64`; see examples/identify_object.py and tools/bench_replay.py.)
"""
import json
import os
import sys
import time
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rigid_body_manipulation_b200", "dropin"))

import dynamics as dyn  # noqa: E402  (the drop-in)
from oracle import rnea_oracle as ro  # noqa: E402  (CPU checker / baseline only)
from rigid_body_manipulation_b200 import identification, model  # noqa: E402
from rigid_body_manipulation_b200.engine import Model  # noqa: E402
from rigid_body_manipulation_b200.lie import se3_from_Rt  # noqa: E402
from rigid_body_manipulation_b200.planner import traj_5th_spline  # noqa: E402


def main():
    c = model.load_packaged("sequential", "hammer")
    plan = traj_5th_spline([0.2, 1.4, 0.6, np.pi, 0.0, 6 * np.pi], c.key_qpos, c.timestep, int(3.0 / c.timestep))  # base.yaml:18-26
    n = plan.n_steps
    trajs = plan.trajectory()
    frames = np.arange(0, n, 10)  # 50 fps at dt = 0.002
    phi = identification.sensor_frame_params(c.target, c.pose_sen_obj_Rt)
    out = {"steps": n, "frames": len(frames)}

    # ---- reference algorithm on the CPU (oracle port), the way simulate() calls it --------------------------------
    hp = [ro.SE3(ro.SO3(r[:9].reshape(3, 3).copy()), r[9:].copy()) for r in c.hposes_Rt]
    inv_cpu = partial(ro.inverse, hposes_body_parent=hp, simats_body=c.simats, uscrews_body=c.uscrews, twist_0=c.twist_0, dtwist_0=c.dtwist_0)
    sen = ro.SE3(ro.SO3(c.pose_sen_Rt[:9].reshape(3, 3).copy()), c.pose_sen_Rt[9:].copy())
    t0 = time.perf_counter()
    taus, Ys = [], []
    for k in range(n):
        tau, _, _, _ = inv_cpu(trajs[k])       # feed-forward (simulate.py:188)
        _, _, tw, dtw = inv_cpu(trajs[k])      # second call on the measured state (simulate.py:194); same cost
        taus.append(tau)
        if k % 10 == 0:
            Vs, dVs = ro.sensor_frame_twists(sen, tw[6], dtw[6])
            Ys.append(ro.regressor(Vs, dVs))
    Y = np.array(Ys)
    f = identification.perturb_wrench(Y @ phi)
    phi_cpu = np.linalg.lstsq(Y.reshape(-1, 10), f.reshape(-1), rcond=None)[0]
    out["reference_cpu_s"] = time.perf_counter() - t0

    # ---- drop-in scalar API, same loop ---------------------------------------------------------------------------
    hp2 = [se3_from_Rt(r) for r in c.hposes_Rt]
    inv_gpu = partial(dyn.inverse, hposes_body_parent=hp2, simats_body=c.simats, uscrews_body=c.uscrews, twist_0=c.twist_0, dtwist_0=c.dtwist_0)
    sen2 = se3_from_Rt(c.pose_sen_Rt)
    inv_gpu(trajs[0])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    taus2, Ys2 = [], []
    for k in range(n):
        tau, _, _, _ = inv_gpu(trajs[k])
        _, _, tw, dtw = inv_gpu(trajs[k])
        taus2.append(tau)
        if k % 10 == 0:
            Ad = sen2.adjoint()
            Ys2.append(dyn.get_regressor_matrix(Ad @ tw[6], Ad @ dtw[6]))
    out["dropin_scalar_s"] = time.perf_counter() - t0
    out["dropin_scalar_us_per_inverse_call"] = 1e6 * out["dropin_scalar_s"] / (2 * n)

    # ---- batched ---------------------------------------------------------------------------------------------------
    m = Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0, pose_sen_llj=c.pose_sen_Rt)
    def batched():
        tau_b, traj_b = m.rnea_planned(plan, want_traj=True)                      # plan + feed-forward for all 1500 steps
        sel = torch.as_tensor(frames, device="cuda")
        qf, qdf, qddf = (traj_b[k].index_select(1, sel).contiguous() for k in range(3))
        fd = torch.as_tensor(f, device="cuda").t().contiguous()
        ident = identification.solve(m.regressor_gram(qf, qdf, qddf, fd))          # regressor + Gram + solve
        q0 = torch.as_tensor(c.key_qpos.reshape(6, 1), device="cuda")
        A, B = m.linearize(q0, torch.zeros_like(q0), None, dt=c.timestep)         # the one LQR linearisation
        return tau_b.t().cpu().numpy(), ident, A.cpu().numpy(), B.cpu().numpy()

    batched()  # warm-up (workspace allocation, first-launch module load)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tau_h, ident, A, B = batched()
    out["batched_s"] = time.perf_counter() - t0

    taus, taus2 = np.array(taus), np.array(taus2)
    scale = np.abs(taus).max(axis=1, keepdims=True)
    out["parity_dropin_vs_cpu"] = float((np.abs(taus2 - taus) / scale).max())
    out["parity_batched_vs_cpu"] = float((np.abs(tau_h - taus) / scale).max())
    out["parity_regressor"] = float(np.abs(np.array(Ys2) - Y).max() / np.abs(Y).max())
    out["parity_phi_batched_vs_lstsq"] = float(np.abs(ident.phi - phi_cpu).max())
    out["speedup_dropin"] = out["reference_cpu_s"] / out["dropin_scalar_s"]
    out["speedup_batched"] = out["reference_cpu_s"] / out["batched_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
