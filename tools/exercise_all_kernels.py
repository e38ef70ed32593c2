"""Launch every kernel of librbm_b200.so once on small inputs (fast + generic paths, fp64 + fp32, ragged sizes).
Used as the target of `compute-sanitizer --tool memcheck|racecheck|synccheck` and as a quick smoke run."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rigid_body_manipulation_b200 import engine, model  # noqa: E402
from rigid_body_manipulation_b200.planner import traj_5th_spline  # noqa: E402


def main():
    c = model.load_packaged("sequential", "hammer")
    rng = np.random.default_rng(0)
    for force_generic in (False, True):
        m = engine.Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0, pose_sen_llj=c.pose_sen_Rt, force_generic=force_generic)
        for n in (1, 300, 2049):
            traj = np.stack([rng.uniform(-3, 3, (n, 6)), rng.standard_normal((n, 6)), rng.standard_normal((n, 6))], axis=1)
            for dt in (torch.float64, torch.float32):
                dev = torch.as_tensor(traj, dtype=dt, device="cuda")
                q, qd, qdd = (dev[:, k, :].t().contiguous() for k in range(3))
                tau, V, dV = m.rnea(q, qd, qdd, want_twists=True)
                m.rnea_aos(dev)
                out = m.regressor_from_traj(q, qd, qdd, want_rows=True, want_twists=True, phi=np.arange(10) * 0.1)
                m.regressor_gram(q, qd, qdd, out["wrench"])
                m.rnea_host(traj.astype(np.float64 if dt == torch.float64 else np.float32), chunk=512)
            dev = torch.as_tensor(traj, device="cuda")
            q, qd, qdd = (dev[:, k, :].t().contiguous() for k in range(3))
            m.rnea_full(dev)
            m.linearize(q, qd, qdd, eps=1e-6)
            m.linearize(q, qd, None, eps=1e-6, centered=False)
        plan = traj_5th_spline([0.2, 1.4, 0.6, np.pi, 0.0, 6 * np.pi], [1, 1, 1, 0, 0, 0], 0.002, 1500)
        m.rnea_planned(plan, want_traj=True)
        m.rnea_planned(plan, dtype=torch.float32)
        # the TMA-pipelined Gram kernel needs an aligned, multi-tile fp32 batch
        n = 256 * 7 + 36
        q = torch.randn((6, n), dtype=torch.float32, device="cuda")
        m.regressor_gram(q, q * 0.5, q * 2.0, q * 0.1)
    n = 37
    Rt = np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0.1, 0.2, 0.3.real]), (n, 1))
    G = np.tile(np.eye(6), (n, 1, 1))
    engine.transfer_simat(Rt, G)
    engine.transfer_simat(Rt, G, adjoint_form=True)
    engine.coordinate_transfer_imat(Rt, np.tile(np.eye(3), (n, 1, 1)), np.ones(n))
    engine.spatial_inertia(np.ones(n), np.ones((n, 3)))
    engine.compose_poses(np.zeros((n, 3)), np.tile([1.0, 0, 0, 0], (n, 1)))
    engine.compose_poses(np.zeros((n, 3)), np.tile(np.eye(3).reshape(9), (n, 1)))
    engine.point_motion(np.ones((n, 6)), np.ones((n, 6)), np.ones((n, 3)))
    engine.regressor_rows(np.ones((n, 6)), np.ones((n, 6)))
    engine.sensor_twists(Rt[0], np.ones((n, 6)), np.ones((n, 6)))
    torch.cuda.synchronize()
    print("exercised all kernels")


if __name__ == "__main__":
    main()
