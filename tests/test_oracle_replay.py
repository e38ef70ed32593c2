"""CPU checks of the closed-loop restatement (oracle/replay_oracle.py; reference core/simulate.py:185-290) that the GPU rollout
is compared with in tests/test_gpu_replay.py."""
import numpy as np

from conftest import load_golden
from oracle import replay_oracle as ro
from oracle import rnea_vec as rv
from rigid_body_manipulation_b200 import identification as idn
from rigid_body_manipulation_b200 import model as pm


def _setup():
    c = pm.load_packaged("sequential", "hammer")
    consts = dict(hposes_Rt=c.hposes_Rt, simats=c.simats, uscrews=c.uscrews, twist_0=c.twist_0, dtwist_0=c.dtwist_0)
    return c, consts, ro.sensor_inertia(c.simat_object_llj, c.pose_sen_Rt)


def test_sensed_inertia_and_wrench_model():
    c, consts, G_s = _setup()
    phi = ro.inertia_to_phi(G_s)
    # the two routes to the sensed body's parameters agree: moved spatial inertia vs the CAD numbers in the sensor frame
    assert np.abs(phi - idn.sensor_frame_params(c.target, c.pose_sen_obj_Rt)).max() < 1e-9
    # F = G dV - ad(V)^T G V  ==  Y(V, dV) phi   
    rng = np.random.default_rng(0)
    V, dV = rng.standard_normal((5, 6)), rng.standard_normal((5, 6)) * 3
    Y = rv.regressor_batched(V, dV)
    for k in range(5):
        assert np.abs(ro._ft_reading(G_s, V[k : k + 1], dV[k : k + 1]) - Y[k] @ phi).max() < 1e-12


def test_frame_schedule_and_short_rollout():
    c, consts, G_s = _setup()
    g = load_golden("ref_config1_hammer.npz")  # reference-generated plan (1500 steps of base.yaml)
    K = ro.lqr_gain(consts, c.key_qpos, np.zeros(6), [10, 10, 10, 1e4, 1e4, 1e4])
    assert K.shape == (6, 12) and np.all(np.diag(K[:, :6]) > 0) and np.all(np.diag(K[:, 6:]) > 0)
    out = ro.closed_loop_replay(consts, c.pose_sen_Rt, G_s, g["traj"][:220], K, c.key_qpos)
    assert list(out["step"]) == list(range(0, 220, 10))          # 50 fps on a 2 ms step
    # first logged frame: at rest at the keyframe, qacc / wrench as left by a forward pass without control -> free fall along the
    # vertical slider (world z = joint 3), the sensor reads the inertial reaction, not the weight
    assert np.abs(out["act"][0][2] - [0, 0, -9.81, 0, 0, 0]).max() < 1e-9
    assert np.abs(out["wrench"][0]).max() < 1e-9
    # afterwards the feed-forward holds the arm: the F/T sensor carries the object's weight (m g)
    m_obj = G_s[0, 0]
    assert abs(np.linalg.norm(out["wrench"][1][:3]) - m_obj * 9.81) < 0.05 * m_obj * 9.81
    # the lagged acceleration: frame k logs qacc of the forward pass one step earlier
    assert np.abs(out["act"][1:, 0] - g["traj"][out["step"][1:], 0]).max() < 1e-3


def test_replay_log_flags_diverged_environments():
    import torch

    from rigid_body_manipulation_b200.replay import ReplayLog

    F, n = 5, 4
    tr, f = torch.zeros(F, 3, 6, n, dtype=torch.float64), torch.zeros(F, 6, n, dtype=torch.float64)
    tr[2, 0, 3, 1] = 1e5          # a joint position ran away
    f[1, 2, 2] = float("nan")     # a non-finite wrench
    tr[0, 2, 0, 3] = 1e9          # a large logged acceleration alone does not count
    log = ReplayLog(frame_steps=torch.arange(F, dtype=torch.int32), trajectory=tr, twists_sen=f, dtwists_sen=f, fts_sen=f,
                    final=torch.zeros(3, 6, n, dtype=torch.float64), timestep=0.002)
    assert log.diverged().tolist() == [False, True, True, False]
    assert np.allclose(log.time, [0.0, 0.002, 0.004, 0.006, 0.008])
