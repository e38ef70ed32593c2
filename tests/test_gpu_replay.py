"""Closed-loop replay (reference core/simulate.py:185-290): the one-launch GPU rollout against the literal CPU
restatement of the loop (oracle/replay_oracle.py).  Both sides replace MuJoCo by the same published semantics, so this pins the
kernel to the restatement, not to a MuJoCo run (stated in DESIGN.md)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import model_from_golden
from oracle import replay_oracle as ro
from rigid_body_manipulation_b200 import identification as idn
from rigid_body_manipulation_b200 import model as pm
from rigid_body_manipulation_b200 import planner, replay

pytestmark = pytest.mark.gpu

INPUT_GAIN = [10.0, 10.0, 10.0, 1e4, 1e4, 1e4]  # configurations/base.yaml controller.input_gain
DISP = [0.2, 1.4, 0.6, np.pi, 0.0, 18.8495559215]  # configurations/base.yaml planner.displacements


def _setup(target="hammer", n_steps=1500):
    c = pm.load_packaged("sequential", target)
    consts = dict(hposes_Rt=c.hposes_Rt, simats=c.simats, uscrews=c.uscrews, twist_0=c.twist_0, dtwist_0=c.dtwist_0)
    G_s = ro.sensor_inertia(c.simat_object_llj, c.pose_sen_Rt)
    plan = planner.QuinticPlan(DISP, c.key_qpos, 0.002, n_steps)
    return c, consts, G_s, plan


@pytest.mark.parametrize("force_generic", [False, True])
def test_rollout_matches_the_literal_loop(force_generic):
    from rigid_body_manipulation_b200.engine import Model

    n_steps = 400 if force_generic else 1500
    c, consts, G_s, plan = _setup(n_steps=n_steps)
    m = Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0, pose_sen_llj=c.pose_sen_Rt, force_generic=force_generic)
    K = replay.lqr_gain(m, c.key_qpos, INPUT_GAIN)
    Kr = ro.lqr_gain(consts, c.key_qpos, np.zeros(6), INPUT_GAIN)
    assert np.abs(K - Kr).max() < 1e-4 * np.abs(Kr).max()
    phi = ro.inertia_to_phi(G_s)
    # environment 0 starts at the keyframe, environment 1 from a perturbed state
    q0 = np.stack([c.key_qpos, c.key_qpos + [0.01, -0.02, 0.015, 0.05, -0.04, 0.03]], axis=1)
    qd0 = np.stack([np.zeros(6), [0.02, 0.01, -0.03, 0.1, -0.2, 0.05]], axis=1)
    log = replay.closed_loop_replay(m, plan, Kr, phi, torch.as_tensor(q0, device="cuda"), torch.as_tensor(qd0, device="cuda"))
    steps = log.frame_steps.cpu().numpy()
    plan_traj = plan.trajectory()
    for e in range(2):
        ref = ro.closed_loop_replay(consts, c.pose_sen_Rt, G_s, plan_traj, Kr, q0[:, e], qd0[:, e])
        assert np.array_equal(steps, ref["step"])
        np.testing.assert_allclose(log.time, ref["time"], rtol=0, atol=0)
        got = log.env(e)
        for name, key in (("trajectory", "act"), ("twists_sen", "twist_sen"), ("dtwists_sen", "dtwist_sen"), ("fts_sen", "wrench")):
            scale = np.abs(ref[key]).max()
            assert np.abs(got[name] - ref[key]).max() < 1e-9 * scale, (e, name, np.abs(got[name] - ref[key]).max() / scale)
        fin = log.final[..., e].cpu().numpy()
        assert np.abs(fin[0] - ref["q_final"]).max() < 1e-10 and np.abs(fin[1] - ref["qd_final"]).max() < 1e-10


def test_config1_replay_identifies_the_object():
    """configs[0] end to end: keyframe start, base.yaml plan and gains, 150 logged frames, noise-free and noisy identification."""
    from rigid_body_manipulation_b200.engine import Model

    c, consts, G_s, plan = _setup("hammer")
    m = Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0, pose_sen_llj=c.pose_sen_Rt)
    K = replay.lqr_gain(m, c.key_qpos, INPUT_GAIN)
    phi = idn.sensor_frame_params(c.target, c.pose_sen_obj_Rt)
    log = replay.closed_loop_replay(m, plan, K, phi, n_envs=3)
    assert log.frame_steps.shape[0] == 150 and int(log.frame_steps[0]) == 0 and int(log.frame_steps[1]) == 10
    # tracking: the feed-forward carries the motion (the reference's feedback is weak and enters with the residual's sign)
    tgt = plan.trajectory()[log.frame_steps.cpu().numpy()]
    err = np.abs(log.trajectory[:, 0, :, 0].cpu().numpy() - tgt[:, 0])
    assert err[:, :3].max() < 1e-3 and err[:, 3:].max() < 0.05
    # identical environments give identical logs
    assert torch.equal(log.fts_sen[..., 0], log.fts_sen[..., 2])
    clean = replay.identify(m, log, 0, perturb=False)
    assert abs(clean.phi[0] - phi[0]) < 1e-3 * phi[0]            # mass
    assert np.abs(clean.phi[1:4] - phi[1:4]).max() < 2e-3        # first moments (the one-step lag between qacc and qpos biases them)
    noisy = replay.identify(m, log, 0, perturb=True)
    assert abs(noisy.phi[0] - phi[0]) < 0.05 * phi[0]
    assert idn.score(clean.phi, phi, c.target.aabb_scale) < idn.score(noisy.phi, phi, c.target.aabb_scale)


def test_argument_checks():
    from rigid_body_manipulation_b200.engine import Model

    c, _, _, plan = _setup(n_steps=10)
    m = Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0, pose_sen_llj=c.pose_sen_Rt)
    q0 = torch.zeros((6, 4), dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):
        m.closed_loop(plan, np.zeros((6, 6)), np.zeros(10), q0)
    with pytest.raises(ValueError):
        m.closed_loop(plan, np.zeros((6, 12)), np.zeros(9), q0)
    with pytest.raises(ValueError):
        m.closed_loop(plan, np.zeros((6, 12)), np.zeros(10), q0, pos_residual_divisor=0.0)
    out = m.closed_loop(plan, np.zeros((6, 12)), np.zeros(10), q0, fps=0.0)   # fps 0: only the very first frame
    assert out["frames"].shape[0] == 1


def test_grouped_gram_equals_per_environment_grams():
    from rigid_body_manipulation_b200.engine import Model

    c, consts, G_s, plan = _setup("uniform_gearbox", n_steps=500)
    phi = idn.sensor_frame_params(c.target, c.pose_sen_obj_Rt)
    for force_generic in (False, True):
        m = Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0, pose_sen_llj=c.pose_sen_Rt, force_generic=force_generic)
        K = replay.lqr_gain(m, c.key_qpos, INPUT_GAIN)
        g = torch.Generator(device="cuda").manual_seed(5)
        q0 = torch.as_tensor(c.key_qpos, device="cuda").reshape(6, 1) + 0.01 * torch.randn((6, 37), generator=g, device="cuda", dtype=torch.float64)
        log = replay.closed_loop_replay(m, plan, -K, phi, q0.contiguous(), pos_residual_divisor=1.0)   # stabilising sign: see examples/
        ests = replay.identify_all(m, log, perturb=False)
        assert len(ests) == 37
        for e in (0, 17, 36):
            one = replay.identify(m, log, e, perturb=False)
            assert ests[e].n_samples == one.n_samples == log.frame_steps.shape[0]
            assert np.abs(ests[e].phi - one.phi).max() < 1e-9 * np.abs(one.phi).max()
        # strided views of the log are consumed in place: the packs equal the oracle's Gram of the logged samples
        from oracle import rnea_vec as rv
        e = 5
        tr = log.trajectory[..., e].cpu().numpy()
        out = rv.inverse_batched(tr, c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0)
        Vs, dVs = rv.sensor_frame_twists_batched(c.pose_sen_Rt, out["twists"][:, -1], out["dtwists"][:, -1])
        ref = rv.gram_pack(rv.regressor_batched(Vs, dVs), log.fts_sen[..., e].cpu().numpy())
        got = m.regressor_gram_grouped(log.trajectory[:, 0], log.trajectory[:, 1], log.trajectory[:, 2], log.fts_sen)[e].cpu().numpy()
        assert np.abs(got - ref).max() < 1e-9 * np.abs(ref).max()
    noisy = replay.identify_all(m, log, perturb=True, seed=3)
    masses = np.array([x.phi[0] for x in noisy])
    assert abs(masses.mean() - phi[0]) < 0.05 * phi[0] and masses.std() > 0


@pytest.mark.parametrize("target", ["hammer", "uniform_gearbox"])
def test_kernel_rollout_reproduces_the_reference_simulate_run(target):
    """tests/golden/ref_simulate_hammer.npz = what the reference's own simulate() returned (run unmodified on the MuJoCo stand-in,
    oracle/gen_golden_simulate.py): gain from its LQR class, 150 frames of sensor twists, noisy F/T readings and regressors."""
    from rigid_body_manipulation_b200.engine import Model, regressor_rows

    g = load_golden(f"ref_simulate_{target}.npz")
    m = Model(g["hposes_Rt"], g["simats"], g["uscrews"], g["twist_0"], g["dtwist_0"], pose_sen_llj=g["pose_sen_llj"])
    pl = planner.QuinticPlan(g["displacements"], g["key_qpos"], float(g["timestep"]), int(g["n_steps"]))
    K = replay.lqr_gain(m, g["key_qpos"], INPUT_GAIN)
    assert np.abs(K - g["gain_matrix"]).max() < 1e-4 * np.abs(g["gain_matrix"]).max()     # kernel linearisation + scipy DARE vs the reference's class
    log = replay.closed_loop_replay(m, pl, g["gain_matrix"], ro.inertia_to_phi(g["G_sensed"]), n_envs=2)
    got = log.env(1)
    assert got["twists_sen"].shape == g["twist_sen"].shape
    assert np.abs(got["twists_sen"] - g["twist_sen"]).max() < 1e-9 * np.abs(g["twist_sen"]).max()
    assert np.abs(got["dtwists_sen"] - g["dtwist_sen"]).max() < 1e-9 * np.abs(g["dtwist_sen"]).max()
    assert np.abs(idn.perturb_wrench(got["fts_sen"], 0.05, 0) - g["ft_sen"]).max() < 1e-9 * np.abs(g["ft_sen"]).max()
    Y = regressor_rows(log.twists_sen[..., 1].contiguous(), log.dtwists_sen[..., 1].contiguous()).cpu().numpy()
    assert np.abs(Y - g["regressors"]).max() < 1e-9 * np.abs(g["regressors"]).max()
    est = replay.identify(m, log, 1, perturb=True, seed=0)                                  # same noise stream as the reference's run
    phi_ref = np.linalg.lstsq(g["regressors"].reshape(-1, 10), g["ft_sen"].reshape(-1), rcond=None)[0]
    assert np.abs(est.phi - phi_ref).max() < 1e-6 * np.abs(phi_ref).max()
