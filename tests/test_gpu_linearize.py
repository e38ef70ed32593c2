"""-m gpu parity of the batched LQR linearisation (rbm_linearize_f64) against the literal CPU restatement of
mjd_transitionFD on the same plant (oracle/lqr_oracle.py; reference dynamics/dynamics.py:41-46, controllers/lqr.py:43-49)."""
import numpy as np
import pytest
import torch

from gpu_util import load_golden, model_from_golden, sample_states
from oracle import lqr_oracle as lo

pytestmark = pytest.mark.gpu


def consts_of(g):
    return dict(hposes_Rt=g["hposes_Rt"], simats=g["simats"], uscrews=g["uscrews"], twist_0=g["twist_0"], dtwist_0=g["dtwist_0"])


def run(m, q, qd, u, **kw):
    to = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a.T), device="cuda")
    A, B, qdd = m.linearize(to(q), to(qd), to(u), want_qdd=True, **kw)
    torch.cuda.synchronize()
    return A.cpu().numpy(), B.cpu().numpy(), qdd.t().cpu().numpy()


@pytest.mark.parametrize("fname", ["ref_inverse_hammer.npz", "ref_inverse_kill_la_kill.npz"])
@pytest.mark.parametrize("force_generic", [False, True])
def test_matches_transition_fd_oracle(fname, force_generic):
    g = load_golden(fname)
    m = model_from_golden(g, force_generic=force_generic)
    n = 48
    tr = sample_states(np.random.default_rng(3), n)
    q, qd = tr[:, 0], tr[:, 1]
    u = np.random.default_rng(4).standard_normal((n, 6)) * np.array([100, 100, 400, 1, 1, 1.0])
    c = consts_of(g)
    # (a) a well-conditioned step: truncation O(eps^2), round-off O(1e-16/eps) -> both sides agree tightly
    A, B, qdd = run(m, q, qd, u, dt=0.002, eps=1e-5, centered=True)
    Ar, Br = lo.transition_fd(c, q, qd, u, dt=0.002, eps=1e-5, centered=True)
    assert np.abs(qdd - lo.forward_dynamics(c, q, qd, u)).max() < 1e-9 * np.abs(qdd).max()
    assert np.abs(A - Ar).max() < 2e-8
    assert np.abs(B - Br).max() < 2e-8
    # (b) the reference's own defaults (StateSpaceConfig: eps 1e-8, centred): both sides carry ~1e-7 FD round-off
    A, B, _ = run(m, q, qd, u, dt=0.002, eps=1e-8, centered=True)
    Ar, Br = lo.transition_fd(c, q, qd, u, dt=0.002, eps=1e-8, centered=True)
    assert np.abs(A - Ar).max() < 5e-6
    assert np.abs(B - Br).max() < 5e-6
    # (c) forward differences
    A, B, _ = run(m, q, qd, u, dt=0.002, eps=1e-6, centered=False)
    Ar, Br = lo.transition_fd(c, q, qd, u, dt=0.002, eps=1e-6, centered=False)
    assert np.abs(A - Ar).max() < 1e-4  # first-order truncation differs between differentiating ID and FD of the step
    assert np.abs(B - Br).max() < 2e-7


def test_structure_of_the_euler_map():
    g = load_golden("ref_inverse_hammer.npz")
    m = model_from_golden(g)
    n, dt = 1000, 0.002
    tr = sample_states(np.random.default_rng(9), n)
    A, B, _ = run(m, tr[:, 0], tr[:, 1], None, dt=dt, eps=1e-6)
    Minv = lo.mass_matrix(consts_of(g), tr[:, 0])
    Minv = np.linalg.inv(Minv)
    assert np.abs(B[:, 6:, :] - dt * Minv).max() < 1e-12
    assert np.abs(B[:, :6, :] - dt * dt * Minv).max() < 1e-14
    # q+ = q + dt qd+  =>  top block rows are dt * bottom rows + [I 0]
    I = np.eye(6)
    assert np.abs(A[:, :6, :6] - (I + dt * A[:, 6:, :6])).max() < 1e-12
    assert np.abs(A[:, :6, 6:] - dt * A[:, 6:, 6:]).max() < 1e-12
    # prismatic gantry: neither its positions nor (Galilean invariance) its velocities change the dynamics; the kernel skips those
    # finite differences (SequentialDesc::q_matters / qd_matters) -- confirm with the literal transition-FD oracle
    assert np.abs(A[:, 6:, :3]).max() < 1e-6
    assert np.abs(A[:, 6:, 6:9] - np.eye(6)[None, :, :3]).max() < 1e-6
    Ar, _ = lo.transition_fd(consts_of(g), tr[:32, 0], tr[:32, 1], None, dt=dt, eps=1e-5)
    assert np.abs(Ar[:, 6:, :3]).max() < 1e-8 and np.abs(Ar[:, 6:, 6:9] - np.eye(6)[None, :, :3]).max() < 1e-8
    assert np.abs(A[:32] - Ar).max() < 2e-8


def test_keyframe_linearisation_feeds_lqr():
    """The reference's actual use (controllers/lqr.py:34-49): one linearisation at the keyframe at rest, then DARE -> K."""
    from scipy import linalg

    g = load_golden("ref_inverse_hammer.npz")
    m = model_from_golden(g)
    q = g["key_qpos"][None]
    A, B, _ = run(m, q, np.zeros((1, 6)), None, dt=0.002, eps=1e-8)
    Ar, Br = lo.transition_fd(consts_of(g), q, np.zeros((1, 6)), None, dt=0.002, eps=1e-8)
    R = np.diag([10.0, 10, 10, 1e4, 1e4, 1e4])  # configurations/base.yaml:33-39

    def gain(A, B):
        P = linalg.solve_discrete_are(A, B, np.eye(12), R)
        return linalg.pinv(R + B.T @ P @ B) @ B.T @ P @ A

    K, Kr = gain(A[0], B[0]), gain(Ar[0], Br[0])
    assert np.abs(K - Kr).max() < 1e-4 * np.abs(Kr).max()


def test_one_million_states_runs_and_is_consistent():
    """BASELINE.json config 4 size: 2^20 states; spot-check 64 of them against the oracle."""
    g = load_golden("ref_inverse_hammer.npz")
    m = model_from_golden(g)
    n = 1 << 20
    tr = sample_states(np.random.default_rng(17), n)
    A, B, _ = run(m, tr[:, 0], tr[:, 1], None, dt=0.002, eps=1e-6)
    idx = np.random.default_rng(1).choice(n, 64, replace=False)
    Ar, Br = lo.transition_fd(consts_of(g), tr[idx, 0], tr[idx, 1], None, dt=0.002, eps=1e-6)
    assert np.abs(A[idx] - Ar).max() < 5e-8
    assert np.abs(B[idx] - Br).max() < 5e-8
    assert np.isfinite(A).all() and np.isfinite(B).all()


@pytest.mark.parametrize("force_generic", [False, True])
def test_forward_dynamics_and_step_match_the_oracle_and_the_linearisation(force_generic):
    """rbm_forward_dynamics_f64: qdd and the semi-implicit Euler transition against the CPU restatement; and the transition's own
    finite differences (GPU step, eps 1e-6) against the A, B the linearisation kernel returns."""
    g = load_golden("ref_inverse_uniform_gearbox.npz")
    m = model_from_golden(g, force_generic=force_generic)
    c = consts_of(g)
    n, dt = 64, 0.002
    tr = sample_states(np.random.default_rng(21), n)
    q, qd = tr[:, 0], tr[:, 1]
    u = np.random.default_rng(22).standard_normal((n, 6)) * np.array([100, 100, 400, 1, 1, 1.0])
    to = lambda a: torch.as_tensor(np.ascontiguousarray(a.T), device="cuda")
    qdd = m.forward_dynamics(to(q), to(qd), to(u)).t().cpu().numpy()
    ref = lo.forward_dynamics(c, q, qd, u)
    assert np.abs(qdd - ref).max() < 1e-9 * np.abs(ref).max()
    qn, qdn = m.step(to(q), to(qd), to(u), dt=dt)
    y = lo.step(c, q, qd, u, dt)
    assert np.abs(np.concatenate([qn.t().cpu().numpy(), qdn.t().cpu().numpy()], 1) - y).max() < 1e-11 * np.abs(y).max()
    # in-place stepping (aliasing allowed) gives the same bits
    qi, qdi = to(q), to(qd)
    m.step(qi, qdi, to(u), dt=dt, inplace=True)
    assert torch.equal(qi, qn) and torch.equal(qdi, qdn)
    # finite differences of the GPU transition == the linearisation kernel's A, B
    A, B, _ = run(m, q, qd, u, dt=dt, eps=1e-6, centered=True)
    eps = 1e-6
    x0 = np.concatenate([q, qd], 1)
    for i in range(12):
        d = np.zeros(12)
        d[i] = eps
        xp, xm = x0 + d, x0 - d
        yp = torch.cat(m.step(to(xp[:, :6]), to(xp[:, 6:]), to(u), dt=dt)).t().cpu().numpy()
        ym = torch.cat(m.step(to(xm[:, :6]), to(xm[:, 6:]), to(u), dt=dt)).t().cpu().numpy()
        assert np.abs((yp - ym) / (2 * eps) - A[:, :, i]).max() < 5e-8
    for k in range(6):
        d = np.zeros(6)
        d[k] = eps
        yp = torch.cat(m.step(to(q), to(qd), to(u + d), dt=dt)).t().cpu().numpy()
        ym = torch.cat(m.step(to(q), to(qd), to(u - d), dt=dt)).t().cpu().numpy()
        assert np.abs((yp - ym) / (2 * eps) - B[:, :, k]).max() < 5e-8


def test_closed_loop_tracking_with_lqr():
    """End-to-end consistency of the pieces the reference's control loop is made of (core/simulate.py:185-270,
    controllers/lqr.py:38-51), batched over 256 perturbed rollouts: planner -> feed-forward tau (RNEA) -> LQR gain from the
    kernel's A, B (DARE on the host) -> transition kernel.  With the exact model the feed-forward alone tracks the plan up to the
    integrator's O(dt) error, and the feedback pulls perturbed starts back; without feedback they are not pulled back.
    (The stabilising sign u = tau_ff + K (x_target - x) is used; the reference's own loop subtracts, simulate.py:268.)"""
    from scipy import linalg

    from rigid_body_manipulation_b200.planner import traj_5th_spline

    g = load_golden("ref_inverse_hammer.npz")
    m = model_from_golden(g)
    dt, n_steps, nroll = 0.002, 400, 256
    plan = traj_5th_spline([0.2, 0.4, 0.6, np.pi, 0.3 * np.pi, 1.5 * np.pi], g["key_qpos"], dt, n_steps)  # uniform.yaml-like motion
    tau_ff, traj = m.rnea_planned(plan, want_traj=True)  # (6, T), (3, 6, T)
    A, B, _ = run(m, g["key_qpos"][None], np.zeros((1, 6)), None, dt=dt, eps=1e-6)
    R = np.diag([1.0] * 6)
    Q = np.diag([1e6] * 6 + [1e3] * 6)  # stiff enough to converge within the 0.8 s motion (prototyped on the CPU oracle)
    P = linalg.solve_discrete_are(A[0], B[0], Q, R)
    K = torch.as_tensor(linalg.solve(R + B[0].T @ P @ B[0], B[0].T @ P @ A[0]), device="cuda")  # (6, 12)
    rng = np.random.default_rng(0)
    x0 = np.concatenate([g["key_qpos"], np.zeros(6)])[:, None] + rng.standard_normal((12, nroll)) * np.array([[0.02]] * 6 + [[0.05]] * 6)

    def rollout(gain):
        q = torch.as_tensor(x0[:6].copy(), device="cuda")
        qd = torch.as_tensor(x0[6:].copy(), device="cuda")
        for k in range(n_steps):
            xt = torch.cat([traj[0, :, k], traj[1, :, k]])[:, None]
            u = tau_ff[:, k : k + 1] + gain * (K @ (xt - torch.cat([q, qd])))
            m.step(q, qd, u.contiguous(), dt=dt, inplace=True)
        xt = torch.cat([traj[0, :, -1], traj[1, :, -1]])[:, None]
        return (torch.cat([q, qd]) - xt).abs().amax(dim=0).cpu().numpy()

    err_fb, err_open = rollout(1.0), rollout(0.0)
    assert np.isfinite(err_fb).all()
    assert err_fb.max() < 0.05, err_fb.max()          # pulled back onto the plan (what remains is the integrator's O(dt) lag)
    assert np.median(err_open) > 5 * np.median(err_fb)  # without feedback the initial offsets persist / grow


def test_generic_path_with_a_moving_base():
    """with twist_0 != 0 the velocity block of A contains V_0 x (S qd) coupling terms; the generic evaluator used to
    switch the base twist off for the velocity columns.  Checked against the literal transition FD on the same plant."""
    g = load_golden("ref_inverse_hammer.npz")
    from rigid_body_manipulation_b200.engine import Model

    tw0 = np.array([0.3, -0.2, 0.1, 0.4, -0.5, 0.6])
    m = Model(g["hposes_Rt"], g["simats"], g["uscrews"], tw0, g["dtwist_0"], pose_sen_llj=g["pose_sen_llj"])
    assert m.kernel_path == "generic"
    c = consts_of(g)
    c["twist_0"] = tw0
    n = 24
    tr = sample_states(np.random.default_rng(11), n)
    q, qd = tr[:, 0], tr[:, 1]
    u = np.random.default_rng(12).standard_normal((n, 6)) * np.array([100, 100, 400, 1, 1, 1.0])
    A, B, qdd = run(m, q, qd, u, dt=0.002, eps=1e-5, centered=True)
    Ar, Br = lo.transition_fd(c, q, qd, u, dt=0.002, eps=1e-5, centered=True)
    assert np.abs(qdd - lo.forward_dynamics(c, q, qd, u)).max() < 1e-9 * np.abs(qdd).max()
    assert np.abs(A - Ar).max() < 2e-8
    assert np.abs(B - Br).max() < 2e-8
    # the coupling is really there: with the base twist switched off the velocity block differs by far more than the tolerance
    c0 = dict(c, twist_0=np.zeros(6))
    A0, _ = lo.transition_fd(c0, q, qd, u, dt=0.002, eps=1e-5, centered=True)
    assert np.abs(A0[:, 6:, 6:] - Ar[:, 6:, 6:]).max() > 1e-5
    # forward differences use the same (full-base) reference evaluation
    A, B, _ = run(m, q, qd, u, dt=0.002, eps=1e-6, centered=False)
    Ar, Br = lo.transition_fd(c, q, qd, u, dt=0.002, eps=1e-6, centered=False)
    assert np.abs(A - Ar).max() < 1e-4
