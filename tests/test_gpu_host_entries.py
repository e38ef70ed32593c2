"""-m gpu: the host (end-to-end) entry points of the C ABI -- AoS, SoA with live rows only, planner-driven -- against the device
kernels and the oracle (reference dynamics/dynamics.py:109-157 behind core/simulate.py:187-188)."""
import numpy as np
import pytest
import torch

from gpu_util import load_golden, model_from_golden, rel_err, sample_states, soa
from oracle import rnea_vec as rv
from rigid_body_manipulation_b200 import planner

pytestmark = pytest.mark.gpu


def _oracle_tau(g, traj):
    return rv.inverse_batched(traj, g["hposes_Rt"], g["simats"], g["uscrews"], g["twist_0"], g["dtwist_0"])["tau"]


def test_live_inputs_mask_matches_the_structure():
    g = load_golden("ref_inverse_hammer.npz")
    live = model_from_golden(g).live_inputs()
    assert live.shape == (3, 6) and live.sum() == 15 and not live[0, :3].any() and live[0, 3:].all() and live[1:].all()
    assert model_from_golden(g, force_generic=True).live_inputs().all()
    # the claim behind the mask, checked on the oracle: tau does not move when the gantry positions do
    traj = sample_states(np.random.default_rng(5), 64)
    moved = traj.copy()
    moved[:, 0, :3] += np.random.default_rng(6).uniform(-3, 3, (64, 3))
    assert np.abs(_oracle_tau(g, traj) - _oracle_tau(g, moved)).max() < 1e-9


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-9), (np.float32, 1e-4)])
@pytest.mark.parametrize("n,chunk", [(1, 0), (1000, 0), (70_001, 4096), (300_000, 0), (131_072 * 2 + 17, 131_072)])
@pytest.mark.parametrize("force_generic", [False, True])
def test_soa_host_entry_matches_the_oracle(dtype, tol, n, chunk, force_generic):
    g = load_golden("ref_inverse_hammer.npz")
    if force_generic and n > 100_000:
        pytest.skip("generic path: small sizes are enough")
    m = model_from_golden(g, force_generic=force_generic)
    traj = sample_states(np.random.default_rng(n), n)
    ref = _oracle_tau(g, traj)
    q, qd, qdd = (np.ascontiguousarray(traj[:, k, :].T).astype(dtype) for k in range(3))
    # NaN in the dead rows of the fast path proves they never reach the kernel (the generic path reads every row)
    if not force_generic:
        q[:3] = np.nan
    tau = m.rnea_host_soa(q, qd, qdd, chunk=chunk)
    assert tau.shape == (6, n) and tau.dtype == dtype
    assert rel_err(tau.T, ref).max() < tol
    # pinned tensors and a caller-supplied output with a canary row pitch
    qt, qdt, qddt = (torch.as_tensor(a).pin_memory() for a in (q, qd, qdd))
    out = torch.full((6, n), 7.0, dtype=qt.dtype).pin_memory()
    got = m.rnea_host_soa(qt, qdt, qddt, tau=out, chunk=chunk)
    assert got is out and np.array_equal(out.numpy(), tau)


def test_host_entries_reject_bad_buffers():
    g = load_golden("ref_inverse_hammer.npz")
    m = model_from_golden(g)
    traj = sample_states(np.random.default_rng(0), 32)
    q, qd, qdd = (np.ascontiguousarray(traj[:, k, :].T) for k in range(3))
    with pytest.raises(ValueError):
        m.rnea_host_soa(q, qd, qdd[:, :16])
    with pytest.raises(ValueError):
        m.rnea_host_soa(q, qd, qdd, tau=np.empty((6, 31)))
    with pytest.raises(ValueError):
        m.rnea_host_soa(q, qd, qdd, tau=np.empty((6, 32), dtype=np.float32))
    with pytest.raises(ValueError):
        m.rnea_host_soa(torch.as_tensor(q, device="cuda"), torch.as_tensor(qd, device="cuda"), torch.as_tensor(qdd, device="cuda"))
    with pytest.raises(ValueError):
        m.rnea_host(traj, tau=np.empty((31, 6)))
    with pytest.raises(ValueError):
        m.rnea_host(traj, tau=torch.empty((32, 6), dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        m.rnea_host(torch.as_tensor(traj).transpose(1, 2))  # non-contiguous view
    dev = torch.as_tensor(traj, device="cuda")
    with pytest.raises(ValueError):
        m.rnea_aos(dev, tau=torch.empty((31, 6), dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        m.rnea_aos(dev, tau=torch.empty((32, 6), dtype=torch.float32, device="cuda"))


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 1e-4)])
@pytest.mark.parametrize("force_generic", [False, True])
def test_planned_host_entry_reproduces_config1(dtype, tol, force_generic):
    """base.yaml plan (planners/joint_position_planner.py:86-131): tau of all 1500 steps, generated in the kernel and brought to the
    host in chunks, against the reference-generated golden trajectory; then a long fractional-stride batch against the device entry."""
    g = load_golden("ref_inverse_hammer.npz")
    c1 = load_golden("ref_config1_hammer.npz")  # produced by the reference's own planner + dynamics.inverse (oracle/gen_golden.py)
    pg = load_golden("ref_planner.npz")
    m = model_from_golden(g, force_generic=force_generic)
    plan = planner.traj_5th_spline(pg["base_disp"], [1, 1, 1, 0, 0, 0], 0.002, int(pg["base_n_steps"]))
    ref = c1["tau"]
    tau = m.rnea_planned_host(plan, dtype=dtype, chunk=256)
    assert tuple(tau.shape) == (6, plan.n_steps)
    assert rel_err(tau.numpy().T, ref).max() < tol
    n = 300_001
    stride = plan.n_steps / n
    dev = m.rnea_planned(plan, n=n, step0=0.0, stride=stride, dtype=dtype)
    host = m.rnea_planned_host(plan, n=n, step0=0.0, stride=stride, dtype=dtype, chunk=65_536)
    # chunk c starts at step0 + c * chunk * stride: identical arithmetic up to the rounding of that product
    assert rel_err(host.numpy().T, dev.t().cpu().numpy()).max() < (1e-11 if dtype == torch.float64 else 1e-5)
    with pytest.raises(ValueError):
        m.rnea_planned_host(plan, tau=torch.empty((6, 7), dtype=dtype))
