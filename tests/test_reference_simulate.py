"""The closed-loop path pinned to THE REFERENCE'S OWN CODE: core/simulate.py::simulate, controllers/lqr.py, the planner, sensors and pose
registers are executed unmodified from /root/reference on the functional MuJoCo stand-in (oracle/mujoco_standin.py).

  * live (needs the checkout; skipped on the GPU box): a short run of the reference's simulate() == oracle/replay_oracle.py, i.e. the
    restated loop (frame schedule, lagged qacc / sensordata, control law, sensor-frame twists, noise model) is the reference's;
  * golden (tests/golden/ref_simulate_hammer.npz, written by oracle/gen_golden_simulate.py from the full 1500-step reference run): the
    kernels' rollout algorithm (host build of csrc/rbm_dynamics.cuh here, the CUDA kernel in tests/test_gpu_replay.py) reproduces what the
    reference's simulate() returned.
What the stand-in restates of MuJoCo (plant step, F/T sensor) is listed in its header; that part stays unpinned against MuJoCo itself."""
import os

import numpy as np
import pytest

from conftest import load_golden
from oracle import reference_loader as rl
from oracle import replay_oracle as ro
from oracle import rnea_vec as rv
from rigid_body_manipulation_b200 import engine, identification as idn, planner

os.environ.setdefault("TQDM_DISABLE", "1")


def _consts(g):
    return dict(hposes_Rt=g["hposes_Rt"], simats=g["simats"], uscrews=g["uscrews"], twist_0=g["twist_0"], dtwist_0=g["dtwist_0"])


@pytest.mark.skipif(not rl.available(), reason="reference checkout not present")
def test_reference_simulate_equals_the_restated_loop():
    from oracle.gen_golden_simulate import run_reference_simulation

    res, controller, pl, m = run_reference_simulation("kill_la_kill", duration=0.5)        # 250 steps, 25 frames
    K = controller.gain_matrix
    Kr = ro.lqr_gain(m.consts, m.key_qpos, np.zeros(6), [10.0, 10.0, 10.0, 1e4, 1e4, 1e4])
    assert np.abs(K - Kr).max() < 1e-12 * np.abs(Kr).max()            # the reference's controller on the stand-in == the restatement
    plan_traj = np.array([pl.plan(k) for k in range(pl.n_steps)])
    out = ro.closed_loop_replay(m.consts, m.pose_sen_Rt, m.G_sensed, plan_traj, K, m.key_qpos)
    fr = res["frames"]
    assert len(fr) == len(out["step"]) == 25
    for key, mine in (("twist_sen", out["twist_sen"]), ("dtwist_sen", out["dtwist_sen"]), ("ft_sen", idn.perturb_wrench(out["wrench"], 0.05, 0))):
        theirs = np.array([f[key] for f in fr])
        assert np.abs(theirs - mine).max() < 1e-12 * np.abs(mine).max(), key
    assert np.abs(np.asarray(res["regressors"]) - out["regressor"]).max() < 1e-12 * np.abs(out["regressor"]).max()


@pytest.mark.parametrize("target", ["hammer", "uniform_gearbox"])
def test_rollout_algorithm_reproduces_the_reference_run(target):
    host_harness = pytest.importorskip("host_harness")
    if host_harness.nvcc_path() is None:  # pragma: no cover
        pytest.skip("nvcc is needed to build the host harness")
    g = load_golden(f"ref_simulate_{target}.npz")
    an = engine.analyze_model(g["hposes_Rt"], g["simats"], g["uscrews"], g["twist_0"], g["dtwist_0"], pose_sen_llj=g["pose_sen_llj"])
    assert an[0] == "seq_iso"
    pl = planner.QuinticPlan(g["displacements"], g["key_qpos"], float(g["timestep"]), int(g["n_steps"]))
    phi = ro.inertia_to_phi(g["G_sensed"])
    out = host_harness.closed_loop(an, pl, g["gain_matrix"], phi, g["key_qpos"][None])
    fr = out["frames"][..., 0]
    assert fr.shape[0] == g["twist_sen"].shape[0] == 150
    assert np.abs(fr[:, 18:24] - g["twist_sen"]).max() < 1e-9 * np.abs(g["twist_sen"]).max()
    assert np.abs(fr[:, 24:30] - g["dtwist_sen"]).max() < 1e-9 * np.abs(g["dtwist_sen"]).max()
    noisy = idn.perturb_wrench(fr[:, 30:36], 0.05, 0)                                    # simulate.py:279-290
    assert np.abs(noisy - g["ft_sen"]).max() < 1e-9 * np.abs(g["ft_sen"]).max()
    Y = rv.regressor_batched(fr[:, 18:24], fr[:, 24:30])
    assert np.abs(Y - g["regressors"]).max() < 1e-9 * np.abs(g["regressors"]).max()
    # and the identification the reference's logger would run on its own data (loggers.py:127-129) == ours on ours
    phi_ref = np.linalg.lstsq(g["regressors"].reshape(-1, 10), g["ft_sen"].reshape(-1), rcond=None)[0]
    phi_own = np.linalg.lstsq(Y.reshape(-1, 10), noisy.reshape(-1), rcond=None)[0]
    assert np.abs(phi_ref - phi_own).max() < 1e-7 * np.abs(phi_ref).max()
