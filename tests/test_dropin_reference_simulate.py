"""`not gpu`: THE REFERENCE'S OWN CALLER ON TOP OF THE DROP-IN (missing #6 / next #7).

core/simulate.py::simulate is loaded UNMODIFIED from /root/reference, but `import dynamics as dyn`, `from transformations import Poses`
(core/simulate.py:13-15), the LQR controller (`from dynamics import StateSpace`, controllers/lqr.py:10) and the planner resolve to
rigid_body_manipulation_b200/dropin.  This container has no GPU, so the engine object inside the drop-in modules is replaced by
tests/host_engine.py: the kernels' own __host__ __device__ arithmetic (csrc/rbm_rnea.cuh generic_rnea, rbm_dynamics.cuh
linearize_state, rbm_setup.cuh transfer_simat / compose / point motion / regressor rows) compiled for the host -- same code, no CUDA.
MuJoCo is the functional stand-in of oracle/mujoco_standin.py, exactly as in the run that produced the golden file.

What is compared: everything simulate() returned on the drop-in (150 frames of sensor twists / twist rates / noisy F/T readings,
the regressors, the controller's gain) against tests/golden/ref_simulate_hammer.npz = the same call on the reference's own modules."""
import os
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import ROOT, load_golden
from oracle import reference_loader as rl

os.environ.setdefault("TQDM_DISABLE", "1")
DROPIN = os.path.join(ROOT, "rigid_body_manipulation_b200", "dropin")

host_harness = pytest.importorskip("host_harness")
pytestmark = pytest.mark.skipif(not rl.available() or host_harness.nvcc_path() is None, reason="needs the reference checkout and nvcc (host harness)")

BASE_YAML = dict(duration=3.0, timestep=-1.0, displacements=[0.2, 1.4, 0.6, 3.141592653589793, 0.0, 18.8495559215],  # configurations/base.yaml
                 input_gain=[10.0, 10.0, 10.0, 1e4, 1e4, 1e4], epsilon=1e-8, centered=True, fps=50)


@pytest.fixture()
def dropin_on_host(monkeypatch):
    import host_engine

    ref = rl.load_simulation_on_dropin(DROPIN)
    # the drop-in modules as imported by the reference's simulate(): swap their engine for the host build of the same arithmetic
    monkeypatch.setattr(ref.dynamics.dynamics, "_engine", host_engine)
    monkeypatch.setattr(ref.transformations.transformations, "_engine", host_engine)
    # ... and the same files under their package-qualified names (mujoco_bridge.constants_from_mujoco imports them that way)
    import rigid_body_manipulation_b200.dropin.dynamics.dynamics as pkg_dyn
    import rigid_body_manipulation_b200.dropin.transformations.transformations as pkg_tf

    monkeypatch.setattr(pkg_dyn, "_engine", host_engine)
    monkeypatch.setattr(pkg_tf, "_engine", host_engine)
    pkg_dyn._MODELS.clear()
    ref.dynamics.dynamics._MODELS.clear()
    return ref


def _run(ref, target, duration):
    from oracle import mujoco_standin as ms

    root = rl.REFERENCE_ROOT
    m = ms.StandinModel(os.path.join(root, "xml_models", "manipulators", "sequential.xml"), os.path.join(root, "xml_models", "targets", target, "object_cad_gt.csv"))
    d = ms.StandinData(m)
    ms.mj_resetDataKeyframe(m, d, 0)                                                   # core/core.py:327
    pcfg = SimpleNamespace(duration=duration, timestep=BASE_YAML["timestep"], pos_offset=d.qpos.copy().tolist(), displacements=list(BASE_YAML["displacements"]))
    planner = ref.planner.JointPositionPlanner(pcfg, m, d)                             # main.py:73 (drop-in planner)
    ccfg = SimpleNamespace(state_space=SimpleNamespace(epsilon=BASE_YAML["epsilon"], centered=BASE_YAML["centered"]), input_gain=list(BASE_YAML["input_gain"]))
    controller = ref.controllers.LinearQuadraticRegulator(ccfg, m, d)                  # main.py:74 (drop-in controller -> drop-in StateSpace)
    logger = SimpleNamespace(fps=BASE_YAML["fps"], cam_id=0, complete_image_dir=Path("/nonexistent"), render=lambda d_, f: None)
    result = ref.simulate.simulate(m, d, logger, planner, controller)                  # main.py:76, the reference's own function
    return result, controller, planner


def test_reference_simulate_runs_on_the_dropin_and_reproduces_its_own_run(dropin_on_host):
    ref = dropin_on_host
    g = load_golden("ref_simulate_hammer.npz")
    res, controller, planner = _run(ref, "hammer", BASE_YAML["duration"])
    assert planner.n_steps == int(g["n_steps"])
    # the gain: drop-in StateSpace (RNEA-based linearisation, host build) + scipy DARE vs the reference's mjd_transitionFD route
    assert np.abs(controller.gain_matrix - g["gain_matrix"]).max() < 1e-4 * np.abs(g["gain_matrix"]).max()
    fr = res["frames"]
    assert len(fr) == g["twist_sen"].shape[0] == 150
    for key, gkey in (("twist_sen", "twist_sen"), ("dtwist_sen", "dtwist_sen"), ("ft_sen", "ft_sen")):
        mine = np.array([f[key] for f in fr])
        # closed loop under a gain that differs by ~1e-5 relative (two linearisation routes): states drift by ~1e-7 over 1500 steps
        assert np.abs(mine - g[gkey]).max() < 1e-5 * np.abs(g[gkey]).max(), key
    assert np.abs(np.asarray(res["regressors"]) - g["regressors"]).max() < 1e-5 * np.abs(g["regressors"]).max()
    assert np.abs(np.array(fr[0]["pose_sen_obj"]) - g["pose_sen_obj"]).max() < 1e-12
    # the identification the reference's logger would run on what simulate() returned (loggers.py:127-129)
    lsq = lambda Y, f: np.linalg.lstsq(np.asarray(Y).reshape(-1, 10), np.asarray(f).reshape(-1), rcond=None)[0]
    phi_dropin = lsq(res["regressors"], [f["ft_sen"] for f in fr])
    phi_ref = lsq(g["regressors"], g["ft_sen"])
    assert np.abs(phi_dropin - phi_ref).max() < 1e-4 * np.abs(phi_ref).max()


def test_same_gain_gives_the_same_run_to_round_off(dropin_on_host):
    """With the controller's gain taken from the golden run the two closed loops see identical inputs, so every quantity the drop-in
    `dynamics.inverse` / `get_regressor_matrix` / `Poses` feed into simulate() must agree with the reference's to round-off."""
    ref = dropin_on_host
    g = load_golden("ref_simulate_hammer.npz")
    from oracle import mujoco_standin as ms

    root = rl.REFERENCE_ROOT
    m = ms.StandinModel(os.path.join(root, "xml_models", "manipulators", "sequential.xml"), os.path.join(root, "xml_models", "targets", "hammer", "object_cad_gt.csv"))
    d = ms.StandinData(m)
    ms.mj_resetDataKeyframe(m, d, 0)
    pcfg = SimpleNamespace(duration=0.6, timestep=-1.0, pos_offset=d.qpos.copy().tolist(), displacements=list(BASE_YAML["displacements"]))
    planner = ref.planner.JointPositionPlanner(pcfg, m, d)
    ms.mjd_transitionFD(m, d, 1e-8, True, np.zeros((12, 12)), np.zeros((12, 6)), None, None)   # what the reference's LQR constructor leaves in `d`
    controller = SimpleNamespace(gain_matrix=g["gain_matrix"])
    logger = SimpleNamespace(fps=BASE_YAML["fps"], cam_id=0, complete_image_dir=Path("/nonexistent"), render=lambda d_, f: None)
    res = ref.simulate.simulate(m, d, logger, planner, controller)
    # the same short run on the reference's own modules
    base = rl.load_simulation()
    m2 = ms.StandinModel(os.path.join(root, "xml_models", "manipulators", "sequential.xml"), os.path.join(root, "xml_models", "targets", "hammer", "object_cad_gt.csv"))
    d2 = ms.StandinData(m2)
    ms.mj_resetDataKeyframe(m2, d2, 0)
    planner2 = base.planner.JointPositionPlanner(SimpleNamespace(duration=0.6, timestep=-1.0, pos_offset=d2.qpos.copy().tolist(),
                                                                 displacements=list(BASE_YAML["displacements"])), m2, d2)
    ms.mjd_transitionFD(m2, d2, 1e-8, True, np.zeros((12, 12)), np.zeros((12, 6)), None, None)
    res2 = base.simulate.simulate(m2, d2, logger, planner2, controller)
    assert len(res["frames"]) == len(res2["frames"]) == 30
    for key in ("twist_sen", "dtwist_sen", "ft_sen"):
        a, b = np.array([f[key] for f in res["frames"]]), np.array([f[key] for f in res2["frames"]])
        assert np.abs(a - b).max() < 1e-11 * np.abs(b).max(), key
    assert np.abs(np.asarray(res["regressors"]) - np.asarray(res2["regressors"])).max() < 1e-11 * np.abs(np.asarray(res2["regressors"])).max()
