"""-m gpu: the reference's own call surface (`import dynamics as dyn`, `from transformations import ...`) served by the
drop-in packages, compared with golden vectors produced by the reference's files (tests/golden/ref_setup_functions.npz,
ref_inverse_*.npz).  Written the way the reference's callers use the API (core/simulate.py:115-156,188,194,202-224)."""
import os
import sys
from functools import partial
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu

DROPIN = os.path.join(ROOT, "rigid_body_manipulation_b200", "dropin")


@pytest.fixture(scope="module")
def mods():
    sys.path.insert(0, DROPIN)
    for k in [k for k in sys.modules if k.split(".")[0] in ("dynamics", "transformations")]:
        del sys.modules[k]
    import dynamics as dyn
    import transformations as tf

    assert os.path.realpath(dyn.__file__).startswith(os.path.realpath(DROPIN))
    yield SimpleNamespace(dyn=dyn, tf=tf)
    sys.path.remove(DROPIN)


def Rt_of(p):
    return np.concatenate([np.asarray(p.rot.as_matrix()).reshape(9), np.asarray(p.trans)])


def test_public_names_match_the_reference(mods):
    for name in ["StateSpaceConfig", "StateSpace", "get_spatial_inertia_matrix", "transfer_simat", "inverse", "extract_linvel_frame_transferred",
                 "extract_linacc_frame_transferred", "get_regressor_matrix", "coordinate_transfer_imat", "coordinate_transfer_simat"]:
        assert hasattr(mods.dyn, name)
    for name in ["tq2se3", "tr2se3", "compose", "homogenize", "Poses"]:
        assert hasattr(mods.tf, name)
    cfg = mods.dyn.StateSpaceConfig()
    assert cfg.epsilon == 1e-8 and cfg.centered is True


def test_setup_functions(mods):
    dyn, tf = mods.dyn, mods.tf
    g = load_golden("ref_setup_functions.npz")
    poses = tf.compose(g["poses_Rt"][:, 9:].copy(), g["poses_Rt"][:, :9].copy())
    assert np.allclose([Rt_of(p) for p in poses], g["poses_Rt"], atol=0)
    assert np.array_equal(dyn.get_spatial_inertia_matrix(g["mass"], g["diag"]), g["simats_diag"])
    assert np.allclose(dyn.transfer_simat(poses, g["simats_diag"]), g["transferred"], rtol=1e-13, atol=1e-13)
    assert np.allclose(dyn.transfer_simat(poses[3], g["simats_diag"][3]), g["transferred_single"], rtol=1e-13, atol=1e-13)
    assert np.allclose(dyn.transfer_simat(poses, g["dense"]), g["transferred_dense"], rtol=1e-12, atol=1e-12)
    for k in range(4):
        assert np.allclose(dyn.coordinate_transfer_imat(poses[k], g["imats"][k], g["mass"][k]), g["ct_imat"][k], rtol=1e-13, atol=1e-13)
        assert np.allclose(dyn.coordinate_transfer_simat(poses[k], g["dense"][k]), g["ct_simat"][k], rtol=1e-12, atol=1e-12)
        assert np.allclose(dyn.extract_linvel_frame_transferred(g["twists"][k], poses[k]), g["linvel"][k], atol=1e-14)
        assert np.allclose(dyn.extract_linvel_frame_transferred(g["twists"][k], poses[k], homogeneous=True), g["linvel_h"][k], atol=1e-14)
        assert np.allclose(dyn.extract_linacc_frame_transferred(g["twists"][k], g["dtwists"][k], poses[k]), g["linacc"][k], atol=1e-13)
        assert np.allclose(dyn.extract_linacc_frame_transferred(g["twists"][k], g["dtwists"][k], poses[k], homogeneous=True), g["linacc_h"][k], atol=1e-13)
        assert np.allclose(dyn.get_regressor_matrix(g["twists"][k], g["dtwists"][k]), g["regressor"][k], atol=1e-13)
    with pytest.raises(AssertionError):
        dyn.get_spatial_inertia_matrix(g["mass"], g["diag"][:-1])
    with pytest.raises(AssertionError):
        dyn.transfer_simat(poses[:3], g["dense"][:4])


def test_compose_variants_and_errors(mods):
    tf = mods.tf
    g = load_golden("ref_setup_functions.npz")
    assert np.allclose([Rt_of(p) for p in tf.compose(g["trans"], g["quats"])], g["comp_q"], atol=1e-15)
    assert np.allclose([Rt_of(p) for p in tf.compose(g["trans"])], g["comp_none"], atol=0)
    single = tf.compose(g["trans"][0], g["quats"][0])
    assert not isinstance(single, list) and np.allclose(Rt_of(single), g["comp_single"], atol=1e-15)
    assert np.array_equal(tf.homogenize(g["trans"][0]), g["hom"]) and np.array_equal(tf.homogenize(g["trans"][1], 0), g["hom0"])
    with pytest.raises(ValueError):
        tf.compose(g["trans"][:1], np.array([[1.0, 1.0, 0, 0]]))  # non-unit quaternion
    with pytest.raises(ValueError):
        tf.compose(g["trans"][:1], np.diag([1.0, 1.0, -1.0]).reshape(1, 9))  # reflection
    with pytest.raises(AssertionError):
        tf.compose(g["trans"][:2], g["quats"][:3])
    # translation rows stay views of the caller's buffer (the reference's "dynamic" world poses)
    buf = g["trans"][:2].copy()
    ps = tf.compose(buf, g["quats"][:2])
    buf[0, 0] = 123.0
    assert ps[0].trans[0] == 123.0


@pytest.mark.parametrize("fname", ["ref_inverse_hammer.npz", "ref_inverse_generic_nj4.npz", "ref_inverse_generic_nj9.npz"])
def test_inverse_scalar_api_like_simulate(mods, fname):
    dyn = mods.dyn
    from rigid_body_manipulation_b200.lie import se3_from_Rt

    g = load_golden(fname)
    nj = g["uscrews"].shape[0]
    extra = dict(wrench_tip=g["wrench_tip"], pose_tip_ee=se3_from_Rt(g["pose_tip"])) if "wrench_tip" in g.files else {}
    inverse = partial(dyn.inverse, hposes_body_parent=[se3_from_Rt(r) for r in g["hposes_Rt"]], simats_body=g["simats"],
                      uscrews_body=g["uscrews"], twist_0=g["twist_0"], dtwist_0=g["dtwist_0"], **extra)  # simulate.py:150-156
    pose_sen = se3_from_Rt(g["pose_sen_llj"])
    for s in range(12):
        tgt_ctrl, poses, twists, dtwists = inverse(g["traj"][s])  # simulate.py:188,194
        assert tgt_ctrl.shape == (nj,) and len(poses) == nj + 1 and len(twists) == nj + 1 and len(dtwists) == nj + 1
        assert np.allclose(tgt_ctrl, g["tau"][s], rtol=1e-10, atol=1e-10 * np.abs(g["tau"][s]).max())
        assert np.allclose(np.array(twists), g["twists"][s], rtol=1e-10, atol=1e-12)
        assert np.allclose(np.array(dtwists), g["dtwists"][s], rtol=1e-10, atol=1e-10)
        assert np.allclose([Rt_of(p) for p in poses[:nj]], g["poses"][s], atol=1e-12)
        # the per-frame block of simulate.py:202-224
        twist_sen = pose_sen.adjoint() @ twists[nj]
        dtwist_sen = pose_sen.adjoint() @ dtwists[nj]
        assert np.allclose(dyn.get_regressor_matrix(twist_sen, dtwist_sen), g["regressor"][s], rtol=1e-9, atol=1e-9 * np.abs(g["regressor"][s]).max())
    with pytest.raises(ValueError):
        inverse(g["traj"][0][:, :-1])


def test_mujoco_bridge_and_state_space(mods):
    """constants_from_mujoco on MjModel / MjData look-alike arrays == the constants the reference's setup block produces;
    StateSpace on the same objects == the transition-FD oracle."""
    from oracle import lqr_oracle as lo
    from rigid_body_manipulation_b200.mujoco_bridge import constants_from_mujoco

    z = load_golden("ref_mjmodel_hammer.npz")
    g = load_golden("ref_inverse_hammer.npz")
    m = SimpleNamespace(**{k: z[k] for k in z.files if k not in ("body_names", "site_names")}, body_names=list(z["body_names"]),
                        site_names=list(z["site_names"]), nv=6, nu=6, na=0, nsensordata=51, timestep=0.002)
    d = SimpleNamespace(qpos=z["qpos"], qvel=np.zeros(6), ctrl=np.zeros(6), xpos=z["xpos"], xmat=z["xmat"], xipos=z["xipos"], ximat=z["ximat"],
                        site_xpos=z["site_xpos"], site_xmat=z["site_xmat"], cam_xpos=np.zeros((0, 3)), cam_xmat=np.zeros((0, 9)))
    c = constants_from_mujoco(m, d)
    assert np.allclose([Rt_of(h) for h in c["hposes"]], g["hposes_Rt"], atol=1e-15)
    assert np.abs(c["simats"] - g["simats"]).max() < 1e-13 * np.abs(g["simats"]).max()
    assert np.array_equal(c["uscrews"], g["uscrews"]) and np.array_equal(c["dtwist_0"], g["dtwist_0"])
    assert np.allclose(Rt_of(c["pose_sen_llj"]), g["pose_sen_llj"], atol=1e-15)
    assert np.abs(c["simat_sen_obj"] - g["simat_sen_obj"]).max() < 1e-13 * np.abs(g["simats"]).max()
    ss = mods.dyn.StateSpace(mods.dyn.StateSpaceConfig(), m, d)  # controllers/lqr.py:34
    assert ss.ns == 12 and ss.A.shape == (12, 12) and ss.B.shape == (12, 6) and ss.C.shape == (51, 12) and ss.D.shape == (51, 6)
    consts = dict(hposes_Rt=g["hposes_Rt"], simats=g["simats"], uscrews=g["uscrews"], twist_0=g["twist_0"], dtwist_0=g["dtwist_0"])
    Ar, Br = lo.transition_fd(consts, z["qpos"][None], np.zeros((1, 6)), None, dt=0.002, eps=1e-8)
    assert np.abs(ss.A - Ar[0]).max() < 5e-6 and np.abs(ss.B - Br[0]).max() < 5e-6


def test_controller_and_planner_dropins_drive_the_closed_loop(mods):
    """controllers.LinearQuadraticRegulator (reference controllers/lqr.py:21-51) on the drop-in StateSpace == the gain of the literal
    restatement; planners.JointPositionPlanner's plan feeds the planner-driven kernels directly."""
    for k in [k for k in sys.modules if k.split(".")[0] in ("controllers", "planners")]:
        del sys.modules[k]
    import controllers
    import planners

    from oracle import replay_oracle as ro
    from rigid_body_manipulation_b200 import model as pm
    from rigid_body_manipulation_b200.engine import Model

    assert os.path.realpath(controllers.__file__).startswith(os.path.realpath(DROPIN))
    c = pm.load_packaged("sequential", "hammer")
    state = SimpleNamespace(qpos=c.key_qpos.copy(), qvel=np.zeros(6), ctrl=np.zeros(6))
    gains = [10.0, 10.0, 10.0, 1e4, 1e4, 1e4]                                   # configurations/base.yaml
    ctl = controllers.LinearQuadraticRegulator(controllers.LinearQuadraticRegulatorConfig(input_gain=gains), c, state)
    consts = dict(hposes_Rt=c.hposes_Rt, simats=c.simats, uscrews=c.uscrews, twist_0=c.twist_0, dtwist_0=c.dtwist_0)
    Kr = ro.lqr_gain(consts, c.key_qpos, np.zeros(6), gains)
    assert ctl.gain_matrix.shape == (6, 12) and np.abs(ctl.gain_matrix - Kr).max() < 1e-4 * np.abs(Kr).max()
    dflt = controllers.LinearQuadraticRegulator(controllers.LinearQuadraticRegulatorConfig(), c, state)      # input_gain missing -> ones(nu)
    assert dflt.input_gain == [1.0] * 6 and dflt.gain_matrix.shape == (6, 12)
    # planner drop-in -> planner-driven kernel: tau of every planned step without materialising the trajectory
    pl = planners.JointPositionPlanner(planners.JointPositionPlannerConfig(duration=3.0, displacements=[0.2, 1.4, 0.6, "pi", 0.0, "6 * pi"]), None, state)
    m = Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0, pose_sen_llj=c.pose_sen_Rt)
    tau = m.rnea_planned(pl.plan).t().cpu().numpy()
    g = load_golden("ref_config1_hammer.npz")
    assert np.abs(tau - g["tau"]).max() < 1e-9 * np.abs(g["tau"]).max()
