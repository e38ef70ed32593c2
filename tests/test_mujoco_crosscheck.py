"""Cross-check against MuJoCo's own inverse dynamics (the north star's `mj_inverse` comparison).

SKIPPED wherever MuJoCo is not installed -- which includes the build container and the GPU boxes of this project, so this file
has not been executed there; it is provided so that the comparison runs as soon as `pip install mujoco` is possible.  The model
has no damping / armature / friction, so qfrc_inverse must equal the joint forces of the recursive Newton-Euler path."""
import numpy as np
import pytest

mujoco = pytest.importorskip("mujoco")

from oracle import build_c  # noqa: E402
from rigid_body_manipulation_b200 import model as rbm_model  # noqa: E402
from rigid_body_manipulation_b200.mjcf_export import to_mjcf  # noqa: E402


@pytest.mark.parametrize("target", ["hammer", "uniform_gearbox", "kill_la_kill"])
def test_tau_equals_mj_inverse(target):
    robot, tgt = rbm_model.packaged_robot("sequential"), rbm_model.packaged_target(target)
    c = rbm_model.build_constants(robot, tgt)
    m = mujoco.MjModel.from_xml_string(to_mjcf(robot, tgt))
    d = mujoco.MjData(m)
    assert m.nv == 6
    rng = np.random.default_rng(0)
    n = 64
    traj = np.stack([rng.uniform(-2, 2, (n, 6)), rng.standard_normal((n, 6)), rng.standard_normal((n, 6)) * 3], axis=1)
    tau = build_c.inverse_batched_c(traj, c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0)
    for s in range(n):
        d.qpos[:], d.qvel[:], d.qacc[:] = traj[s, 0], traj[s, 1], traj[s, 2]
        mujoco.mj_inverse(m, d)
        assert np.abs(d.qfrc_inverse - tau[s]).max() < 1e-8 * max(1.0, np.abs(tau[s]).max()), (s, d.qfrc_inverse, tau[s])


@pytest.mark.gpu
def test_gpu_tau_equals_mj_inverse():
    import torch

    from rigid_body_manipulation_b200.engine import Model

    robot, tgt = rbm_model.packaged_robot("sequential"), rbm_model.packaged_target("hammer")
    c = rbm_model.build_constants(robot, tgt)
    m = mujoco.MjModel.from_xml_string(to_mjcf(robot, tgt))
    d = mujoco.MjData(m)
    mdl = Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0)
    rng = np.random.default_rng(1)
    traj = np.stack([rng.uniform(-2, 2, (128, 6)), rng.standard_normal((128, 6)), rng.standard_normal((128, 6)) * 3], axis=1)
    tau = mdl.rnea_aos(torch.as_tensor(traj, device="cuda")).cpu().numpy()
    for s in range(128):
        d.qpos[:], d.qvel[:], d.qacc[:] = traj[s, 0], traj[s, 1], traj[s, 2]
        mujoco.mj_inverse(m, d)
        assert np.abs(d.qfrc_inverse - tau[s]).max() < 1e-8 * max(1.0, np.abs(tau[s]).max())
