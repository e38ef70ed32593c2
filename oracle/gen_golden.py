"""ORACLE / TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN FILES (unmodified, from
/root/reference, through oracle/reference_loader.py) on seeded inputs.  Run in the build
container (the reference checkout does not exist on the GPU box):

    python -m oracle.gen_golden

The fixtures are what pins oracle/rnea_oracle.py, oracle/rnea_vec.py, oracle/rnea_oracle.c
and, on the GPU, the CUDA kernels.  Model constants come from oracle/model_oracle.py
(MuJoCo stand-in) evaluated on the reference's own `sequential.xml` and CAD CSV rows.
"""
from __future__ import annotations

import os

import numpy as np

from . import model_oracle as mo
from . import reference_loader as rl

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
XML = os.path.join(rl.REFERENCE_ROOT, "xml_models")


def sample_states(rng, n):
    """SURVEY.md 8(d) config-2 input distribution."""
    q = np.concatenate([rng.uniform(-1.5, 2.5, (n, 3)), rng.uniform(-6 * np.pi, 6 * np.pi, (n, 3))], axis=1)
    qd = rng.standard_normal((n, 6)) * np.array([1, 1, 1, 3, 3, 3.0])
    qdd = rng.standard_normal((n, 6)) * np.array([3, 3, 3, 10, 10, 10.0])
    return np.stack([q, qd, qdd], axis=1)  # (n, 3, 6)


def to_ref_se3(ns, R, t):
    return ns.liegroups.SE3(ns.liegroups.SO3(np.array(R, float)), np.array(t, float))


def run_reference_inverse(ns, trajs, hposes_Rt, simats, uscrews, twist_0, dtwist_0, wrench_tip=None, pose_tip=None):
    nj = len(uscrews)
    hposes = [to_ref_se3(ns, h[:9].reshape(3, 3), h[9:]) for h in hposes_Rt]
    kw = {}
    if wrench_tip is not None:
        kw["wrench_tip"] = wrench_tip
    if pose_tip is not None:
        kw["pose_tip_ee"] = to_ref_se3(ns, pose_tip[:9].reshape(3, 3), pose_tip[9:])
    n = len(trajs)
    tau = np.zeros((n, nj))
    poses = np.zeros((n, nj, 12))
    tw = np.zeros((n, nj + 1, 6))
    dtw = np.zeros((n, nj + 1, 6))
    for s in range(n):
        t_, p_, v_, dv_ = ns.dynamics.inverse(trajs[s], hposes, simats, uscrews, twist_0, dtwist_0, **kw)
        tau[s] = t_
        for i in range(nj):
            poses[s, i, :9] = p_[i].rot.as_matrix().reshape(9)
            poses[s, i, 9:] = p_[i].trans
        tw[s] = np.array(v_)
        dtw[s] = np.array(dv_)
    return tau, poses, tw, dtw


def sensor_and_regressor(ns, sen_Rt, tw6, dtw6):
    """reference core/simulate.py:202-209,223-224 executed with the reference's liegroups + dynamics."""
    SE3 = ns.liegroups.SE3
    pose = to_ref_se3(ns, sen_Rt[:9].reshape(3, 3), sen_Rt[9:])
    n = len(tw6)
    tws, dtws, Y = np.zeros((n, 6)), np.zeros((n, 6)), np.zeros((n, 6, 10))
    for s in range(n):
        twist_sen = pose.adjoint() @ tw6[s]
        dAd = SE3.curlywedge(twist_sen) @ pose.adjoint()
        dtwist_sen = dAd @ tw6[s] + pose.adjoint() @ dtw6[s]
        tws[s], dtws[s] = twist_sen, dtwist_sen
        Y[s] = ns.dynamics.get_regressor_matrix(twist_sen, dtwist_sen)
    return tws, dtws, Y


def pose_Rt(p):
    return np.concatenate([np.asarray(p.rot.as_matrix(), float).reshape(9), np.asarray(p.trans, float)])


def golden_for_target(ns, name, n, seed):
    c = mo.build_constants(os.path.join(XML, "manipulators", "sequential.xml"), os.path.join(XML, "targets", name, "object_cad_gt.csv"))
    rng = np.random.default_rng(seed)
    trajs = sample_states(rng, n)
    # a few deliberate edge states: keyframe at rest, all zeros, tiny angles on both sides of
    # liegroups' isclose(angle, 0) small-angle branch, large angles
    trajs[0] = 0.0
    trajs[1] = 0.0
    trajs[1, 0] = c.key_qpos
    trajs[2, 0, 3:] = [1e-9, -5e-9, 9.9e-9]
    trajs[3, 0, 3:] = [2e-8, -1e-7, 1e-6]
    trajs[4, 0, 3:] = [6 * np.pi, -6 * np.pi, 100.0]
    hRt = c.hposes_Rt()
    tau, poses, tw, dtw = run_reference_inverse(ns, trajs, hRt, c.simats, c.uscrews, c.twist_0, c.dtwist_0)
    sen = pose_Rt(c.pose_sen_llj)
    tws, dtws, Y = sensor_and_regressor(ns, sen, tw[:, 6], dtw[:, 6])
    gt = c.ground_truth
    np.savez_compressed(
        os.path.join(GOLD, f"ref_inverse_{name}.npz"),
        traj=trajs, tau=tau, poses=poses, twists=tw, dtwists=dtw,
        twist_sen=tws, dtwist_sen=dtws, regressor=Y,
        hposes_Rt=hRt, simats=c.simats, uscrews=c.uscrews, twist_0=c.twist_0, dtwist_0=c.dtwist_0,
        pose_sen_llj=sen, pose_sen_obj=pose_Rt(c.pose_sen_obj), pose_sen_obji=pose_Rt(c.pose_sen_obji),
        simat_sen_obj=c.simat_sen_obj, key_qpos=c.key_qpos,
        gt_mass=gt["mass"], gt_com=gt["com"], gt_iquat=gt["iquat"], gt_diaginertia=gt["diaginertia"],
        gt_fullinertia=np.array(gt["fullinertia"]), gt_globalinertia=np.array(gt["globalinertia"]),
        gt_aabb_scale=gt["aabb_scale"],
    )
    return c


def random_rotation(rng):
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (w * y + x * z)],
            [2 * (w * z + x * y), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (w * x + y * z), 1 - 2 * (x * x + y * y)],
        ]
    )


def golden_generic(ns, n, seed, nj, tag):
    """A model with NO special structure: random home poses with translations, dense SPD spatial
    inertias, general (not one-hot) unit screws mixing translation and rotation, non-zero base
    twist, tip wrench and tip pose -- exercises every term of reference dynamics.py:109-157."""
    rng = np.random.default_rng(seed)
    hRt = np.zeros((nj + 1, 12))
    hRt[0, [0, 4, 8]] = 1.0
    for k in range(1, nj + 1):
        hRt[k, :9] = random_rotation(rng).reshape(9)
        hRt[k, 9:] = rng.uniform(-0.5, 0.5, 3)
    simats = np.zeros((nj + 1, 6, 6))
    for k in range(1, nj + 1):
        A = rng.standard_normal((6, 6))
        simats[k] = A @ A.T + 0.5 * np.eye(6)
    uscrews = np.zeros((nj, 6))
    for k in range(nj):
        kind = k % 3
        if kind == 0:  # pure rotation about a random unit axis
            a = rng.standard_normal(3)
            uscrews[k, 3:] = a / np.linalg.norm(a)
        elif kind == 1:  # pure translation along a random unit axis
            a = rng.standard_normal(3)
            uscrews[k, :3] = a / np.linalg.norm(a)
        else:  # general screw: unit rotation axis + linear part (axis not through the origin, pitch != 0)
            a = rng.standard_normal(3)
            uscrews[k, 3:] = a / np.linalg.norm(a)
            uscrews[k, :3] = rng.uniform(-0.7, 0.7, 3)
    twist_0 = rng.standard_normal(6) * 0.3
    dtwist_0 = np.array([0.3, -0.2, 9.81, 0.1, 0.05, -0.07])
    wrench_tip = rng.standard_normal(6)
    pose_tip = np.concatenate([random_rotation(rng).reshape(9), rng.uniform(-0.3, 0.3, 3)])
    trajs = np.stack([rng.uniform(-3, 3, (n, nj)), rng.standard_normal((n, nj)), rng.standard_normal((n, nj)) * 3], axis=1)
    trajs[0] = 0.0
    trajs[1, 0] = 1e-9  # below liegroups' small-angle threshold for unit axes
    tau, poses, tw, dtw = run_reference_inverse(ns, trajs, hRt, simats, uscrews, twist_0, dtwist_0, wrench_tip, pose_tip)
    sen = np.concatenate([random_rotation(rng).reshape(9), rng.uniform(-0.2, 0.2, 3)])
    tws, dtws, Y = sensor_and_regressor(ns, sen, tw[:, nj], dtw[:, nj])
    np.savez_compressed(
        os.path.join(GOLD, f"ref_inverse_generic_{tag}.npz"),
        traj=trajs, tau=tau, poses=poses, twists=tw, dtwists=dtw, twist_sen=tws, dtwist_sen=dtws, regressor=Y,
        hposes_Rt=hRt, simats=simats, uscrews=uscrews, twist_0=twist_0, dtwist_0=dtwist_0,
        wrench_tip=wrench_tip, pose_tip=pose_tip, pose_sen_llj=sen,
    )


def golden_setup_functions(ns, seed):
    """transfer_simat / get_spatial_inertia_matrix / coordinate_transfer_* / compose / homogenize /
    extract_lin{vel,acc}_frame_transferred executed by the reference on seeded inputs."""
    rng = np.random.default_rng(seed)
    dyn, tf = ns.dynamics, ns.transformations
    n = 16
    Rt = np.array([np.concatenate([random_rotation(rng).reshape(9), rng.uniform(-1, 1, 3)]) for _ in range(n)])
    mass = rng.uniform(0.1, 10, n)
    diag = rng.uniform(0.01, 1.0, (n, 3))
    simats_diag = dyn.get_spatial_inertia_matrix(mass, diag)
    poses = [to_ref_se3(ns, r[:9].reshape(3, 3), r[9:]) for r in Rt]
    transferred = dyn.transfer_simat(poses, simats_diag)
    transferred_single = dyn.transfer_simat(poses[3], simats_diag[3])
    dense = np.array([a @ a.T for a in rng.standard_normal((n, 6, 6))])
    transferred_dense = dyn.transfer_simat(poses, dense)
    imats = np.array([a @ a.T for a in rng.standard_normal((n, 3, 3))])
    ct_imat = np.array([dyn.coordinate_transfer_imat(p, im, m) for p, im, m in zip(poses, imats, mass)])
    ct_simat = np.array([dyn.coordinate_transfer_simat(p, g) for p, g in zip(poses, dense)])
    tw = rng.standard_normal((n, 6))
    dtw = rng.standard_normal((n, 6))
    linvel = np.array([dyn.extract_linvel_frame_transferred(a, p) for a, p in zip(tw, poses)])
    linvel_h = np.array([dyn.extract_linvel_frame_transferred(a, p, homogeneous=True) for a, p in zip(tw, poses)])
    linacc = np.array([dyn.extract_linacc_frame_transferred(a, b, p) for a, b, p in zip(tw, dtw, poses)])
    linacc_h = np.array([dyn.extract_linacc_frame_transferred(a, b, p, homogeneous=True) for a, b, p in zip(tw, dtw, poses)])
    Y = np.array([dyn.get_regressor_matrix(a, b) for a, b in zip(tw, dtw)])
    # compose: quaternion rows, matrix rows, rot=None, single
    quats = rng.standard_normal((n, 4))
    quats /= np.linalg.norm(quats, axis=1, keepdims=True)
    trans = rng.uniform(-1, 1, (n, 3))
    comp_q = np.array([pose_Rt(p) for p in tf.compose(trans, quats)])
    comp_m = np.array([pose_Rt(p) for p in tf.compose(trans, Rt[:, :9].copy())])
    comp_none = np.array([pose_Rt(p) for p in tf.compose(trans)])
    comp_single = pose_Rt(tf.compose(trans[0], quats[0]))
    hom = tf.homogenize(trans[0])
    hom0 = tf.homogenize(trans[1], 0)
    # the reference's own scratch check, test_adjoint_inv_transpose.py:8-27
    SO3, SE3 = ns.liegroups.SO3, ns.liegroups.SE3
    pose = SE3(SO3.from_rpy(10 / 180 * np.pi, 20 / 180 * np.pi, 40 / 180 * np.pi), np.array([1, 2, 3]))
    np.savez_compressed(
        os.path.join(GOLD, "ref_setup_functions.npz"),
        poses_Rt=Rt, mass=mass, diag=diag, simats_diag=simats_diag, transferred=transferred,
        transferred_single=transferred_single, dense=dense, transferred_dense=transferred_dense,
        imats=imats, ct_imat=ct_imat, ct_simat=ct_simat, twists=tw, dtwists=dtw, linvel=linvel, linvel_h=linvel_h,
        linacc=linacc, linacc_h=linacc_h, regressor=Y, quats=quats, trans=trans, comp_q=comp_q, comp_m=comp_m,
        comp_none=comp_none, comp_single=comp_single, hom=hom, hom0=hom0,
        adj_pose_Rt=pose_Rt(pose), adj=pose.adjoint(), adj_inv=pose.inv().adjoint(),
        adj_T_close_inv=np.allclose(pose.adjoint().T, pose.inv().adjoint()),
        adj_inv_close_pinv=np.allclose(pose.inv().adjoint(), np.linalg.pinv(pose.adjoint())),
    )


def golden_planner(ns):
    """reference planners/joint_position_planner.py:86-131 on both shipped configs
    (configurations/base.yaml:18-26, configurations/uniform.yaml:20-28), keyframe offset
    (sequential.xml:81), MuJoCo default timestep 0.002 (joint_position_planner.py:41-42)."""
    out = {}
    cfgs = {
        "base": ([0.2, 1.4, 0.6, 3.141592653589793, 0.0, 18.8495559215], 3.0),
        "uniform": ([0.2, 0.4, 0.6, 3.141592653589793, 0.9424777960769379, 4.71238898038469], 2.0),
    }
    for name, (disp, duration) in cfgs.items():
        dt = 0.002
        n_steps = int(duration / dt)
        plan = ns.planner.traj_5th_spline(disp, [1, 1, 1, 0, 0, 0], dt, n_steps)
        out[f"{name}_traj"] = np.array([plan(s) for s in range(n_steps)])
        out[f"{name}_disp"] = np.array(disp)
        out[f"{name}_n_steps"] = n_steps
    np.savez_compressed(os.path.join(GOLD, "ref_planner.npz"), **out)
    return out


def golden_config1(ns, c, planner_out, name):
    """BASELINE.json config 1, open-loop part: tau / V6 / dV6 / Y for every planned (q, qd, qdd) of the
    base.yaml trajectory (core/simulate.py:187-188); every 10th step is the 50 fps frame grid (:196)."""
    trajs = planner_out["base_traj"]
    hRt = c.hposes_Rt()
    tau, poses, tw, dtw = run_reference_inverse(ns, trajs, hRt, c.simats, c.uscrews, c.twist_0, c.dtwist_0)
    tws, dtws, Y = sensor_and_regressor(ns, pose_Rt(c.pose_sen_llj), tw[:, 6], dtw[:, 6])
    np.savez_compressed(
        os.path.join(GOLD, f"ref_config1_{name}.npz"),
        traj=trajs, tau=tau, twist6=tw[:, 6], dtwist6=dtw[:, 6], twist_sen=tws, dtwist_sen=dtws,
        regressor_frames=Y[::10], frame_steps=np.arange(0, len(trajs), 10),
    )


def golden_mjmodel(name):
    """The MjModel / MjData look-alike arrays (oracle/mjcf_subset.py) for one target, so that the GPU box -- which has
    neither MuJoCo nor the reference's XML -- can exercise the product's MuJoCo bridge (constants_from_mujoco, Poses)."""
    from . import mjcf_subset as mj

    gt = mj.target_ground_truth(mj.read_cad_row(os.path.join(XML, "targets", name, "object_cad_gt.csv")))
    m = mj.compile_manipulator_with_target(os.path.join(XML, "manipulators", "sequential.xml"), gt)
    d = mj.kinematics(m, m.key_qpos)
    np.savez_compressed(
        os.path.join(GOLD, f"ref_mjmodel_{name}.npz"),
        body_names=np.array(m.body_names), site_names=np.array(m.site_names), body_pos=m.body_pos, body_quat=m.body_quat,
        body_ipos=m.body_ipos, body_iquat=m.body_iquat, body_mass=m.body_mass, body_inertia=m.body_inertia, jnt_type=m.jnt_type,
        jnt_axis=m.jnt_axis, jnt_pos=m.jnt_pos, gravity=m.gravity, key_qpos=m.key_qpos, qpos=d.qpos, xpos=d.xpos, xmat=d.xmat,
        xipos=d.xipos, ximat=d.ximat, site_xpos=d.site_xpos, site_xmat=d.site_xmat,
    )


def main():
    os.makedirs(GOLD, exist_ok=True)
    ns = rl.load()
    c_h = golden_for_target(ns, "hammer", 192, seed=20261018)
    golden_for_target(ns, "uniform_gearbox", 192, seed=7)
    golden_for_target(ns, "kill_la_kill", 64, seed=11)  # strongly off-axis CoM and principal frame
    golden_generic(ns, 96, seed=3, nj=6, tag="nj6")
    golden_generic(ns, 48, seed=4, nj=4, tag="nj4")
    golden_generic(ns, 48, seed=5, nj=9, tag="nj9")
    golden_setup_functions(ns, seed=99)
    pl = golden_planner(ns)
    golden_config1(ns, c_h, pl, "hammer")
    golden_mjmodel("hammer")
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
