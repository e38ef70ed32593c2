"""`not gpu`: host-side logic -- the MuJoCo-free model front-end, the planner, the identification solve and sharding."""
import numpy as np
import pytest

from conftest import load_golden
from rigid_body_manipulation_b200 import distributed, identification, model, planner

TARGETS = ["hammer", "uniform_gearbox", "kill_la_kill"]


@pytest.mark.parametrize("t", TARGETS)
def test_packaged_model_reproduces_the_reference_setup_constants(t):
    c = model.load_packaged("sequential", t)
    g = load_golden(f"ref_inverse_{t}.npz")  # constants derived with the reference's own recipe (oracle/model_oracle.py)
    assert np.abs(c.hposes_Rt - g["hposes_Rt"]).max() < 1e-15
    assert np.abs(c.simats - g["simats"]).max() < 1e-13 * np.abs(g["simats"]).max()
    assert np.array_equal(c.uscrews, g["uscrews"])
    assert np.array_equal(c.dtwist_0, g["dtwist_0"]) and not c.twist_0.any()
    assert np.abs(c.pose_sen_Rt - g["pose_sen_llj"]).max() < 1e-15
    assert np.abs(c.pose_sen_obji_Rt - g["pose_sen_obji"]).max() < 1e-15
    assert np.abs(c.simat_object_llj - g["simat_sen_obj"]).max() < 1e-13 * np.abs(g["simats"]).max()
    assert np.allclose(c.target.global_inertia, g["gt_globalinertia"], rtol=1e-14)
    assert np.allclose(c.target.diaginertia, g["gt_diaginertia"], rtol=1e-14)
    assert np.array_equal(c.key_qpos, [1, 1, 1, 0, 0, 0])


def test_all_packaged_targets_build():
    rows = model.packaged_targets()
    assert len(rows) == 23
    for name in rows:
        c = model.load_packaged("sequential", name)
        G = c.simats[6]
        assert np.allclose(G, G.T, atol=1e-12 * np.abs(G).max())
        assert np.linalg.eigvalsh(0.5 * (G + G.T)).min() > 0
        assert abs(G[0, 0] - (8.0 + c.target.mass)) < 1e-12
    with pytest.raises(ValueError):
        model.packaged_target("uniform123_128")  # named by configurations/uniform.yaml:2 but absent from the reference


def test_mjcf_parser_matches_packaged_description(tmp_path):
    xml = """<mujoco><worldbody>
      <body name="l1" euler="0 90 0"><joint name="x" type="slide" axis="0 0 1"/><inertial pos="0 0 0" mass="8." diaginertia=".05 .05 .05"/>
        <body name="l2" pos="0.1 0 0.2" euler="-90 0 0"><joint name="r" type="hinge" axis="0 0 1" pos="0 0.05 0"/><inertial pos="0.01 0 0" mass="2." diaginertia=".01 .02 .03"/>
          <site name="attachment"/></body></body></worldbody><keyframe><key qpos="0.5 0.1"/></keyframe></mujoco>"""
    p = tmp_path / "r.xml"
    p.write_text(xml)
    rob = model.load_mjcf(str(p))
    c = model.build_constants(rob, None)
    assert c.hposes_Rt.shape == (3, 12) and c.uscrews.tolist() == [[0, 0, 1, 0, 0, 0], [0, 0, 0, 0, 0, 1]]
    # joint offset + CoM offset show up as first moments of link 2 about its joint frame: h = m (c - p_joint)
    h = np.array([c.simats[2][5, 1], c.simats[2][3, 2], c.simats[2][4, 0]])
    assert np.allclose(h, 2.0 * (np.array([0.01, 0, 0]) - np.array([0, 0.05, 0])))


def test_planner_matches_reference_golden():
    g = load_golden("ref_planner.npz")
    for name in ("base", "uniform"):
        plan = planner.traj_5th_spline(g[f"{name}_disp"], [1, 1, 1, 0, 0, 0], 0.002, int(g[f"{name}_n_steps"]))
        ref = g[f"{name}_traj"]
        got = plan.trajectory()
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() < 1e-12 * np.abs(ref).max()
        assert np.allclose(plan(17), ref[17], rtol=1e-12, atol=1e-12)
        assert np.allclose(got[0, 0], [1, 1, 1, 0, 0, 0]) and np.abs(got[0, 1:]).max() < 1e-9  # rest-to-rest start


def test_identification_solve_equals_lstsq():
    from oracle import rnea_vec as rv

    rng = np.random.default_rng(0)
    V, dV = rng.standard_normal((500, 6)), rng.standard_normal((500, 6)) * 3
    Y = rv.regressor_batched(V, dV)
    phi = np.array([1.1, 0.1, -0.05, 0.2, 0.02, 0.03, 0.025, 0.001, -0.002, 0.0015])
    f = Y @ phi + rng.standard_normal((500, 6)) * 0.01
    ident = identification.solve(rv.gram_pack(Y, f))
    ref = rv.identify_lstsq(Y, f)  # reference loggers.py:129
    assert np.abs(ident.phi - ref).max() < 1e-10
    assert ident.rank == 10 and ident.n_samples == 500
    assert abs(ident.residual_ss - np.sum((Y.reshape(-1, 10) @ ref - f.reshape(-1)) ** 2)) < 1e-8
    # rank-deficient excitation (no rotation at all): minimum-norm answer like lstsq
    V0, dV0 = np.zeros((50, 6)), np.zeros((50, 6))
    dV0[:, :3] = rng.standard_normal((50, 3))
    Y0 = rv.regressor_batched(V0, dV0)
    f0 = Y0 @ phi
    i0 = identification.solve(rv.gram_pack(Y0, f0))
    assert i0.rank < 10 and np.abs(Y0.reshape(-1, 10) @ i0.phi - f0.reshape(-1)).max() < 1e-9
    # reference score (main.py:21-38): zero for the exact parameters
    assert identification.score(phi, phi, 0.2) == 0.0
    assert identification.score(phi * 1.1, phi, 0.2) > 0


def test_sensor_frame_ground_truth_is_consistent_with_the_folded_inertia():
    """phi of the object in the sensor frame == what the regressor identifies: Y phi == G_obj dV - ad^T G_obj V there."""
    from oracle import rnea_vec as rv

    c = model.load_packaged("sequential", "kill_la_kill")
    # use the reference-recipe inertia (principal off-diagonals dropped) moved into the sensor frame
    Rt = c.pose_sen_Rt
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = Rt[:9].reshape(3, 3), Rt[9:]
    G_sen = model.move_inertia(T, c.simat_object_llj)  # T_sen,llj moves {llj} quantities into {sen}
    m = G_sen[0, 0]
    h = np.array([G_sen[5, 1], G_sen[3, 2], G_sen[4, 0]])
    I = G_sen[3:, 3:]
    phi = np.array([m, *h, I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[1, 2], I[2, 0]])
    rng = np.random.default_rng(1)
    V, dV = rng.standard_normal((5, 6)), rng.standard_normal((5, 6))
    Y = rv.regressor_batched(V, dV)
    for k in range(5):
        rhs = G_sen @ dV[k] - rv.curlywedge(V[k][None])[0].T @ G_sen @ V[k]
        assert np.allclose(Y[k] @ phi, rhs, atol=1e-12)
    assert abs(phi[0] - c.target.mass) < 1e-12


def test_shard_ranges_partition_the_batch():
    for n, w in [(10, 3), (100_000_000, 8), (5, 8), (0, 2)]:
        spans = [distributed.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        distributed.shard_range(10, 3, 3)


def test_split_and_noise_follow_the_reference_recipes():
    """loggers.py:82-108 (`_split`, including its test == valid slice) and simulate.py:281-290 (5 % noise, seed 0)."""
    from operator import itemgetter

    n = 151  # base.yaml: 50 fps x 3 s + 1 frames
    tr, va, te = identification.split_indices(n)
    # literal restatement of the reference method on a list
    data = list(range(n))
    rng = np.random.default_rng(0)
    idx = list(range(n))
    rng.shuffle(idx)
    num_test, num_valid = int(n * 0.1), int(n * 0.1)
    num_train = n - num_test - num_valid
    assert list(tr) == list(itemgetter(*idx[:num_train])(data))
    assert list(va) == list(itemgetter(*idx[num_train : num_train + num_valid])(data))
    assert list(te) == list(va)  # the reference's quirk
    assert len(set(tr) & set(va)) == 0 and len(tr) + len(va) + num_test == n
    f = np.random.default_rng(3).standard_normal((n, 6)) * np.array([10, 10, 10, 1, 1, 1.0])
    g = identification.perturb_wrench(f)
    r = np.random.default_rng(0)
    fs = 0.05 * np.linalg.norm(f[:, :3], axis=1).max()
    ts = 0.05 * np.linalg.norm(f[:, 3:], axis=1).max()
    exp = f.copy()
    exp[:, :3] += fs * r.standard_normal((n, 3))
    exp[:, 3:] += ts * r.standard_normal((n, 3))
    assert np.array_equal(g, exp) and not np.shares_memory(g, f)
