"""Independent cross-check standing in for MuJoCo's mj_inverse (absent): joint forces from an Euler-Lagrange derivation in
world coordinates (oracle/lagrangian.py) vs the reference-produced golden tau."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import lagrangian as lg


@pytest.mark.parametrize("target", ["hammer", "uniform_gearbox", "kill_la_kill"])
def test_lagrangian_agrees_with_reference_rnea(target):
    g = load_golden(f"ref_inverse_{target}.npz")
    consts = dict(hposes_Rt=g["hposes_Rt"], simats=g["simats"], uscrews=g["uscrews"], dtwist_0=g["dtwist_0"])
    for s in [1, 5, 6, 7, 20, 21]:
        q, qd, qdd = g["traj"][s]
        tau = lg.lagrangian_tau(consts, q, qd, qdd)
        ref = g["tau"][s]
        assert np.abs(tau - ref).max() < 2e-8 * max(1.0, np.abs(ref).max()), (s, tau, ref)


# ---------------------------------------------------------------------------------------------------------------------
# the MJCF written by the product's exporter describes the same mechanical system as the Newton-Euler constants
# (checked through M(q), U(q) computed from the XML text alone -- MuJoCo itself is not available here)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("target", ["hammer", "kill_la_kill", None])
def test_exported_mjcf_matches_newton_euler(target):
    from oracle import build_c
    from rigid_body_manipulation_b200 import model as pm
    from rigid_body_manipulation_b200.mjcf_export import to_mjcf

    robot = pm.packaged_robot("sequential")
    tgt = pm.packaged_target(target) if target else None
    c = pm.build_constants(robot, tgt)
    MU = lg.mjcf_mass_matrix_and_potential(to_mjcf(robot, tgt))
    rng = np.random.default_rng(5)
    for _ in range(4):
        q, qd, qdd = rng.uniform(-2, 2, 6), rng.standard_normal(6), rng.standard_normal(6) * 3
        tau_l = lg.lagrangian_tau(MU, q, qd, qdd)
        tau = build_c.inverse_batched_c(np.stack([q, qd, qdd])[None], c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0)[0]
        assert np.abs(tau_l - tau).max() < 2e-7 * max(1.0, np.abs(tau).max()), (tau_l, tau)
