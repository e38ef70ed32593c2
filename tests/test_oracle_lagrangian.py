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
