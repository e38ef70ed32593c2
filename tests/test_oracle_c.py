"""Pins the C restatement (oracle/rnea_oracle.c) against the golden vectors produced by the reference's own files."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import build_c

CASES = ["ref_inverse_hammer.npz", "ref_inverse_uniform_gearbox.npz", "ref_inverse_kill_la_kill.npz", "ref_inverse_generic_nj6.npz",
         "ref_inverse_generic_nj4.npz", "ref_inverse_generic_nj9.npz"]


@pytest.mark.parametrize("fname", CASES)
def test_c_oracle_matches_reference_golden(fname):
    g = load_golden(fname)
    kw = dict(wrench_tip=g["wrench_tip"], pose_tip_Rt=g["pose_tip"]) if "wrench_tip" in g.files else {}
    tau, V, dV = build_c.inverse_batched_c(g["traj"], g["hposes_Rt"], g["simats"], g["uscrews"], g["twist_0"], g["dtwist_0"], want_twists=True, **kw)
    nj = g["uscrews"].shape[0]
    scale = np.abs(g["tau"]).max(axis=1, keepdims=True)
    assert (np.abs(tau - g["tau"]) / scale).max() < 1e-13
    assert np.abs(V - g["twists"][:, nj]).max() < 1e-12 * max(1.0, np.abs(g["twists"]).max())
    assert np.abs(dV - g["dtwists"][:, nj]).max() < 1e-12 * max(1.0, np.abs(g["dtwists"]).max())


def test_c_oracle_empty_and_bad_nj():
    g = load_golden("ref_inverse_hammer.npz")
    out = build_c.inverse_batched_c(np.zeros((0, 3, 6)), g["hposes_Rt"], g["simats"], g["uscrews"], g["twist_0"], g["dtwist_0"])
    assert out.shape == (0, 6)
