"""-m gpu: the C ABI called directly (ctypes, raw device pointers) the way a non-torch binding would: padded row pitch
(ld > n), a non-default stream, CUDA-graph capture of the hot call, several models alive at once."""
import ctypes as C

import numpy as np
import pytest
import torch

from gpu_util import load_golden, model_from_golden, rel_err, sample_states
from oracle import build_c
from rigid_body_manipulation_b200 import _lib

pytestmark = pytest.mark.gpu


def vp(t):
    return C.c_void_p(t.data_ptr())


def test_padded_pitch_and_side_stream():
    g = load_golden("ref_inverse_hammer.npz")
    m = model_from_golden(g)
    lib = _lib.load()
    n, ld = 1000, 1536
    traj = sample_states(np.random.default_rng(0), n)
    ref = build_c.inverse_batched_c(traj, g["hposes_Rt"], g["simats"], g["uscrews"], g["twist_0"], g["dtwist_0"])
    bufs = []
    for k in range(3):
        b = torch.full((6, ld), float("nan"), dtype=torch.float64, device="cuda")
        b[:, :n] = torch.as_tensor(traj[:, k, :].T, device="cuda")
        bufs.append(b)
    tau = torch.full((6, ld), -7.0, dtype=torch.float64, device="cuda")
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    rc = lib.rbm_rnea_f64(m._h, vp(bufs[0]), vp(bufs[1]), vp(bufs[2]), vp(tau), None, None, n, ld, C.c_void_p(st.cuda_stream))
    assert rc == 0, _lib.last_error()
    st.synchronize()
    assert rel_err(tau[:, :n].t().cpu().numpy(), ref).max() < 1e-9
    assert (tau[:, n:] == -7.0).all()  # the padding is never touched
    # ld < n is rejected
    assert lib.rbm_rnea_f64(m._h, vp(bufs[0]), vp(bufs[1]), vp(bufs[2]), vp(tau), None, None, n, n - 1, None) == _lib.RBM_ERR_INVALID


def test_cuda_graph_capture_and_replay():
    g = load_golden("ref_inverse_uniform_gearbox.npz")
    m = model_from_golden(g)
    n = 4096
    traj = sample_states(np.random.default_rng(1), n)
    dev = torch.as_tensor(traj, device="cuda")
    q, qd, qdd = (dev[:, k, :].t().contiguous() for k in range(3))
    tau = torch.empty_like(q)
    m.rnea(q, qd, qdd, tau=tau)  # warm-up outside capture
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        m.rnea(q, qd, qdd, tau=tau)
        m.rnea(q, qd, tau, tau=qdd)  # a dependent second launch inside the same graph (stream order must hold under PDL)
    expect1 = m.rnea(q, qd, dev[:, 2, :].t().contiguous())
    expect2 = m.rnea(q, qd, expect1)
    q.add_(0.0)
    qdd.copy_(dev[:, 2, :].t())
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(tau, expect1)
    assert torch.equal(qdd, expect2)


def test_many_models_coexist():
    names = ["hammer", "uniform_gearbox", "kill_la_kill"]
    gs = [load_golden(f"ref_inverse_{t}.npz") for t in names]
    ms = [model_from_golden(g) for g in gs]
    for g, m in zip(gs, ms):
        tau = m.rnea_aos(torch.as_tensor(g["traj"], device="cuda"))
        assert rel_err(tau.cpu().numpy(), g["tau"]).max() < 1e-9
    ms[0].close()
    tau = ms[1].rnea_aos(torch.as_tensor(gs[1]["traj"], device="cuda"))
    assert rel_err(tau.cpu().numpy(), gs[1]["tau"]).max() < 1e-9
