"""-m gpu, needs >= 2 GPUs (skipped otherwise): one process per GPU, samples sharded, per-rank fused Gram, ONE all-reduce of
the 112-double pack -- through torch.distributed (NCCL) and through the library's own rbm_allreduce_gram -- then the
same solve on every rank; tau needs no communication."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from rigid_body_manipulation_b200 import distributed, identification, model
        from rigid_body_manipulation_b200.engine import Model

        c = model.load_packaged("sequential", "uniform_gearbox")
        m = Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0, pose_sen_llj=c.pose_sen_Rt, device=rank)
        rng = np.random.default_rng(99)  # the global batch, identical on every rank
        q = np.concatenate([rng.uniform(-1.5, 2.5, (3, n_total)), rng.uniform(-19, 19, (3, n_total))])
        qd, qdd = rng.standard_normal((6, n_total)), rng.standard_normal((6, n_total)) * 3
        phi = np.array([0.585, -0.0032, 1.9e-5, -8e-6, 3.85e-3, 2.9e-3, 3.0e-3, 1e-5, 2e-5, -1e-5])
        a, b = distributed.shard_range(n_total, rank, world)
        dev = lambda x: torch.as_tensor(np.ascontiguousarray(x[:, a:b]), device="cuda")
        qs, qds, qdds = dev(q), dev(qd), dev(qdd)
        f = m.regressor_from_traj(qs, qds, qdds, want_rows=False, phi=phi)["wrench"]
        pack_t = m.regressor_gram(qs, qds, qdds, f).clone()
        pack_c = pack_t.clone()
        distributed.allreduce_gram(pack_t)                      # torch.distributed / NCCL
        reducer = distributed.NcclGramReducer(device=rank)      # the C ABI's own collective
        reducer(pack_c)
        torch.cuda.synchronize()
        ident = identification.solve(pack_c)
        tau = m.rnea(qs, qds, qdds)
        out.put((rank, pack_t.cpu().numpy(), pack_c.cpu().numpy(), ident.phi, float(tau.abs().sum().item()), (a, b)))
        reducer.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_identification():
    import torch.multiprocessing as mp

    from rigid_body_manipulation_b200 import model
    from rigid_body_manipulation_b200.engine import Model

    world, n_total = 2, 400_001
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([out.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single-GPU reference of the whole batch
    c = model.load_packaged("sequential", "uniform_gearbox")
    m = Model(c.hposes_Rt, c.simats, c.uscrews, c.twist_0, c.dtwist_0, pose_sen_llj=c.pose_sen_Rt, device=0)
    rng = np.random.default_rng(99)
    q = np.concatenate([rng.uniform(-1.5, 2.5, (3, n_total)), rng.uniform(-19, 19, (3, n_total))])
    qd, qdd = rng.standard_normal((6, n_total)), rng.standard_normal((6, n_total)) * 3
    phi = np.array([0.585, -0.0032, 1.9e-5, -8e-6, 3.85e-3, 2.9e-3, 3.0e-3, 1e-5, 2e-5, -1e-5])
    dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), device="cuda:0")
    qs, qds, qdds = dev(q), dev(qd), dev(qdd)
    f = m.regressor_from_traj(qs, qds, qdds, want_rows=False, phi=phi)["wrench"]
    whole = m.regressor_gram(qs, qds, qdds, f).cpu().numpy()
    tau_sum = float(m.rnea(qs, qds, qdds).abs().sum().item())
    assert res[0][5] == (0, 200_001) and res[1][5] == (200_001, 400_001)
    for rank, pack_t, pack_c, phi_hat, _, _ in res:
        assert np.array_equal(pack_t, pack_c)  # both collectives give the same bits
        assert np.abs(pack_t - whole).max() < 1e-11 * np.abs(whole).max()
        assert pack_t[111] == n_total
        assert np.abs(phi_hat - phi).max() < 1e-9
    assert np.array_equal(res[0][1], res[1][1])
    assert abs(res[0][4] + res[1][4] - tau_sum) < 1e-9 * tau_sum
