"""`not gpu`: the N > 1 host logic on CPU with the gloo backend, world_size 2 -- shard, per-rank Gram pack, one sum
all-reduce of the 112-double pack, identical solve on every rank."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import rnea_vec as rv
        from rigid_body_manipulation_b200 import distributed, identification

        rng = np.random.default_rng(7)  # same stream on every rank: the global batch
        V, dV = rng.standard_normal((n_total, 6)), rng.standard_normal((n_total, 6)) * 2
        phi = np.arange(1, 11) * 0.01
        Y = rv.regressor_batched(V, dV)
        f = Y @ phi
        a, b = distributed.shard_range(n_total, rank, world)
        pack = torch.as_tensor(rv.gram_pack(Y[a:b], f[a:b]))  # stands in for the device kernel's per-rank pack
        distributed.allreduce_gram(pack)
        ident = identification.solve(pack)
        t = distributed.max_over_ranks(float(rank + 1), device="cpu")
        q.put((rank, pack.numpy().copy(), ident.phi, t, (a, b)))
    finally:
        dist.destroy_process_group()


def test_sharded_gram_allreduce_world2():
    from oracle import rnea_vec as rv

    world, n_total = 2, 1001
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(7)
    V, dV = rng.standard_normal((n_total, 6)), rng.standard_normal((n_total, 6)) * 2
    Y = rv.regressor_batched(V, dV)
    phi = np.arange(1, 11) * 0.01
    whole = rv.gram_pack(Y, Y @ phi)
    assert res[0][4] == (0, 501) and res[1][4] == (501, 1001)
    for rank, pack, phi_hat, tmax, _ in res:
        assert np.allclose(pack, whole, rtol=1e-12, atol=1e-12)
        assert pack[111] == n_total
        assert np.abs(phi_hat - phi).max() < 1e-9
        assert tmax == 2.0
    assert np.array_equal(res[0][1], res[1][1])  # every rank holds the same reduced pack
