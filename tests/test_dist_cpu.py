"""Host-side logic of the N>1 path on CPU (gloo, world_size 2): contiguous sharding of independent units,
one broadcast per key, optional result gather.  No GPU: the per-shard compute here is the ORACLE standing
in for the CUDA call (this is a test of the plumbing, not of the kernels)."""
import os
import socket

import numpy as np
import pytest

from fhe_study_b200.dist import shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 5, 8, 13, 4096, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q_out):
    import torch
    import torch.distributed as dist

    import oracle
    from fhe_study_b200.dist import broadcast_key, run_sharded

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        kn, l, total = 16, 64, 13  # ragged: 13 ciphertexts over 2 ranks
        # rank 0 owns the key; everybody else starts with garbage and receives it by broadcast
        if rank == 0:
            ksk = torch.from_numpy(oracle.uniform(1, kn * l * (kn + 1)).view(np.int64).copy())
        else:
            ksk = torch.full((kn * l * (kn + 1),), -1, dtype=torch.int64)
        broadcast_key(ksk, src=0)
        cts = torch.from_numpy(oracle.uniform(2, (total, kn + 1)).view(np.int64).copy())
        ksk_np = ksk.numpy().view(np.uint64)

        def fn(local):
            out = oracle.key_switch(kn, kn, l, ksk_np, local.numpy().view(np.uint64).reshape(-1))
            return torch.from_numpy(out.view(np.int64).reshape(local.shape[0], kn + 1).copy())

        local = run_sharded(cts, kn + 1, fn, gather=False)
        full = run_sharded(cts, kn + 1, fn, gather=True)
        want = oracle.key_switch(kn, kn, l, oracle.uniform(1, kn * l * (kn + 1)), oracle.uniform(2, (total, kn + 1)).reshape(-1))
        want = want.reshape(total, kn + 1)
        b, e = shard_range(total, rank, world)
        ok = bool((local.numpy().view(np.uint64) == want[b:e]).all() and (full.numpy().view(np.uint64) == want).all())
        q_out.put((rank, ok, tuple(local.shape)))
    finally:
        dist.destroy_process_group()


def test_broadcast_shard_gather_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, (7, 17)), (1, True, (6, 17))]
