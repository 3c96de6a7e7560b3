"""world_size-2 gloo test (CPU) of the multi-rank host logic: chain sharding and the cross-walker
all-reduce of log-determinants and signs (the only collective on the path)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, total, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import sla_b200
    from sla_b200 import walkers
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = walkers.shard_chains(total, world, rank)
    # stand-in per-chain results that depend only on the global chain id
    logdet = torch.tensor([-(10.0 + i) for i in ids], dtype=torch.float64)
    sign = torch.tensor([np.exp(0.1j * i) for i in ids], dtype=torch.complex128)
    out = walkers.reduce_walkers(logdet, sign)
    # the fields of a chain must not depend on the sharding
    f = sla_b200.synthetic.hs_fields(list(ids), 5, 8)
    q.put((rank, list(ids), out.tolist(), f.sum(axis=(1, 2)).tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_walker_reduction_gloo(sla):
    world, total = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = sum((r[1] for r in res), [])
    assert ids == list(range(total))                         # disjoint cover, contiguous blocks
    assert abs(len(res[0][1]) - len(res[1][1])) <= 1
    expect = [sum(-(10.0 + i) for i in range(total)), sum(np.cos(0.1 * i) for i in range(total)),
              sum(np.sin(0.1 * i) for i in range(total)), float(total)]
    for r in res:
        assert np.allclose(r[2], expect)                     # every rank holds the global sums
    ref = sla.synthetic.hs_fields(list(range(total)), 5, 8).sum(axis=(1, 2)).tolist()
    assert sum((r[3] for r in res), []) == ref


def test_shard_chains_properties(sla):
    from sla_b200 import walkers
    for total in (0, 1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            parts = [walkers.shard_chains(total, world, r) for r in range(world)]
            assert sum((list(p) for p in parts), []) == list(range(total))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert list(walkers.weak_chain_ids(4, 2)) == [8, 9, 10, 11]
    with pytest.raises(ValueError):
        walkers.shard_chains(4, 2, 2)
