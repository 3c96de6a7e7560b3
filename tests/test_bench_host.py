"""Host-side logic of bench.py (no GPU): workload selection per rank, config block, ncu record lookup."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_weak_and_strong_sharding(world):
    """Weak: every rank owns the config's per-GPU share with GLOBAL chain ids (C5: 512 per GPU); strong: the config's
    total (C5: 4096) is split into contiguous, balanced, disjoint blocks."""
    seen = []
    for rank in range(world):
        cfg, ids, total = bench.get_cfg("C5", None, "weak", world, rank)
        assert cfg.chains == 512 and total == 512 * world and list(ids) == list(range(rank * 512, (rank + 1) * 512))
        cfg, ids, total = bench.get_cfg("C5", None, "strong", world, rank)
        assert total == 4096 and cfg.chains == len(ids) == 4096 // world
        seen += list(ids)
    assert seen == list(range(4096))
    cfg, ids, total = bench.get_cfg("C2", None, "weak", world, 0)
    assert cfg.chains == 64 and total == 64 * world and cfg.N == 256 and cfg.L == 200 and cfg.n_stab == 10


def test_strong_sharding_uneven():
    sizes = [len(bench.get_cfg("C2", 10, "strong", 4, r)[1]) for r in range(4)]
    assert sizes == [3, 3, 2, 2]


def test_config_block_names_the_workload():
    cfg, ids, total = bench.get_cfg("C3", None, "weak", 2, 1)
    c = bench.config_json(cfg, 2, None, total, "weak")
    assert c["eltype"] == "ComplexF64" and c["inverse"] == "inv_IpUV!" and c["chains_total"] == 256
    assert c["N"] == 256 and "C3" in c["workload"] and "model" not in c
    assert c["flops_per_eval"] > 0 and c["ldr_pivoting"].startswith("geqp3")
    json.dumps(c)


def test_ncu_records_are_keyed_by_kernel_and_config():
    rec = bench.ncu_record("gemm_kernel", "C2")
    assert rec and rec["dram_bytes_per_launch"] > 0 and "commit" in rec and os.path.exists(os.path.join(ROOT, rec["source"].split(" ")[0]))
    assert bench.ncu_record("gemm_kernel", "no-such-config") is None


def test_default_steps_cover_every_config():
    import sla_b200
    assert set(bench.DEFAULT_STEPS) == set(sla_b200.synthetic.CONFIGS)
