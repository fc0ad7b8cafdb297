"""CPU, world_size 2 over gloo: the multi-GPU plan (disjoint shards, no data-path collective) and the
max-over-ranks timing reduction used by bench.py."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_package
    load_package()
    from go_jpeg2000_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = [(i * 7919) % 101 + 1 for i in range(37)]            # e.g. compressed bytes per frame
    mine = shard.shard_units(costs, world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)                        # test-only check; the data path never gathers
    t = shard.reduce_max_time(10.0 + 5.0 * rank)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, mine, gathered, t, sum(costs[i] for i in mine), sum(costs)))


def test_two_rank_sharding_and_timing():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    (r0, a, g0, t0, l0, tot), (r1, b, g1, t1, l1, _) = res
    assert g0 == g1 == [a, b]                                     # both ranks derived the same global plan
    assert sorted(a + b) == list(range(37)) and not set(a) & set(b)
    assert abs(l0 - l1) <= 101 and l0 + l1 == tot                 # balanced within one unit's cost
    assert t0 == t1 == 15.0                                       # max over ranks


def test_shard_units_edge_cases():
    sys.path.insert(0, ROOT)
    from conftest import load_package
    load_package()
    from go_jpeg2000_b200 import shard
    assert shard.shard_units([], 4) == [[], [], [], []]
    assert shard.shard_units([5], 3) == [[0], [], []]
    plan = shard.shard_units([1] * 8, 8)
    assert sorted(sum(plan, [])) == list(range(8)) and all(len(p) == 1 for p in plan)
    with pytest.raises(ValueError):
        shard.shard_units([1], 0)
