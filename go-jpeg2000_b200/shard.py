"""shard.py -- how a batch is split across GPUs: units (frames or tiles) are independent, so each
rank takes a disjoint subset and NO data-path collective exists.  The only communication is the control-plane
barrier and the max-over-ranks of the device time that bench.py reports.

Units are balanced by a cost (compressed bytes by default: entropy decode time is roughly proportional to them)
with the longest-processing-time-first greedy rule, deterministically, so every rank computes the same plan
without talking to the others.
"""


def shard_units(costs, world, rank=None):
    """costs: sequence of non-negative numbers, one per unit.  Returns the list of unit-index lists per rank
    (or only `rank`'s list).  Every unit is assigned exactly once; ties are broken by index so that the plan is
    identical on all ranks."""
    if world < 1:
        raise ValueError("world must be >= 1")
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0] * world
    plan = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        plan[r].append(i)
        loads[r] += costs[i]
    for p in plan:
        p.sort()
    return plan if rank is None else plan[rank]


def reduce_max_time(ms, group=None):
    """max over ranks of a per-rank device time in milliseconds (torch.distributed, any backend)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(ms)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(ms)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])
