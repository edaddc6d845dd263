"""Scene sharding across ranks and the final result gather.

Scenes are independent units of the forward path: graph_gather only offsets indices (lanegcn.py:196-203), no
edge or Att pair crosses scenes (:675), GroupNorm is per row.  So a batch of B scenes is cut into
``world`` CONTIGUOUS shards balanced by node count, every rank runs the whole forward on its shard with no
data-path collective, and the per-actor outputs are gathered once at the end (reference analogue: the MPI
allgather of per-rank results, train.py:245-255).  Backend-agnostic: NCCL over NVLink on the GPU box, gloo in
the CPU tests.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


def partition(costs: Sequence[int], world: int) -> List[range]:
    """Contiguous partition of len(costs) scenes into ``world`` ranges with near-equal total cost
    (cost = node count).  Ranges may be empty when there are fewer scenes than ranks."""
    n, total = len(costs), float(sum(costs))
    bounds, acc, j = [0], 0.0, 0
    for r in range(1, world):
        target = total * r / world
        while j < n and acc + costs[j] / 2.0 <= target:
            acc += costs[j]
            j += 1
        bounds.append(j)
    bounds.append(n)
    return [range(bounds[r], bounds[r + 1]) for r in range(world)]


def shard_batch(data: Dict[str, list], rank: int, world: int) -> Dict[str, list]:
    """The slice of a collated batch (dict of per-scene lists, data.py:555-561) that ``rank`` owns."""
    costs = [int(g["num_nodes"]) for g in data["graph"]]
    r = partition(costs, world)[rank]
    return {k: v[r.start: r.stop] for k, v in data.items()}


class GatherPlan:
    """Who owns how many actors per scene, agreed once per batch (host metadata, no device work)."""

    def __init__(self, per_rank: List[List[int]]):
        self.per_rank = per_rank
        self.max_actors = max(1, max(sum(s) for s in per_rank))


def make_plan(local_sizes: Sequence[int], group=None) -> GatherPlan:
    """One all_gather_object of the per-scene actor counts of every rank's shard (staging-time metadata: the
    counts are known on the host before anything runs on the device)."""
    world = dist.get_world_size(group)
    per_rank: List = [None] * world
    dist.all_gather_object(per_rank, [int(x) for x in local_sizes], group=group)
    return GatherPlan(per_rank)


def gather_outputs(out: Dict[str, List[torch.Tensor]], plan: GatherPlan = None, group=None,
                   lazy: bool = False) -> Dict[str, List[torch.Tensor]]:
    """All ranks receive every scene's ``cls``/``reg`` in global scene order with ONE collective: cls [A,K] and
    reg [A,K,T,2] are packed side by side into a [max_actors, K + K*T*2] buffer (1,464 B per actor in the
    reference config) and all_gathered; no host synchronisation when ``plan`` is given."""
    if plan is None:
        plan = make_plan(list(out["cls"].sizes) if getattr(out["cls"], "sizes", None) is not None else [len(x) for x in out["cls"]], group)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = plan.per_rank[rank]
    if len(out["cls"]):
        cat_of = lambda l: l.cat if getattr(l, "cat", None) is not None else torch.cat(list(l), 0)  # noqa: E731
        cls, reg = cat_of(out["cls"]), cat_of(out["reg"])
        k, tail = cls.shape[1], tuple(reg.shape[1:])
        dev = cls.device
    else:  # a rank without scenes still takes part; shapes follow the reference config
        k, tail = 6, (6, 30, 2)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        cls, reg = torch.zeros(0, k, device=dev), torch.zeros((0,) + tail, device=dev)
    width = k + int(torch.tensor(tail).prod())
    buf = torch.zeros(plan.max_actors, width, dtype=torch.float32, device=dev)
    n = sum(mine)
    if n:
        buf[:n, :k] = cls
        buf[:n, k:] = reg.reshape(n, -1)
    full = torch.empty(world * plan.max_actors, width, dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(full, buf, group=group)
    # compact the ranks' valid rows (2 small copies); with lazy=True the per-scene views are created on first use
    # (splitting a 128-scene result into 256 views costs ~0.4 ms of host time per step)
    from .lanegcn import scene_list

    counts = [sum(s) for s in plan.per_rank]
    if all(c == plan.max_actors for c in counts):
        valid = full
    else:
        valid = torch.cat([full[r * plan.max_actors: r * plan.max_actors + c] for r, c in enumerate(counts)], 0)
    sizes = [x for s in plan.per_rank for x in s]
    cls = valid[:, :k].contiguous()
    reg = valid[:, k:].contiguous().view((-1,) + tail)
    off = torch.zeros(1, dtype=torch.int32)
    return {"cls": scene_list(cls, sizes, off, lazy=lazy), "reg": scene_list(reg, sizes, off, lazy=lazy)}
