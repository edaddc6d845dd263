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


def gather_outputs(out: Dict[str, List[torch.Tensor]], group=None) -> Dict[str, List[torch.Tensor]]:
    """All ranks receive every scene's ``cls``/``reg`` in global scene order.  One padded all_gather per
    key (payload <= 1,464 B per actor), plus one tiny all_gather of the per-scene actor counts."""
    world = dist.get_world_size(group)
    dev = out["cls"][0].device if out["cls"] else torch.device("cpu")
    if len(out["cls"]) == 0 and dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
    sizes = torch.tensor([len(x) for x in out["cls"]], dtype=torch.int64, device=dev)
    n_scenes = torch.tensor([len(sizes)], dtype=torch.int64, device=dev)
    all_n = [torch.zeros_like(n_scenes) for _ in range(world)]
    dist.all_gather(all_n, n_scenes, group=group)
    max_s = max(int(x) for x in all_n)
    pad_sizes = torch.zeros(max(max_s, 1), dtype=torch.int64, device=dev)
    pad_sizes[: len(sizes)] = sizes
    all_sizes = [torch.zeros_like(pad_sizes) for _ in range(world)]
    dist.all_gather(all_sizes, pad_sizes, group=group)
    per_rank = [s[: int(n)].tolist() for s, n in zip(all_sizes, all_n)]
    max_a = max(1, max(sum(s) for s in per_rank))
    res = {}
    for key, tail in (("cls", (6,)), ("reg", (6, 30, 2))):
        if out[key]:
            tail = tuple(out[key][0].shape[1:])
            mine = torch.cat(out[key], 0)
        else:
            mine = torch.zeros((0,) + tail, dtype=torch.float32, device=dev)
        tail_t = torch.tensor(list(tail), dtype=torch.int64, device=dev)  # agree on the trailing shape
        tails = [torch.zeros_like(tail_t) for _ in range(world)]
        dist.all_gather(tails, tail_t, group=group)
        tail = tuple(int(x) for x in max(tails, key=lambda t: int(t.prod())))
        buf = torch.zeros((max_a,) + tail, dtype=torch.float32, device=dev)
        if mine.numel():
            buf[: mine.shape[0]] = mine
        bufs = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(bufs, buf, group=group)
        scenes = []
        for b, s in zip(bufs, per_rank):
            scenes += list(torch.split(b[: sum(s)], s))
        res[key] = scenes
    return res
