"""CPU: host-side logic — parameter contract, synthetic generator, sharding, the world_size-2 gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import shapes
from lanegcn_b200 import lanegcn as L
from lanegcn_b200 import shard, synth


def test_state_dict_names_and_shapes_match_reference():
    want = shapes()
    got = {k: list(v.shape) for k, v in L.Net(L.config).state_dict().items()}
    assert got == want and len(got) == 405
    assert sum(int(np.prod(s)) for s in got.values()) == 3701161


def test_state_dict_round_trip():
    net = L.Net(L.config)
    sd = synth.seeded_state_dict(shapes(), 3)
    net.load_state_dict(sd)
    back = net.state_dict()
    assert all(torch.equal(back[k], sd[k]) for k in sd)


def test_module_api_surface():
    for name in ("config", "Net", "ActorNet", "MapNet", "A2M", "M2M", "M2A", "A2A", "Att", "PredNet",
                 "actor_gather", "graph_gather", "get_model"):
        assert hasattr(L, name), name
    for key in ("n_map", "n_actor", "num_scales", "actor2map_dist", "map2actor_dist", "actor2actor_dist",
                "num_mods", "num_preds"):
        assert key in L.config
    import inspect

    assert list(inspect.signature(L.Att.forward).parameters)[:8] == [
        "self", "agts", "agt_idcs", "agt_ctrs", "ctx", "ctx_idcs", "ctx_ctrs", "dist_th"]
    assert list(inspect.signature(L.A2M.forward).parameters)[:6] == [
        "self", "feat", "graph", "actors", "actor_idcs", "actor_ctrs"]
    assert list(inspect.signature(L.M2A.forward).parameters)[:7] == [
        "self", "actors", "actor_idcs", "actor_ctrs", "nodes", "node_idcs", "node_ctrs"]


def test_no_cpu_fallback():
    net = L.Net(L.config)
    with pytest.raises(RuntimeError, match="no CPU path"):
        net(synth.collate(synth.make_scenes(1, "tiny")))
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        L.Att(128, 128)(torch.zeros(2, 128), [torch.arange(2)], [torch.zeros(2, 2)], torch.zeros(2, 128),
                        [torch.arange(2)], [torch.zeros(2, 2)], 1.0)


def test_synth_schema_and_determinism():
    a, b = synth.make_scene(5, "small"), synth.make_scene(5, "small")
    g = a["graph"]
    assert g["num_nodes"] == 378 and g["ctrs"].dtype == np.float32 and g["pre"][0]["u"].dtype == np.int16
    assert len(g["pre"]) == len(g["suc"]) == 6 and a["feats"].shape == (10, 20, 3)
    assert all(np.array_equal(g["pre"][s]["v"], b["graph"]["pre"][s]["v"]) for s in range(6))
    assert np.array_equal(a["ctrs"], b["ctrs"])
    for d in ("left", "right"):  # at most one left / right edge per destination node
        assert len(np.unique(g[d]["u"])) == len(g[d]["u"])
    big = synth.make_scene(0, "city-100k", roads=2, seq=20)  # still > int16 range? no: stays small -> int16
    assert big["graph"]["pre"][0]["u"].dtype == np.int16


def test_partition_balanced_and_contiguous():
    costs = [1512] * 128
    for world in (1, 2, 4, 8):
        parts = shard.partition(costs, world)
        assert [len(p) for p in parts] == [128 // world] * world
        assert parts[0].start == 0 and parts[-1].stop == 128
        assert all(parts[i].stop == parts[i + 1].start for i in range(world - 1))
    parts = shard.partition([100, 1, 1, 1, 100, 1], 2)
    assert [list(p) for p in parts] == [[0, 1, 2, 3], [4, 5]] or sum(len(p) for p in parts) == 6
    parts = shard.partition([5, 5], 4)
    assert sum(len(p) for p in parts) == 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scenes = synth.make_scenes(5, "tiny", seed0=40)
        scenes[3] = synth.make_scene(99, "small")  # uneven node and actor counts
        data = synth.collate(scenes)
        mine = shard.shard_batch(data, rank, world)
        # stand-in forward: a per-scene function of the inputs (the kernels need a GPU; the sharding does not)
        out = {"cls": [c.sum(1, keepdim=True).repeat(1, 6) for c in mine["ctrs"]],
               "reg": [c.view(-1, 1, 1, 2).repeat(1, 6, 30, 1) for c in mine["ctrs"]]}
        full = shard.gather_outputs(out)
        ok = len(full["cls"]) == 5
        for i, c in enumerate(data["ctrs"]):
            ok &= torch.equal(full["cls"][i], c.sum(1, keepdim=True).repeat(1, 6))
            ok &= torch.equal(full["reg"][i], c.view(-1, 1, 1, 2).repeat(1, 6, 30, 1))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shard_and_gather_gloo(world):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)
