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


def test_stage_scenes_matches_reference_batching(lib):
    """lgcn_stage_scenes (packed scenes -> the four staging buffers in capacity layout) against a numpy restatement of
    collate + graph_gather's concatenation order (data.py:555-561, lanegcn.py:171-209), with padded scene slots."""
    import ctypes

    from lanegcn_b200 import _C, forward_engine as FE, lanegcn as L, synth

    scenes = synth.make_scenes(3, "tiny", seed0=1) + synth.make_scenes(2, "small", seed0=5)
    data = L.pack_batch(synth.collate(scenes))
    packed = data["_packed"]
    B, S, n_kv = len(packed), 6, 28
    N, A, tot = sum(p.n_nodes for p in packed), sum(p.n_actors for p in packed), sum(p.n_index for p in packed)
    Nc, Ac, Ic, Bc = FE.round_cap(N, 128), FE.round_cap(A), FE.round_cap(tot, 1024), 7
    off = np.cumsum([0, 2 * Nc, 2 * Nc, 2 * Nc, Nc, Nc, 60 * Ac, 2 * Ac, 4 * Bc, 2 * Bc])
    fl, idx = np.full(off[-1], np.nan, np.float32), np.full(Ic, -7, np.int16)
    t64, t32 = np.zeros(2 * n_kv * Bc + 1, np.int64), np.zeros(2 * (Bc + 1) + 4, np.int32)
    ptrs = (ctypes.c_void_p * B)(*[p.ptr for p in packed])
    _C.check(lib.lgcn_stage_scenes(ptrs, B, Nc, Ac, Ic, Bc, S, 2, fl.ctypes.data, idx.ctypes.data, t64.ctypes.data,
                                   t32.ctypes.data, 4))
    g = data["graph"]
    for r, (key, k) in enumerate([("ctrs", 2), ("feats", 2), ("turn", 2), ("control", 1), ("intersect", 1)]):
        assert np.array_equal(fl[off[r]:off[r] + k * N], np.concatenate([x[key].numpy().ravel() for x in g])), key
    for r, (key, k) in enumerate([("feats", 60 * A), ("ctrs", 2 * A), ("rot", 4 * B), ("orig", 2 * B)]):
        assert np.array_equal(fl[off[5 + r]:off[5 + r] + k], np.concatenate([x.numpy().ravel() for x in data[key]])), key
    locs = []
    for k1, sc in L._edge_names(S):
        src = [x[k1] for x in g] if sc is None else [x[k1][sc] for x in g]
        for k2 in ("u", "v"):
            locs += [d[k2].numpy() for d in src]
    assert np.array_equal(idx[:tot], np.concatenate(locs))
    lens = np.zeros((n_kv, Bc), np.int64)
    lens[:, :B] = np.array([len(x) for x in locs]).reshape(n_kv, B)
    assert np.array_equal(t64[:n_kv * Bc + 1], np.concatenate(([0], np.cumsum(lens.ravel()))))
    noff, aoff = np.full(Bc + 1, N), np.full(Bc + 1, A)
    noff[:B + 1] = np.concatenate(([0], np.cumsum([p.n_nodes for p in packed])))
    aoff[:B + 1] = np.concatenate(([0], np.cumsum([p.n_actors for p in packed])))
    assert np.array_equal(t64[n_kv * Bc + 1:], np.tile(noff[:-1], n_kv))
    assert np.array_equal(t32, np.concatenate((noff, aoff, [N, A, 0, 0])).astype(np.int32))
    # capacity violations are errors, not overruns
    assert lib.lgcn_stage_scenes(ptrs, B, N - 1, Ac, Ic, Bc, S, 2, fl.ctypes.data, idx.ctypes.data, t64.ctypes.data,
                                 t32.ctypes.data, 1) != 0
    assert lib.lgcn_stage_scenes(ptrs, B, Nc, Ac, Ic, Bc, S, 4, fl.ctypes.data, idx.ctypes.data, t64.ctypes.data,
                                 t32.ctypes.data, 1) != 0
