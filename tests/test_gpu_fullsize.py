"""GPU (B200): the drop-in forward against the CPU oracle AT THE BENCHMARKED SIZES — BASELINE configs[1]
(batch 32 argo-1.5k) and configs[2] (batch 128 argo-1.5k, what bench.py times).  Per-stage taps and the final
cls/reg are held to the north-star tolerance (1e-4 relative / 1e-5 absolute); the worst error/tolerance per stage
is written to gpurun_out/parity_fullsize.json (copied to profiles/ per round)."""
import json
import os

import pytest
import torch

from helpers import ATOL, RTOL, STAGES, weights
from lanegcn_b200 import synth
from lanegcn_b200 import lanegcn as L
from oracle import lanegcn_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def net(cuda, lib):
    n = L.Net(L.config)
    n.load_state_dict(weights())
    return n.to(cuda).eval()


@pytest.fixture(scope="module")
def oracle_runs():
    """Oracle forward (with per-stage taps) per batch size, computed once per module."""
    cache = {}

    def get(b):
        if b not in cache:
            scenes = synth.make_scenes(b, "argo-1.5k", seed0=0)
            taps = {}
            with torch.no_grad():
                out = O.net_forward(weights(), synth.collate(scenes), taps)
            cache[b] = (scenes, taps, out)
        return cache[b]
    return get


def _ratio(got, want):
    got, want = got.detach().cpu().double(), torch.as_tensor(want).double()
    assert got.shape == want.shape
    err = (got - want).abs()
    return float((err / (ATOL + RTOL * want.abs())).max()), float(err.max())


def _record(tag, table):
    d = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(d):
        return
    p = os.path.join(d, "parity_fullsize.json")
    cur = json.load(open(p)) if os.path.exists(p) else {}
    cur[tag] = table
    json.dump(cur, open(p, "w"), indent=1, sort_keys=True)


def _engine(lib, which):
    """which: 'tcgen05' (default path, aggregate-first LaneConv), 'tcgen05-split', 'simt'"""
    eng = 0 if which == "simt" else 1
    prev = (lib.lgcn_set_gemm_engine(eng), L.LANECONV_FUSED)
    if lib.lgcn_get_gemm_engine() != eng:
        lib.lgcn_set_gemm_engine(prev[0])
        pytest.skip("engine not built")
    L.LANECONV_FUSED = which == "tcgen05"
    return prev


def _restore(lib, prev):
    lib.lgcn_set_gemm_engine(prev[0])
    L.LANECONV_FUSED = prev[1]


def _check_forward(net, lib, oracle_runs, b, which):
    scenes, taps, want = oracle_runs(b)
    prev = _engine(lib, which)
    try:
        got_taps = {}
        out = net.forward_taps(synth.collate(scenes), got_taps)
    finally:
        _restore(lib, prev)
    table = {}
    for s in STAGES:
        table[s] = _ratio(got_taps[s], taps[s])
    table["cls"] = _ratio(torch.cat(list(out["cls"])), torch.cat(want["cls"]))
    table["reg"] = _ratio(torch.cat(list(out["reg"])), torch.cat(want["reg"]))
    _record(f"b{b}_{which}", {k: {"max_err_over_tol": round(r, 4), "max_abs_err": e} for k, (r, e) in table.items()})
    print(f"\nB={b} {which}: " + "  ".join(f"{k} {r:.2f}" for k, (r, _) in table.items()))
    bad = {k: r for k, (r, _) in table.items() if not r <= 1.0}
    assert not bad, f"B={b} {which}: error/tolerance above 1 for {bad}"
    assert [len(x) for x in out["cls"]] == [len(s["ctrs"]) for s in scenes]


@pytest.mark.parametrize("which", ["tcgen05", "tcgen05-split", "simt"])
def test_net_forward_b32_vs_oracle_all_engines(cuda, lib, net, oracle_runs, which):
    """BASELINE configs[1]: batch 32, every stage + cls/reg, on all three engines."""
    _check_forward(net, lib, oracle_runs, 32, which)


def test_net_forward_b128_vs_oracle(cuda, lib, net, oracle_runs):
    """BASELINE configs[2] — the size bench.py times (193,536 nodes, 2,560 actors, 1,512 row tiles)."""
    _check_forward(net, lib, oracle_runs, 128, "tcgen05")


def test_modules_b32_fed_with_oracle_inputs(cuda, lib, net, oracle_runs):
    """Each module on the ORACLE's input of that stage at batch 32 (errors do not accumulate), default engine."""
    scenes, taps, _ = oracle_runs(32)
    batch = synth.collate(scenes)
    graph = L.graph_gather(batch["graph"])
    actor_ctrs = L._as_scene_list([c.to(cuda) for c in batch["ctrs"]])
    actor_idcs = L.scene_list(torch.arange(len(actor_ctrs.cat), device=cuda), [len(c) for c in actor_ctrs])
    table = {}
    nodes, _, _ = net.map_net(graph)
    table["map_net"] = _ratio(nodes, taps["map_net"])
    table["a2m"] = _ratio(net.a2m(taps["map_net"].to(cuda), graph, taps["actor_net"].to(cuda), actor_idcs, actor_ctrs), taps["a2m"])
    table["m2m"] = _ratio(net.m2m(taps["a2m"].to(cuda), graph), taps["m2m"])
    table["m2a"] = _ratio(net.m2a(taps["actor_net"].to(cuda), actor_idcs, actor_ctrs, taps["m2m"].to(cuda), graph["idcs"],
                                  graph["ctrs"]), taps["m2a"])
    table["a2a"] = _ratio(net.a2a(taps["m2a"].to(cuda), actor_idcs, actor_ctrs), taps["a2a"])
    _record("b32_modules_on_oracle_inputs", {k: {"max_err_over_tol": round(r, 4), "max_abs_err": e} for k, (r, e) in table.items()})
    bad = {k: r for k, (r, _) in table.items() if not r <= 1.0}
    assert not bad, bad


def test_actor_gather_matches_oracle(cuda, lib):
    """lanegcn.py:155-168: [sum A, 3, 20] transposed concat (bit-exact: pure data movement) + per-scene index lists."""
    batch = synth.collate(synth.make_scenes(5, "tiny", seed0=40) + synth.make_scenes(2, "small", seed0=3))
    want, want_idcs = O.actor_gather(batch["feats"])
    for feats in (batch["feats"], [f.to(cuda) for f in batch["feats"]]):   # host lists and device lists
        got, idcs = L.actor_gather(feats)
        assert got.is_cuda and got.dtype == torch.float32 and got.is_contiguous()
        assert torch.equal(got.cpu(), want)
        assert len(idcs) == len(want_idcs)
        for a, b in zip(idcs, want_idcs):
            assert a.dtype == torch.int64 and torch.equal(a.cpu(), b)
