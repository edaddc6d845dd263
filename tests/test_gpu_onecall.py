"""GPU (B200): the one-call forward (lgcn_forward: device-side sizes, capacity buckets, one CUDA graph per bucket)
against the oracle and against the module path."""
import numpy as np
import pytest
import torch

from helpers import STAGES, assert_close, golden_scenes, weights
from lanegcn_b200 import forward_engine as FE
from lanegcn_b200 import synth
from lanegcn_b200 import lanegcn as L
from oracle import graph_oracle, lanegcn_oracle as O

pytestmark = pytest.mark.gpu


def make_net(cuda):
    n = L.Net(L.config)
    n.load_state_dict(weights())
    return n.to(cuda).eval()


@pytest.fixture(scope="module")
def net(cuda, lib):
    return make_net(cuda)


def _modules_forward(net, data, monkeypatch):
    monkeypatch.setenv("LGCN_FORWARD", "modules")
    try:
        return net(data)
    finally:
        monkeypatch.delenv("LGCN_FORWARD")


@pytest.mark.parametrize("name", ["tiny_b3", "argo_b1"])
def test_index_side_bit_exact(cuda, lib, net, name):
    """CSR, batched int64 indices and the three pair lists built INSIDE lgcn_forward (device-side sizes, capacity
    layout with padded scene slots) are bit-identical to the oracle's; tiny_b3 pins the empty-scene offset quirk."""
    batch = synth.collate(golden_scenes(name))
    b = net.stage(batch)
    assert b.slot is not None, "the default configuration must take the one-call path"
    net.forward_device(b)
    assert net.check(b) is False
    torch.cuda.synchronize()
    og = O.graph_gather(O.to_long(batch["graph"]))
    edges = graph_oracle.edge_lists(og)
    n = og["feats"].shape[0]
    rowptr, col = graph_oracle.merged_csr(edges, n)
    assert np.array_equal(b.slot.buffer(0, torch.int32, n + 1).cpu().numpy(), rowptr)
    assert np.array_equal(b.slot.buffer(1, torch.int32, len(col)).cpu().numpy(), col)
    want64 = np.concatenate([np.concatenate([u, v]) for u, v in edges])
    assert np.array_equal(b.slot.buffer(2, torch.int64, len(want64)).cpu().numpy(), want64)
    nctr, actr = [g["ctrs"].numpy() for g in batch["graph"]], [c.numpy() for c in batch["ctrs"]]
    st = b.slot.status_host.tolist()
    for i, (a, c, th) in enumerate([(nctr, actr, 7.0), (actr, nctr, 6.0), (actr, actr, 100.0)]):
        hi, wi = graph_oracle.pair_list(a, c, th)
        assert st[1 + i] == len(hi)
        assert np.array_equal(b.slot.buffer(3 + i, torch.int32, len(hi)).cpu().numpy(), hi)
        assert np.array_equal(b.slot.buffer(6 + i, torch.int32, len(wi)).cpu().numpy(), wi)


@pytest.mark.parametrize("name", ["tiny_b3", "argo_b1"])
def test_onecall_stages_vs_oracle(cuda, lib, net, name):
    batch = synth.collate(golden_scenes(name))
    taps, want = {}, {}
    with torch.no_grad():
        ref = O.net_forward(weights(), batch, want)
    out = net.forward_taps(batch, taps)
    for s in STAGES:
        assert_close(taps[s], want[s], s, rtol=1e-4, atol=2e-5 if s == "actor_net" else 1e-5)
    assert_close(torch.cat(out["cls"]), torch.cat(ref["cls"]), "cls")
    assert_close(torch.cat(out["reg"]), torch.cat(ref["reg"]), "reg")


def test_graph_replay_equals_eager_and_modules(cuda, lib, net, monkeypatch):
    """Graph replay == eager run of the same sequence (bitwise), and == the module path up to the stock-PyTorch
    ActorNet / PredNet (which see a padded row count on the one-call path)."""
    batch = synth.collate(synth.make_scenes(4, "small", seed0=50))
    a = net(batch)                     # captures
    a2 = net(batch)                    # replays
    net.use_cuda_graphs = False
    try:
        e = net(batch)
    finally:
        net.use_cuda_graphs = True
    m = _modules_forward(net, batch, monkeypatch)
    for k in ("cls", "reg"):
        assert torch.equal(torch.cat(a[k]), torch.cat(a2[k])), "replay is not deterministic"
        assert torch.equal(torch.cat(a[k]), torch.cat(e[k])), "graph replay differs from the eager sequence"
        assert_close(torch.cat(a[k]), torch.cat(m[k]), "one-call vs modules " + k, rtol=1e-5, atol=1e-5)


def test_bucket_is_reused_for_batches_of_different_sizes(cuda, lib, monkeypatch):
    """Batches with different node / actor / edge / pair counts that fit the same capacities replay ONE graph; every
    result equals the module path's for that batch."""
    net = make_net(cuda)
    base = synth.make_scenes(6, "small", seed0=70)
    variants = [base, base[:5] + synth.make_scenes(1, "tiny", seed0=3), synth.make_scenes(6, "small", seed0=90)]
    outs = [net(synth.collate(v)) for v in variants]
    keys = {k for k in net._buckets}
    assert len(keys) <= 2, f"expected the batches to share capacity buckets, got {len(keys)}"
    graphs = sum(s.graph is not None for bk in net._buckets.values() for s in bk.slots)
    assert graphs <= 2 * len(keys)
    for v, o in zip(variants, outs):
        m = _modules_forward(net, synth.collate(v), monkeypatch)
        for k in ("cls", "reg"):
            assert [len(x) for x in o[k]] == [len(x) for x in m[k]]
            assert_close(torch.cat(o[k]), torch.cat(m[k]), k, rtol=1e-5, atol=1e-5)


def test_pair_capacity_overflow_is_detected_and_rerun(cuda, lib, monkeypatch):
    net = make_net(cuda)
    batch = synth.collate(synth.make_scenes(3, "small", seed0=11))
    want = _modules_forward(net, batch, monkeypatch)
    real = FE.pair_caps
    calls = []

    def tiny_first(node_sizes, actor_sizes, learned, cap_actors):
        calls.append(learned)
        return (64, 64, 64) if learned is None else real(node_sizes, actor_sizes, learned, cap_actors)
    monkeypatch.setattr(FE, "pair_caps", tiny_first)
    b = net.stage(batch)
    net.forward_device(b)
    assert net.check(b) is True, "64-pair capacities must overflow on this batch"
    got = net(batch)                   # learned counts -> larger bucket -> valid result
    assert calls[0] is None and calls[-1] is not None
    for k in ("cls", "reg"):
        assert_close(torch.cat(got[k]), torch.cat(want[k]), k, rtol=1e-5, atol=1e-5)
    got2 = list(L.prefetch_forward(net, iter([batch, batch]), to_host=True))
    for g in got2:
        assert_close(torch.cat(g["reg"]), torch.cat(want["reg"]), "prefetch reg", rtol=1e-5, atol=1e-5)


def test_no_pairs_raises_like_reference(cuda, lib):
    net = make_net(cuda)
    scenes = synth.make_scenes(2, "tiny", seed0=5)
    for s in scenes:
        s["ctrs"] = s["ctrs"] + np.float32(5000.0)     # every actor far from every lane node: A2M / M2A lists empty
    with pytest.raises(RuntimeError, match="no agent/context pair"):
        net(synth.collate(scenes))


def test_bad_edge_index_raises(cuda, lib):
    net = make_net(cuda)
    scenes = synth.make_scenes(2, "tiny", seed0=5)
    scenes[1]["graph"]["suc"][2]["v"][3] = 30000
    with pytest.raises(RuntimeError, match="out of range"):
        net(synth.collate(scenes))
    monkey = pytest.MonkeyPatch()
    monkey.setenv("LGCN_FORWARD", "modules")
    try:
        with pytest.raises(RuntimeError, match="out of range"):
            net(synth.collate(scenes))
    finally:
        monkey.undo()


def test_weight_updates_reach_the_captured_graph(cuda, lib):
    net = make_net(cuda)
    batch = synth.collate(golden_scenes("tiny_b3"))
    a = torch.cat(net(batch)["reg"])
    with torch.no_grad():
        net.m2m.fuse["ctr"][0].weight.mul_(1.25)
        net.map_net.seg[2].linear.weight.add_(0.01)
    b = torch.cat(net(batch)["reg"])
    assert not torch.equal(a, b)
    net.load_state_dict(weights())
    assert torch.equal(torch.cat(net(batch)["reg"]), a)
    with torch.no_grad():
        net.a2a.att[1].agt.weight.data.mul_(0.5)      # .data writes do not bump the version counter ...
    net.invalidate_packs()                            # ... so they need an explicit invalidation
    assert not torch.equal(torch.cat(net(batch)["reg"]), a)


def test_second_device_if_present(lib):
    """Per-device library state (function attributes, split-weight ring, SM count): a forward on cuda:1 after cuda:0."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    batch = synth.collate(golden_scenes("tiny_b3"))
    outs = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        with torch.cuda.device(dev):
            outs.append(torch.cat(make_net(dev)(batch)["reg"]).cpu())
            x = torch.randn(300, 128, device=dev)
            w = torch.randn(128, 128, device=dev) / 11
            y = torch.empty_like(x)
            from lanegcn_b200 import _C
            _C.check(lib.lgcn_linear128(x.data_ptr(), None, None, None, None, None, 1, None, 0, w.data_ptr(), 1, None, None,
                                        None, 0, y.data_ptr(), 128, 300, torch.cuda.current_stream().cuda_stream))
            assert_close(y, x.double().cpu() @ w.double().cpu().T, "linear on cuda:%d" % d)
    assert_close(outs[1], outs[0], "cuda:1 vs cuda:0", rtol=1e-5, atol=1e-5)


# --------------------------------------------------------------------------- ActorNet / PredNet kernels (f2, f3)
@pytest.mark.parametrize("path", ["tc", "fp32"])
@pytest.mark.parametrize("n", [1, 3, 4, 5, 64, 257])
def test_actor_net_kernel_vs_oracle_and_torch(cuda, lib, net, n, path, monkeypatch):
    monkeypatch.setenv("LGCN_ACTOR", path)   # tc: output Res1d on the tensor core (default); fp32: the single fp32 kernel
    g = torch.Generator().manual_seed(n)
    feats = torch.randn(n, 20, 3, generator=g) * 0.5
    feats[:, :, 2] = (torch.rand(n, 20, generator=g) > 0.2).float()
    feats[: n // 2, :7] = 0.0                                 # padded history prefix, as in the dataset
    with torch.no_grad():
        want = O.actor_net(weights(), feats.transpose(1, 2).contiguous())
    got = net.actor_net(feats.transpose(1, 2).contiguous().to(cuda))       # module API: channels-first [A,3,20]
    assert_close(got, want, f"actor_net kernel n={n}", rtol=1e-4, atol=2e-5)
    got2 = net.actor_net.forward_ntc(feats.to(cuda))
    assert torch.equal(got, got2)
    ref = net.actor_net.forward_torch(feats.transpose(1, 2).contiguous().to(cuda))   # cuDNN fp32 spelling
    assert_close(got, ref, "actor_net kernel vs torch", rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("n", [1, 15, 16, 17, 100])
def test_pred_net_kernel_vs_oracle(cuda, lib, net, n):
    g = torch.Generator().manual_seed(10 + n)
    actors = torch.randn(n, 128, generator=g)
    ctrs = torch.randn(n, 2, generator=g) * 30
    sizes = [n // 2, n - n // 2] if n > 1 else [1]
    idcs = list(torch.arange(n).split(sizes))
    with torch.no_grad():
        want = O.pred_net(weights(), actors, idcs, list(ctrs.split(sizes)))
    got = net.pred_net(actors.to(cuda), [i.to(cuda) for i in idcs], [c.to(cuda) for c in ctrs.split(sizes)])
    assert [len(x) for x in got["cls"]] == sizes
    assert_close(torch.cat(got["cls"]), torch.cat(want["cls"]), "cls")
    assert_close(torch.cat(got["reg"]), torch.cat(want["reg"]), "reg")
    # world transform folded in (lanegcn.py:145-150)
    rot = torch.randn(len(sizes), 2, 2, generator=g)
    orig = torch.randn(len(sizes), 2, generator=g) * 100
    off = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int32)
    cls, reg = net.pred_net.core(actors.to(cuda), ctrs.to(cuda), off.to(cuda), rot.to(cuda), orig.to(cuda))
    want_w = torch.cat([torch.matmul(r, rot[i]) + orig[i].view(1, 1, 1, -1) for i, r in enumerate(want["reg"])])
    assert_close(reg, want_w, "reg (world)")
    assert torch.equal(cls, torch.cat(got["cls"]))


def test_packed_scene_staging_equals_dict_staging(cuda, lib, net):
    """Net.stage from packed scenes (lgcn_stage_scenes: one C call + 4 H2D copies) == Net.stage from the dict of lists."""
    scenes = synth.make_scenes(3, "small", seed0=31) + synth.make_scenes(2, "tiny", seed0=32)
    want = net(synth.collate(scenes))
    data = L.pack_batch(synth.collate(scenes))
    b = net.stage(data)
    assert b.slot is not None and b.h2d_bytes > 0
    got = net(data)
    for k in ("cls", "reg"):
        assert [len(x) for x in got[k]] == [len(x) for x in want[k]]
        assert torch.equal(torch.cat(got[k]), torch.cat(want[k]))
    host = list(L.prefetch_forward(net, iter([data, data, data]), to_host=True))
    assert len(host) == 3 and all(torch.equal(torch.cat(h["reg"]), torch.cat(want["reg"]).cpu()) for h in host)
