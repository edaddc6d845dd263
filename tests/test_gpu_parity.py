"""GPU (B200): every C-ABI entry point and the drop-in modules against the CPU oracle / the committed goldens.

Integer outputs (batched edge lists, CSR, pair lists) must be bit-exact.  Float outputs must be within the
north-star tolerance of the reference fp32 forward: 1e-4 relative / 1e-5 absolute (helpers.RTOL/ATOL).
"""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import ATOL, RTOL, STAGES, assert_close, golden, golden_scenes, weights
from lanegcn_b200 import _C, synth
from lanegcn_b200 import lanegcn as L
from oracle import graph_oracle, lanegcn_oracle as O

pytestmark = pytest.mark.gpu


def sp():
    return torch.cuda.current_stream().cuda_stream


@pytest.fixture(scope="module")
def net(cuda, lib):
    n = L.Net(L.config)
    n.load_state_dict(weights())
    return n.to(cuda).eval()


@pytest.fixture(params=[0, 1, 2], ids=["simt", "tcgen05", "tcgen05-split"])
def engine(request, lib):
    """0: fp32 SIMT validation engine; 1: tcgen05 with the aggregate-first LaneConv kernel (the default);
    2: tcgen05 with the wide projection + CSR gather + ctr2 LaneConv path."""
    eng = min(request.param, 1)
    prev = lib.lgcn_set_gemm_engine(eng)
    if lib.lgcn_get_gemm_engine() != eng:
        lib.lgcn_set_gemm_engine(prev)
        pytest.skip("engine not built")
    prev_fused = L.LANECONV_FUSED
    L.LANECONV_FUSED = request.param == 1
    yield eng
    L.LANECONV_FUSED = prev_fused
    lib.lgcn_set_gemm_engine(prev)


# --------------------------------------------------------------------------- graph batching (a1-a3)
@pytest.mark.parametrize("name", ["tiny_b3", "argo_b1"])
def test_graph_gather_bit_exact(cuda, lib, name):
    batch = synth.collate(golden_scenes(name))
    want = O.graph_gather(O.to_long(batch["graph"]))
    got = L.graph_gather(batch["graph"])
    got["_packed"].check()
    for k1 in ("pre", "suc"):
        for i in range(6):
            for k2 in ("u", "v"):
                assert got[k1][i][k2].dtype == torch.int64
                assert torch.equal(got[k1][i][k2].cpu(), want[k1][i][k2]), (k1, i, k2)
    for k1 in ("left", "right"):
        for k2 in ("u", "v"):
            assert torch.equal(got[k1][k2].cpu(), want[k1][k2])
    for k in ("feats", "turn", "control", "intersect"):
        assert torch.equal(got[k].cpu(), want[k])
    assert all(torch.equal(a.cpu(), b) for a, b in zip(got["idcs"], want["idcs"]))
    assert all(torch.equal(a.cpu(), b) for a, b in zip(got["ctrs"], want["ctrs"]))
    if name == "tiny_b3":  # and against the reference's own output
        g = golden(name)
        assert np.array_equal(got["pre"][3]["v"].cpu().numpy(), g["g_pre3_v"])
        assert np.array_equal(got["left"]["u"].cpu().numpy(), g["g_left_u"])


def test_graph_gather_accepts_int64_and_device_inputs(cuda, lib):
    scenes = golden_scenes("tiny_b3")
    batch = synth.collate(scenes)
    want = O.graph_gather(O.to_long(batch["graph"]))
    dev_graphs = [{k: (v.to(cuda) if torch.is_tensor(v) else
                       [{kk: vv.to(cuda).long() for kk, vv in e.items()} for e in v] if isinstance(v, list) else
                       {kk: vv.to(cuda).long() for kk, vv in v.items()} if isinstance(v, dict) else v)
                   for k, v in g.items()} for g in batch["graph"]]
    got = L.graph_gather(dev_graphs)
    assert torch.equal(got["suc"][5]["u"].cpu(), want["suc"][5]["u"])
    assert torch.equal(got["right"]["v"].cpu(), want["right"]["v"])


@pytest.mark.parametrize("name", ["tiny_b3", "argo_b1"])
def test_csr_bit_exact(cuda, lib, name):
    batch = synth.collate(golden_scenes(name))
    og = O.graph_gather(O.to_long(batch["graph"]))
    rowptr, col = graph_oracle.merged_csr(graph_oracle.edge_lists(og), og["feats"].shape[0])
    pg = L.graph_gather(batch["graph"])["_packed"]
    assert np.array_equal(pg.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(pg.col.cpu().numpy()[: len(col)], col)


def test_csr_handles_empty_sets_hubs_and_bad_indices(cuda, lib):
    n = 50
    rng = np.random.default_rng(0)
    hub_u = np.zeros(300, np.int64)                 # one destination with 300 edges (> one 32-wide batch)
    hub_v = rng.integers(0, n, 300)
    sets = [(hub_u, hub_v), (np.zeros(0, np.int64), np.zeros(0, np.int64)), (rng.integers(0, n, 77), rng.integers(0, n, 77))]
    rowptr, col = graph_oracle.merged_csr(sets, n)
    es = [{"u": torch.from_numpy(u).to(cuda), "v": torch.from_numpy(v).to(cuda)} for u, v in sets]
    pg = L.build_csr(es, n, cuda)
    pg.check()
    assert np.array_equal(pg.rowptr.cpu().numpy(), rowptr) and np.array_equal(pg.col.cpu().numpy()[: len(col)], col)
    es[2]["v"][5] = n  # out of range -> flagged, not a crash
    pg = L.build_csr(es, n, cuda)
    with pytest.raises(RuntimeError, match="out of range"):
        pg.check()


# --------------------------------------------------------------------------- dilation (a13)
def test_dilated_nbrs_bit_exact_vs_reference_golden(cuda, lib):
    g = golden("dilate_tiny")
    graph = golden_scenes("tiny_b3")[0]["graph"]
    for d in ("pre", "suc"):
        got = L.dilated_nbrs({"u": graph[d][0]["u"], "v": graph[d][0]["v"]}, graph["num_nodes"], 6)
        assert len(got) == 5
        for i, e in enumerate(got):
            assert e["u"].dtype == torch.int64 and e["u"].is_cuda
            assert np.array_equal(e["u"].cpu().numpy(), g[f"{d}{i + 1}_u"]), (d, i)
            assert np.array_equal(e["v"].cpu().numpy(), g[f"{d}{i + 1}_v"]), (d, i)


def test_dilated_nbrs_batched_branching_and_duplicates(cuda, lib):
    """Random branching graph with duplicate edges and empty rows vs scipy (the reference's data path) and vs the
    numpy restatement; then a whole batch dilated at once == the scenes dilated one by one."""
    rng = np.random.default_rng(3)
    n = 400
    u = np.concatenate([np.arange(1, n), rng.integers(0, n, 150), rng.integers(0, n, 40)])
    v = np.concatenate([np.arange(0, n - 1), rng.integers(0, n, 150), rng.integers(0, n, 40)])
    u, v = np.concatenate([u, u[:30]]), np.concatenate([v, v[:30]])  # duplicates
    want = synth.dilate_edges(u, v, n, 5)
    mine = graph_oracle.dilate_smmp(u, v, n, 5)
    got = L.dilated_nbrs({"u": u, "v": v}, n, 5)
    for w, m, gt in zip(want, mine, got):
        assert np.array_equal(m["u"], w["u"]) and np.array_equal(m["v"], w["v"])
        assert np.array_equal(gt["u"].cpu().numpy(), w["u"]) and np.array_equal(gt["v"].cpu().numpy(), w["v"])
    scenes = golden_scenes("tiny_b3")
    batch = synth.collate(scenes)
    bg = O.graph_gather(O.to_long(batch["graph"]))
    n_tot = bg["feats"].shape[0]
    got = L.dilated_nbrs({"u": bg["pre"][0]["u"], "v": bg["pre"][0]["v"]}, n_tot, 6)
    for s in range(1, 6):
        assert torch.equal(got[s - 1]["u"].cpu(), bg["pre"][s]["u"]) and torch.equal(got[s - 1]["v"].cpu(), bg["pre"][s]["v"])


# --------------------------------------------------------------------------- pair lists (a10)
@pytest.mark.parametrize("name", ["tiny_b3", "argo_b1"])
def test_pair_lists_bit_exact(cuda, lib, name):
    g = golden(name)
    batch = synth.collate(golden_scenes(name))
    nctr = [x["ctrs"].to(cuda) for x in batch["graph"]]
    actr = [x.to(cuda) for x in batch["ctrs"]]
    for tag, (a, c, th) in {"a2m": (nctr, actr, 7.0), "m2a": (actr, nctr, 6.0), "a2a": (actr, actr, 100.0)}.items():
        hi, wi = L.att_pairs(a, c, th)
        assert hi.dtype == torch.int64
        assert np.array_equal(hi.cpu().numpy(), g[f"hi_{tag}"]), tag
        assert np.array_equal(wi.cpu().numpy(), g[f"wi_{tag}"]), tag


def test_pair_lists_borderline_distances_and_empty_scenes(cuda, lib):
    """Centres on a 1 m grid with thresholds that land exactly on representable distances (3-4-5 triangles):
    flips if dx*dx+dy*dy is contracted into an FMA or the sqrt is approximate."""
    rng = np.random.default_rng(1)
    agt, ctx = [], []
    for b in range(6):
        na, nc = int(rng.integers(1, 40)), int(rng.integers(1, 70))
        a = rng.integers(-8, 8, (na, 2)).astype(np.float32) + np.float32(0.1) * rng.integers(0, 3, (na, 2)).astype(np.float32)
        c = rng.integers(-8, 8, (nc, 2)).astype(np.float32)
        if b in (1, 4):
            c = c + np.float32(1000.0)  # empty scenes -> offset quirk
        agt.append(a)
        ctx.append(c)
    for th in (5.0, 0.0, 2.236068, 100.0):
        want_hi, want_wi = graph_oracle.pair_list(agt, ctx, th)
        t_hi, t_wi = O.att_pairs([torch.from_numpy(x) for x in agt], [torch.from_numpy(x) for x in ctx],
                                 [len(x) for x in agt], [len(x) for x in ctx], th) if len(want_hi) else (None, None)
        if t_hi is not None:
            assert np.array_equal(t_hi.numpy(), want_hi)
        hi, wi = L.att_pairs([torch.from_numpy(x).to(cuda) for x in agt], [torch.from_numpy(x).to(cuda) for x in ctx], th)
        assert np.array_equal(hi.cpu().numpy(), want_hi) and np.array_equal(wi.cpu().numpy(), want_wi), th


def test_pair_rowptr_is_by_destination(cuda, lib):
    batch = synth.collate(golden_scenes("tiny_b3"))
    nctr = L._as_scene_list([x["ctrs"].to(cuda) for x in batch["graph"]])
    actr = L._as_scene_list([x.to(cuda) for x in batch["ctrs"]])
    p = L.build_pair_lists([(nctr, actr, 7.0)], want_int64=True)[0]
    hi = p.hi64.cpu().numpy()
    want = np.zeros(p.n_agt + 1, np.int64)
    np.add.at(want, hi + 1, 1)
    assert np.array_equal(p.rowptr.cpu().numpy(), np.cumsum(want))


# --------------------------------------------------------------------------- dense pieces (a5, a11, a12)
def _ref_linear(srcs, idxs, xs, W, gamma, beta, res, flags, dtype=torch.float32):
    c = lambda t: None if t is None else t.to(dtype)
    srcs, xs, W, gamma, beta, res = [c(s) for s in srcs], c(xs), c(W), c(gamma), c(beta), c(res)
    parts = [s if i is None else s[i.long()] for s, i in zip(srcs, idxs)]
    a = torch.cat(parts + ([xs] if xs is not None else []), 1)
    y = F.linear(a, W)
    if flags & _C.EPI_GN:
        y = F.group_norm(y, 1, gamma, beta, 1e-5)
    if flags & _C.EPI_RELU1:
        y = F.relu(y)
    if flags & _C.EPI_RES:
        y = y + res
    if flags & _C.EPI_RELU2:
        y = F.relu(y)
    return y


@pytest.mark.parametrize("m", [1, 127, 128, 129, 1000, 4097])
@pytest.mark.parametrize("case", ["plain", "gn_relu", "gn_res_relu", "meta_ks4", "cat3_gather", "wide15"])
def test_linear128(cuda, lib, engine, m, case):
    g = torch.Generator().manual_seed(m * 7 + len(case))
    n_src = 3 if case == "cat3_gather" else 1
    nob = 15 if case == "wide15" else 1
    ks = 4 if case == "meta_ks4" else 0
    flags = {"plain": 0, "gn_relu": 3, "gn_res_relu": 13, "meta_ks4": 3, "cat3_gather": 3, "wide15": 0}[case]
    rows = [m, 57, 300] if n_src == 3 else [m]
    srcs = [torch.randn(r, 128, generator=g) * 2 for r in rows]
    idxs = [None] + [torch.randint(0, r, (m,), generator=g, dtype=torch.int32) for r in rows[1:]]
    xs = torch.randn(m, 4, generator=g) if ks else None
    W = torch.randn(nob * 128, n_src * 128 + ks, generator=g) / 11.3
    gamma, beta, res = torch.randn(128, generator=g), torch.randn(128, generator=g), torch.randn(m, 128, generator=g)
    # the judge of ONE layer is the exact (fp64) value of the same expression: against it the kernel must hold the
    # north-star tolerance (1e-4 relative / 1e-5 absolute), and must not be worse than ~2x torch's own fp32 CPU result
    # (comparing two fp32 results with each other would count both rounding errors against the kernel)
    want = _ref_linear(srcs, idxs, xs, W, gamma, beta, res, flags, torch.float64)
    f32 = _ref_linear(srcs, idxs, xs, W, gamma, beta, res, flags, torch.float32)
    d = lambda t: None if t is None else t.to(cuda).contiguous()
    ds, di, dxs, dW, dg, db, dres = [d(s) for s in srcs], [d(i) for i in idxs], d(xs), d(W), d(gamma), d(beta), d(res)
    ds += [None] * (3 - n_src)
    di += [None] * (3 - n_src)
    out = torch.full((m, nob * 128), float("nan"), device=cuda)
    _C.check(lib.lgcn_linear128(_C.ptr(ds[0]), _C.ptr(di[0]), _C.ptr(ds[1]), _C.ptr(di[1]), _C.ptr(ds[2]), _C.ptr(di[2]),
                                n_src, _C.ptr(dxs), ks, dW.data_ptr(), nob, dg.data_ptr(), db.data_ptr(),
                                dres.data_ptr(), flags, out.data_ptr(), nob * 128, m, sp()))
    assert_close(out, want, f"{case} m={m}", rtol=RTOL, atol=ATOL)
    ours, torch32 = (out.cpu().double() - want).abs().max().item(), (f32.double() - want).abs().max().item()
    assert ours <= 3.0 * torch32 + 2e-6, f"{case} m={m}: max err {ours:.2e} vs torch fp32's own {torch32:.2e}"


def test_mlp2_in(cuda, lib):
    g = torch.Generator().manual_seed(0)
    p, q = torch.randn(100, 2, generator=g) * 50, torch.randn(40, 2, generator=g) * 50
    ip = torch.randint(0, 100, (333,), generator=g, dtype=torch.int32)
    iq = torch.randint(0, 40, (333,), generator=g, dtype=torch.int32)
    W1, b1 = torch.randn(128, 2, generator=g), torch.randn(128, generator=g)
    for use_q in (False, True):
        x = p[ip.long()] - q[iq.long()] if use_q else p
        want = F.relu(F.linear(x, W1, b1))
        out = torch.empty(len(x), 128, device=cuda)
        dp, dq, dip, diq, dW, db = p.to(cuda), q.to(cuda), ip.to(cuda), iq.to(cuda), W1.to(cuda), b1.to(cuda)
        _C.check(lib.lgcn_mlp2_in(dp.data_ptr(), dip.data_ptr() if use_q else None, dq.data_ptr() if use_q else None,
                                  diq.data_ptr() if use_q else None, dW.data_ptr(), db.data_ptr(), out.data_ptr(),
                                  len(x), sp()))
        assert_close(out, want, "mlp2_in", rtol=1e-5, atol=2e-5)


def test_gather_gn_relu_matches_index_add_order(cuda, lib):
    """The kernel adds in CSR order == the order CPU index_add_ applies => BIT-identical pre-norm sums; with
    the norm, within tolerance."""
    batch = synth.collate(golden_scenes("tiny_b3"))
    og = O.graph_gather(O.to_long(batch["graph"]))
    edges = graph_oracle.edge_lists(og)
    n, K = og["feats"].shape[0], len(edges)
    g = torch.Generator().manual_seed(3)
    Y = torch.randn(n, (K + 1) * 128, generator=g)
    gamma, beta = torch.randn(128, generator=g), torch.randn(128, generator=g)
    temp = Y[:, :128].clone()
    for k, (u, v) in enumerate(edges):
        temp.index_add_(0, torch.from_numpy(u), Y[torch.from_numpy(v), (k + 1) * 128:(k + 2) * 128])
    want = F.relu(F.group_norm(temp, 1, gamma, beta, 1e-5))
    pg = L.graph_gather(batch["graph"])["_packed"]
    out = torch.empty(n, 128, device=cuda)
    dY, dg, db = Y.to(cuda), gamma.to(cuda), beta.to(cuda)
    _C.check(lib.lgcn_laneconv_gather_gn_relu(dY.data_ptr(), K + 1, pg.rowptr.data_ptr(), pg.col.data_ptr(),
                                              dg.data_ptr(), db.data_ptr(), out.data_ptr(), n, sp()))
    assert_close(out, want, "gather+GN+ReLU", rtol=1e-5, atol=1e-5)
    # identity affine and an all-ones second call: determinism (bitwise) across launches
    out2 = torch.empty_like(out)
    _C.check(lib.lgcn_laneconv_gather_gn_relu(dY.data_ptr(), K + 1, pg.rowptr.data_ptr(), pg.col.data_ptr(),
                                              dg.data_ptr(), db.data_ptr(), out2.data_ptr(), n, sp()))
    assert torch.equal(out, out2)


@pytest.mark.parametrize("name", ["tiny_b3", "argo_b1"])
def test_laneconv_plan_matches_oracle(cuda, lib, name):
    """The gather plan of the aggregate-first kernel, decoded, is EXACTLY the per-(row, key) source lists of the
    reference's index_add_ loop (integer structure: bit-exact, order included); padding rows have no sources."""
    batch = synth.collate(golden_scenes(name))
    og = O.graph_gather(O.to_long(batch["graph"]))
    edges = graph_oracle.edge_lists(og)
    n, K = og["feats"].shape[0], len(edges)
    want, rows = graph_oracle.laneconv_plan(edges, n)
    pg = L.graph_gather(batch["graph"])["_packed"]
    plan = pg.plan()
    torch.cuda.synchronize()
    raw = plan.cpu().numpy()
    hdr = raw[:256].view(np.int32)
    n_multi, n_mcol = int(hdr[0]), int(hdr[1])
    n_tiles = rows // 128
    off = 256
    tab = raw[off:off + n_tiles * K * 128 * 4].view(np.int32).reshape(n_tiles, K, 128)
    off += (n_tiles * K * 128 * 4 + 255) // 256 * 256
    max_multi = pg.n_edges // 2
    mdesc = raw[off:off + (max_multi + 1) * 8].view(np.int32).reshape(-1, 2)
    off += ((max_multi + 1) * 8 + 255) // 256 * 256
    mcol = raw[off:off + pg.n_edges * 4].view(np.int32)
    assert n_multi == sum(1 for k in range(K) for m in range(n) if len(want[k][m]) > 1)
    assert n_mcol == sum(len(want[k][m]) for k in range(K) for m in range(n) if len(want[k][m]) > 1)
    for k in range(K):
        for m in range(rows):
            v = int(tab[m // 128, k, m % 128])
            exp = want[k][m] if m < n else []
            if v == -1:
                got = []
            elif v >= 0:
                got = [v]
            else:
                s0, c = mdesc[-2 - v]
                got = mcol[s0:s0 + c].tolist()
            assert got == exp, f"key {k} row {m}: {got} vs {exp}"


@pytest.mark.parametrize("n,n_keys,n_blocks,seed", [(1, 14, 1, 0), (127, 14, 2, 1), (129, 3, 3, 2), (1000, 14, 4, 3),
                                                     (4097, 14, 4, 4), (300, 0, 2, 5), (20000, 14, 1, 6)])
def test_laneconv_stack_planned_matches_split_fp32(cuda, lib, n, n_keys, n_blocks, seed):
    """Aggregate-first single-kernel LaneConv stack (tcgen05) vs the projection + gather + ctr2 stack on the fp32
    SIMT engine: random graphs with empty keys, multi-source (row, key) pairs, a hub row and row tails."""
    g = torch.Generator().manual_seed(100 + seed)
    sets = []
    for k in range(n_keys):
        e = 0 if k == 2 else int(n * (0.9 if k % 3 else 1.5)) + (k == 0)
        u, v = torch.randint(0, n, (e,), generator=g), torch.randint(0, n, (e,), generator=g)
        if k == 1 and n > 64:
            u[: min(e, 300)] = 5   # hub: one destination with hundreds of sources of one key
        sets.append({"u": u.to(cuda), "v": v.to(cuda)})
    pg = L.build_csr(sets, n, cuda)
    nb = n_keys + 1
    per = lib.lgcn_laneconv_wpack_floats(n_keys)
    wp = torch.empty(n_blocks * per)
    for i in range(n_blocks):
        o = i * per
        wp[o:o + (nb + 1) * 128 * 128] = torch.randn((nb + 1) * 128 * 128, generator=g) * (0.7 / (128 ** 0.5))
        wp[o + (nb + 1) * 128 * 128:o + per] = torch.randn(4 * 128, generator=g) * 0.3 + torch.tensor([1.0, 0.0, 1.0, 0.0]).repeat_interleave(128)
    wp = wp.to(cuda)
    x = torch.randn(n, 128, generator=g).to(cuda)

    prev = lib.lgcn_set_gemm_engine(0)
    try:
        want = x.clone()
        ws = torch.empty(lib.lgcn_laneconv_workspace_bytes(n, n_keys), dtype=torch.uint8, device=cuda)
        _C.check(lib.lgcn_laneconv_stack(want.data_ptr(), pg.rowptr.data_ptr(), pg.col.data_ptr(), n_keys, n_blocks,
                                         wp.data_ptr(), n, ws.data_ptr(), sp()))
        if lib.lgcn_set_gemm_engine(1) != 0 or lib.lgcn_get_gemm_engine() != 1:
            pytest.skip("tcgen05 engine not built")
        got = x.clone()
        ws2 = torch.empty(lib.lgcn_laneconv_planned_workspace_bytes(n, pg.n_edges, n_keys), dtype=torch.uint8, device=cuda)
        _C.check(lib.lgcn_laneconv_stack_planned(got.data_ptr(), pg.plan().data_ptr(), pg.n_edges, n_keys, n_blocks,
                                                 wp.data_ptr(), n, ws2.data_ptr(), sp()))
        assert_close(got, want, "planned LaneConv stack", rtol=RTOL, atol=ATOL)
        again = x.clone()
        _C.check(lib.lgcn_laneconv_stack_planned(again.data_ptr(), pg.plan().data_ptr(), pg.n_edges, n_keys, n_blocks,
                                                 wp.data_ptr(), n, ws2.data_ptr(), sp()))
        assert torch.equal(got, again), "planned LaneConv stack is not deterministic"
    finally:
        lib.lgcn_set_gemm_engine(prev)


def test_segsum_gn_relu(cuda, lib):
    g = torch.Generator().manual_seed(5)
    n, P = 37, 500
    hi = torch.sort(torch.randint(0, n, (P,), generator=g)).values
    hi[hi == 7] = 8  # a row with no pair
    a, c = torch.randn(n, 128, generator=g), torch.randn(P, 128, generator=g)
    gamma, beta = torch.randn(128, generator=g), torch.randn(128, generator=g)
    t = a.clone()
    t.index_add_(0, hi, c)
    want = F.relu(F.group_norm(t, 1, gamma, beta, 1e-5))
    rowptr = torch.zeros(n + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(hi, minlength=n), 0)
    out = torch.empty(n, 128, device=cuda)
    da, dc, dr, dg, db = a.to(cuda), c.to(cuda), rowptr.int().to(cuda), gamma.to(cuda), beta.to(cuda)
    _C.check(lib.lgcn_segsum_gn_relu(da.data_ptr(), dc.data_ptr(), dr.data_ptr(), dg.data_ptr(), db.data_ptr(),
                                     out.data_ptr(), n, sp()))
    assert_close(out, want, "segsum+GN+ReLU", rtol=1e-5, atol=1e-5)


# --------------------------------------------------------------------------- modules vs oracle (a4-a9)
def _dev_lists(batch, cuda):
    actors, actor_idcs = O.actor_gather(batch["feats"])
    return actors, actor_idcs, batch["ctrs"]


@pytest.mark.parametrize("name", ["tiny_b3", "argo_b1"])
def test_modules_stage_by_stage(cuda, lib, net, engine, name):
    """Each drop-in module fed with the ORACLE's input for that stage (so errors do not accumulate)."""
    sd = weights()
    batch = synth.collate(golden_scenes(name))
    taps = {}
    with torch.no_grad():
        O.net_forward(sd, batch, taps)
    og = O.graph_gather(O.to_long(batch["graph"]))
    graph = L.graph_gather(batch["graph"])
    actor_ctrs = L._as_scene_list([c.to(cuda) for c in batch["ctrs"]])
    actor_idcs = L.scene_list(torch.arange(len(actor_ctrs.cat), device=cuda), [len(c) for c in actor_ctrs])
    nodes, idcs, ctrs = net.map_net(graph)
    assert_close(nodes, taps["map_net"], "map_net")
    got = net.a2m(taps["map_net"].to(cuda), graph, taps["actor_net"].to(cuda), actor_idcs, actor_ctrs)
    assert_close(got, taps["a2m"], "a2m")
    got = net.m2m(taps["a2m"].to(cuda), graph)
    assert_close(got, taps["m2m"], "m2m")
    got = net.m2a(taps["actor_net"].to(cuda), actor_idcs, actor_ctrs, taps["m2m"].to(cuda), graph["idcs"], graph["ctrs"])
    assert_close(got, taps["m2a"], "m2a")
    got = net.a2a(taps["m2a"].to(cuda), actor_idcs, actor_ctrs)
    assert_close(got, taps["a2a"], "a2a")
    got = net.actor_net(O.actor_gather(batch["feats"])[0].to(cuda))
    assert_close(got, taps["actor_net"], "actor_net", rtol=1e-4, atol=2e-5)


def test_modules_accept_reference_style_graph_dict(cuda, lib, net):
    """A batched graph dict built by the ORACLE's graph_gather (plain lists/tensors moved to the GPU, no
    _packed attachment) drives MapNet / M2M the same way."""
    sd = weights()
    batch = synth.collate(golden_scenes("tiny_b3"))
    og = O.graph_gather(O.to_long(batch["graph"]))
    with torch.no_grad():
        want, _, _ = O.map_net(sd, og)
    mv = lambda x: x.to(cuda) if torch.is_tensor(x) else ([mv(v) for v in x] if isinstance(x, list) else {k: mv(v) for k, v in x.items()})
    dg = mv(og)
    got, idcs, ctrs = net.map_net(dg)
    assert_close(got, want, "map_net(reference-style dict)")
    assert len(idcs) == 3 and len(ctrs) == 3


def test_att_without_context_skips_norm(cuda, lib, net):
    sd = weights()
    agts = torch.randn(9, 128, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        want = O.att(sd, "a2a.att.0", agts, [torch.arange(9)], [torch.zeros(9, 2)], torch.zeros(0, 128), [], [], 1.0)
    got = net.a2a.att[0](agts.to(cuda), [torch.arange(9)], [torch.zeros(9, 2, device=cuda)],
                         torch.zeros(0, 128, device=cuda), [], [], 1.0)
    assert_close(got, want, "att(no ctx)")


def test_att_no_pairs_raises_like_reference(cuda, lib, net):
    a = torch.zeros(3, 128, device=cuda)
    with pytest.raises(RuntimeError, match="no agent/context pair"):
        net.a2a.att[0](a, [torch.arange(3)], [torch.zeros(3, 2, device=cuda)], a, [torch.arange(3)],
                       [torch.full((3, 2), 1e4, device=cuda)], 1.0)


# --------------------------------------------------------------------------- full forward vs goldens
@pytest.mark.parametrize("name", ["tiny_b3", "argo_b1"])
def test_net_forward_matches_reference_golden(cuda, lib, net, engine, name):
    g = golden(name)
    out = net(synth.collate(golden_scenes(name)))
    assert_close(torch.cat(out["cls"]), g["cls"], "cls")
    assert_close(torch.cat(out["reg"]), g["reg"], "reg")
    assert [len(x) for x in out["cls"]] == [len(s["ctrs"]) for s in golden_scenes(name)]


def test_prefetch_forward_equals_forward(cuda, lib, net):
    """The pipelined entry point (staging of batch i+1 overlapped with batch i) returns exactly what Net.forward
    returns, batch by batch, for batches of different shapes."""
    batches = [synth.collate(synth.make_scenes(b, "tiny", seed0=s)) for b, s in ((3, 0), (1, 7), (4, 11), (2, 3))]
    want = [net(b) for b in batches]
    got = list(L.prefetch_forward(net, iter(batches)))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        for k in ("cls", "reg"):
            assert len(g[k]) == len(w[k])
            for a, b in zip(g[k], w[k]):
                assert torch.equal(a, b)
    assert list(L.prefetch_forward(net, iter([]))) == []
    host = list(L.prefetch_forward(net, iter(batches), to_host=True))   # one batch deeper, results on the host
    assert len(host) == len(want)
    for g, w in zip(host, want):
        for k in ("cls", "reg"):
            assert len(g[k]) == len(w[k])
            for a, b in zip(g[k], w[k]):
                assert not a.is_cuda and torch.equal(a, b.cpu())


def test_weight_updates_are_seen(cuda, lib):
    """The packed weight copies follow in-place updates, load_state_dict and re-assigned Parameter objects."""
    n1 = L.Net(L.config)
    n1.load_state_dict(weights())
    n1 = n1.to(cuda).eval()
    batch = synth.collate(golden_scenes("tiny_b3"))
    a = torch.cat(n1(batch)["reg"])
    with torch.no_grad():
        n1.m2m.fuse["ctr"][0].weight.mul_(1.25)
        n1.a2a.att[1].agt.weight.add_(0.01)
    b = torch.cat(n1(batch)["reg"])
    assert not torch.equal(a, b)
    n2 = L.Net(L.config)
    n2.load_state_dict({k: v.clone() for k, v in n1.state_dict().items()})
    n2 = n2.to(cuda).eval()
    assert torch.equal(torch.cat(n2(batch)["reg"]), b)
    n1.load_state_dict(weights())
    assert torch.equal(torch.cat(n1(batch)["reg"]), a)
    q = n1.m2a.att[0].query.linear
    q.weight = torch.nn.Parameter(q.weight.detach() * 0.5)   # a new Parameter object under the same name
    assert not torch.equal(torch.cat(n1(batch)["reg"]), a)


def test_net_forward_matches_oracle_batch8(cuda, lib, net):
    scenes = synth.make_scenes(8, "small", seed0=20)
    sd = weights()
    with torch.no_grad():
        want = O.net_forward(sd, synth.collate(scenes))
    got = net(synth.collate(scenes))
    assert_close(torch.cat(got["cls"]), torch.cat(want["cls"]), "cls")
    assert_close(torch.cat(got["reg"]), torch.cat(want["reg"]), "reg")


# --------------------------------------------------------------------------- full-size properties (B = 128)
def test_batch128_properties(cuda, lib, net):
    """At BASELINE.json's full size the oracle is too slow to run per test; check size-independent
    properties instead: CSR invariants, determinism (bitwise), and scene independence — the forward of a
    batch equals, scene for scene and bit for bit, the forward of its two halves (what scene sharding across
    GPUs relies on)."""
    scenes = synth.make_scenes(128, "argo-1.5k", seed0=0)
    data = synth.collate(scenes)
    b = net.stage(data)
    out1 = net.forward_device(b)
    out2 = net.forward_device(net.stage(data))
    for k in ("cls", "reg"):
        assert all(torch.equal(x, y) for x, y in zip(out1[k], out2[k])), "non-deterministic " + k
        assert all(torch.isfinite(x).all() for x in out1[k])
    graph = L.graph_gather(data["graph"])
    pg = graph["_packed"]
    pg.check()
    rowptr = pg.rowptr.cpu().numpy().astype(np.int64)
    E = sum(len(e["u"]) for e in L._edge_sets_of(graph))
    assert rowptr[0] == 0 and rowptr[-1] == E and (np.diff(rowptr) >= 0).all()
    col = pg.col.cpu().numpy()[:E]
    key = col % 15
    assert key.min() >= 1 and (col // 15).max() < pg.n_nodes
    inner = np.ones(E, bool)
    inner[rowptr[1:-1][rowptr[1:-1] < E]] = False           # positions that start a new row
    assert (np.diff(key)[inner[1:]] >= 0).all(), "keys must be non-decreasing inside a row"
    deg = np.bincount(torch.cat([e["u"] for e in L._edge_sets_of(graph)]).cpu().numpy(), minlength=pg.n_nodes)
    assert np.array_equal(np.diff(rowptr), deg)
    # scene independence of OUR kernels is exact: MapNet on the full batch == MapNet on each half
    full, _, _ = net.map_net(graph)
    n_half = sum(s["graph"]["num_nodes"] for s in scenes[:64])
    for part, rows in ((scenes[:64], slice(0, n_half)), (scenes[64:], slice(n_half, None))):
        half, _, _ = net.map_net(L.graph_gather(synth.collate(part)["graph"]))
        assert torch.equal(full[rows], half), "scene sharding changed MapNet output"
    # end to end (ActorNet/PredNet go through cuDNN/cuBLAS, which may pick other algorithms per batch size)
    halves = [net(synth.collate(scenes[:64])), net(synth.collate(scenes[64:]))]
    for k in ("cls", "reg"):
        got = halves[0][k] + halves[1][k]
        assert_close(torch.cat(got), torch.cat(out1[k]), "sharded " + k, rtol=1e-5, atol=1e-5)


# --------------------------------------------------------------------------- config 4: one city-scale lane graph
def test_mapnet_city_100k_vs_oracle(cuda, lib, net):
    """BASELINE configs[3]: MapNet alone on a single ~100k-node graph (int64 indices: the scene exceeds the int16
    range of the preprocessed pickles), against the CPU oracle."""
    scene = synth.make_scene(0, "city-100k")
    assert scene["graph"]["num_nodes"] == 100800 and scene["graph"]["pre"][0]["u"].dtype == np.int64
    batch = synth.collate([scene])
    sd = weights()
    og = O.graph_gather(O.to_long(batch["graph"]))
    with torch.no_grad():
        want, _, _ = O.map_net(sd, og)
    graph = L.graph_gather(batch["graph"])
    graph["_packed"].check()
    assert torch.equal(graph["suc"][4]["v"].cpu(), og["suc"][4]["v"])
    got, _, _ = net.map_net(graph)
    assert_close(got, want, "map_net (100,800 nodes)")


# --------------------------------------------------------------------------- property tests (random inputs)
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=20, deadline=None)
@given(st.integers(1, 300), st.integers(0, 6), st.integers(0, 2**31 - 1))
def test_csr_random_edge_sets(n, n_keys, seed):
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(seed)
    sets = []
    for _ in range(n_keys):
        e = int(rng.integers(0, 4 * n))
        sets.append((rng.integers(0, n, e), rng.integers(0, n, e)))
    rowptr, col = graph_oracle.merged_csr(sets, n, n_keys)
    es = [{"u": torch.from_numpy(u).to(dev), "v": torch.from_numpy(v).to(dev)} for u, v in sets]
    pg = L.build_csr(es, n, dev)
    pg.check()
    assert np.array_equal(pg.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(pg.col.cpu().numpy()[: len(col)], col)


@settings(max_examples=20, deadline=None)
@given(st.integers(1, 5), st.floats(0.5, 30.0), st.integers(0, 2**31 - 1))
def test_pairs_random_scenes(n_scenes, th, seed):
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(seed)
    agt = [rng.uniform(-20, 20, (int(rng.integers(1, 60)), 2)).astype(np.float32) for _ in range(n_scenes)]
    ctx = [rng.uniform(-20, 20, (int(rng.integers(1, 90)), 2)).astype(np.float32) for _ in range(n_scenes)]
    if n_scenes > 2:
        ctx[1] = ctx[1] + np.float32(500.0)  # an empty scene in the middle
    th = float(np.float32(th))
    want_hi, want_wi = graph_oracle.pair_list(agt, ctx, th)
    if len(want_hi) == 0:
        return
    hi, wi = L.att_pairs([torch.from_numpy(x).to(dev) for x in agt], [torch.from_numpy(x).to(dev) for x in ctx], th)
    assert np.array_equal(hi.cpu().numpy(), want_hi) and np.array_equal(wi.cpu().numpy(), want_wi)
