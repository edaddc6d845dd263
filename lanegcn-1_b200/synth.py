"""Seeded synthetic Argoverse-shaped scenes in the reference's *preprocessed sample* schema.

The Argoverse dataset and argoverse-api are offline, so every test and benchmark input is made here.
A scene is a dict with exactly the keys/dtypes ``ArgoDataset.__getitem__`` returns in preprocessed mode
(reference ``data.py:67-71``; graph built like ``data.py:220-361``; int16 indices like
``preprocess_data.py:230-238``; left/right as produced by ``preprocess_data.py:287-392``):

    city, orig f32[2], theta, rot f32[2,2], feats f32[A,20,3], ctrs f32[A,2], gt_preds f32[A,30,2],
    has_preds bool[A,30], idx,
    graph: ctrs f32[N,2], num_nodes, feats f32[N,2], turn f32[N,2], control f32[N], intersect f32[N],
           pre/suc: list of num_scales {u,v}, left/right {u,v}, lane_idcs, {pre,suc,left,right}_pairs

Geometry: ``roads`` straight roads, each ``parallel`` same-direction lanes 3.5 m apart, each lane a chain
of ``seq`` lanes of ``segs`` segments (~1.6 m).  Lane i's successor is the next lane of its chain; a few
extra branch/merge successor links between chains make nodes with several successors/predecessors so
the multi-scale dilation produces unsorted column order (scipy's SpGEMM order, reference
``data.py:520-534``).  Scale 0 edges follow ``data.py:272-295`` (u = destination, v = source; in-lane
links first, then the lane-boundary links, per lane).  Scales 1..5 are hops 2,4,8,16,32 by repeated
squaring of the boolean adjacency, with scipy, exactly as the reference data path does.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse

NUM_SCALES = 6

PRESETS = {
    # name: roads, parallel lanes/road, sequential lanes/chain, segments/lane, actors, extra branches
    "tiny": dict(roads=2, parallel=2, seq=5, segs=9, actors=6, branches=3),          # N = 180
    "small": dict(roads=3, parallel=2, seq=7, segs=9, actors=10, branches=4),        # N = 378
    "argo-1.5k": dict(roads=6, parallel=2, seq=14, segs=9, actors=20, branches=6),   # N = 1512
    "city-100k": dict(roads=40, parallel=4, seq=70, segs=9, actors=20, branches=40),  # N = 100800
}


def dilate_edges(u, v, num_nodes, num_scales=NUM_SCALES):
    """Scales 1..num_scales-1 of a scale-0 edge set by repeated squaring (reference data.py:520-534)."""
    mat = sparse.csr_matrix((np.ones(len(u), bool), (u, v)), shape=(num_nodes, num_nodes))
    out = []
    for _ in range(1, num_scales):
        mat = mat @ mat
        coo = mat.tocoo()
        out.append({"u": coo.row.astype(np.int64), "v": coo.col.astype(np.int64)})
    return out


def _lane_graph(rng, roads, parallel, seq, segs, branches, seg_len=1.6, lane_w=3.5):
    n_lanes = roads * parallel * seq
    lane_of = lambda r, p, s: (r * parallel + p) * seq + s  # noqa: E731
    ctrs, feats = [], []
    succ = [[] for _ in range(n_lanes)]
    pred = [[] for _ in range(n_lanes)]
    left_nb = [None] * n_lanes
    right_nb = [None] * n_lanes
    for r in range(roads):
        ang = rng.uniform(0, 2 * np.pi)
        d = np.array([np.cos(ang), np.sin(ang)])
        nrm = np.array([-d[1], d[0]])
        c0 = rng.uniform(-40, 40, 2)
        half = 0.5 * seq * segs * seg_len
        for p in range(parallel):
            base = c0 + nrm * lane_w * (p - 0.5 * (parallel - 1)) - d * half
            for s in range(seq):
                i = lane_of(r, p, s)
                t = (s * segs + np.arange(segs + 1)) * seg_len
                pts = base[None] + t[:, None] * d[None]
                pts = pts + rng.normal(0, 0.02, pts.shape)  # keep centres off exact grid values
                ctrs.append(((pts[:-1] + pts[1:]) / 2.0).astype(np.float32))
                feats.append((pts[1:] - pts[:-1]).astype(np.float32))
                if s + 1 < seq:
                    succ[i].append(lane_of(r, p, s + 1))
                    pred[lane_of(r, p, s + 1)].append(i)
                if p + 1 < parallel:
                    left_nb[i] = lane_of(r, p + 1, s)
                if p > 0:
                    right_nb[i] = lane_of(r, p - 1, s)
    for _ in range(branches):
        a, b = rng.integers(0, n_lanes, 2)
        if a != b and b not in succ[a]:
            succ[a].append(int(b))
            pred[b].append(int(a))

    node_idcs, count = [], 0
    for c in ctrs:
        node_idcs.append(np.arange(count, count + len(c)))
        count += len(c)
    N = count

    pre_u, pre_v, suc_u, suc_v = [], [], [], []
    for i in range(n_lanes):
        idcs = node_idcs[i]
        pre_u += list(idcs[1:])
        pre_v += list(idcs[:-1])
        for j in pred[i]:
            pre_u.append(idcs[0])
            pre_v.append(node_idcs[j][-1])
        suc_u += list(idcs[:-1])
        suc_v += list(idcs[1:])
        for j in succ[i]:
            suc_u.append(idcs[-1])
            suc_v.append(node_idcs[j][0])

    turn = np.zeros((N, 2), np.float32)
    control = np.zeros(N, np.float32)
    intersect = np.zeros(N, np.float32)
    for i in range(n_lanes):
        x = rng.uniform()
        if x < 0.1:
            turn[node_idcs[i], 0] = 1
        elif x < 0.2:
            turn[node_idcs[i], 1] = 1
        control[node_idcs[i]] = float(rng.uniform() < 0.1)
        intersect[node_idcs[i]] = float(rng.uniform() < 0.2)

    # left/right node edges: at most one per destination node, the matching segment of the neighbour lane
    # (what preprocess_data.py:287-392 yields for parallel lanes: nearest node of the neighbour lane).
    lu, lv, ru, rv = [], [], [], []
    for i in range(n_lanes):
        if left_nb[i] is not None:
            lu += list(node_idcs[i])
            lv += list(node_idcs[left_nb[i]])
        if right_nb[i] is not None:
            ru += list(node_idcs[i])
            rv += list(node_idcs[right_nb[i]])

    g = dict()
    g["ctrs"] = np.concatenate(ctrs, 0)
    g["num_nodes"] = N
    g["feats"] = np.concatenate(feats, 0)
    g["turn"], g["control"], g["intersect"] = turn, control, intersect
    pre0 = {"u": np.asarray(pre_u, np.int64), "v": np.asarray(pre_v, np.int64)}
    suc0 = {"u": np.asarray(suc_u, np.int64), "v": np.asarray(suc_v, np.int64)}
    g["pre"] = [pre0] + dilate_edges(pre0["u"], pre0["v"], N)
    g["suc"] = [suc0] + dilate_edges(suc0["u"], suc0["v"], N)
    g["left"] = {"u": np.asarray(lu, np.int64), "v": np.asarray(lv, np.int64)}
    g["right"] = {"u": np.asarray(ru, np.int64), "v": np.asarray(rv, np.int64)}
    g["lane_idcs"] = np.concatenate([np.full(len(x), i, np.int64) for i, x in enumerate(node_idcs)])
    pairs = lambda lst: np.asarray(lst, np.int64).reshape(-1, 2)  # noqa: E731
    g["pre_pairs"] = pairs([[i, j] for i in range(n_lanes) for j in pred[i]])
    g["suc_pairs"] = pairs([[i, j] for i in range(n_lanes) for j in succ[i]])
    g["left_pairs"] = pairs([[i, left_nb[i]] for i in range(n_lanes) if left_nb[i] is not None])
    g["right_pairs"] = pairs([[i, right_nb[i]] for i in range(n_lanes) if right_nb[i] is not None])
    return g


def _narrow(x, dtype):
    if isinstance(x, dict):
        return {k: _narrow(v, dtype) for k, v in x.items()}
    if isinstance(x, list):
        return [_narrow(v, dtype) for v in x]
    if isinstance(x, np.ndarray) and x.dtype == np.int64:
        return x.astype(dtype)
    return x


def make_scene(seed: int, preset: str = "argo-1.5k", idx: int | None = None, **overrides):
    """One scene dict (numpy arrays).  Indices are int16 when the scene has < 32768 nodes (the
    preprocessed-pickle format, preprocess_data.py:230-238) and int64 otherwise (raw data.py format)."""
    p = dict(PRESETS[preset])
    p.update(overrides)
    rng = np.random.default_rng(seed)
    g = _lane_graph(rng, p["roads"], p["parallel"], p["seq"], p["segs"], p["branches"])
    N = g["num_nodes"]
    if N < 32768:
        g = _narrow(g, np.int16)
    A = p["actors"]
    anchors = rng.integers(0, N, A)
    ctrs = g["ctrs"][anchors] + rng.normal(0, 1.0, (A, 2)).astype(np.float32)
    ctrs[0] = 0.0
    feats = np.zeros((A, 20, 3), np.float32)
    feats[:, :, :2] = rng.normal(0, 0.5, (A, 20, 2))
    feats[:, :, 2] = 1.0
    theta = float(rng.uniform(0, 2 * np.pi))
    rot = np.asarray([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]], np.float32)
    s = dict()
    s["idx"] = seed if idx is None else idx
    s["city"] = "SYN"
    s["orig"] = rng.uniform(-1000, 1000, 2).astype(np.float32)
    s["theta"] = theta
    s["rot"] = rot
    s["feats"] = feats
    s["ctrs"] = ctrs.astype(np.float32)
    s["gt_preds"] = rng.normal(0, 5.0, (A, 30, 2)).astype(np.float32)
    s["has_preds"] = np.ones((A, 30), bool)
    s["graph"] = g
    return s


def make_scenes(batch: int, preset: str = "argo-1.5k", seed0: int = 0, **overrides):
    return [make_scene(seed0 + i, preset, **overrides) for i in range(batch)]


def _to_torch(x):
    import torch

    if isinstance(x, dict):
        return {k: _to_torch(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_to_torch(v) for v in x]
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x)
    return x


def collate(scenes):
    """List of scene dicts -> dict of lists with numpy arrays turned into CPU tensors
    (same result as reference data.py:555-575 ``collate_fn``, without mutating the inputs)."""
    scenes = [_to_torch(s) for s in scenes]
    return {key: [s[key] for s in scenes] for key in scenes[0].keys()}


def seeded_state_dict(shapes: dict, seed: int = 0):
    """Deterministic, platform-independent weights for a {name: shape} table (numpy RNG, not torch's, so the
    authoring container and the GPU box produce identical tensors without shipping a 15 MB checkpoint).
    Linear/conv weights ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)); GroupNorm weight ~ 1 + 0.1 N(0,1) and every
    bias ~ 0.1 N(0,1), so the affine terms are exercised (torch's default init has weight=1, bias=0)."""
    import torch

    rng = np.random.default_rng(seed)
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        if len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            w = rng.uniform(-1.0, 1.0, shape) / np.sqrt(fan_in)
        elif name.endswith("weight"):
            w = 1.0 + 0.1 * rng.standard_normal(shape)
        else:
            w = 0.1 * rng.standard_normal(shape)
        out[name] = torch.from_numpy(w.astype(np.float32))
    return out


def make_lane_rois(scene: dict, suc_hops: int = 2, pre_hops: int = 1, near: float = 5.0):
    """Synthetic per-agent lane-RoI sub-graphs in the schema of the reference's ``generate_lane_roi``
    (data_lrcnn.py:690-844: keys a2m{u,v}, node_mask, num_nodes, feats [n,8], agent_feat [80], agent_vel, pre/suc[6]{u,v},
    left/right{u,v}; edges = the scene's node edges restricted to the RoI, re-indexed, in row-major (u, v) order).
    The lane selection is a plain breadth-first walk (``suc_hops`` successors, ``pre_hops`` predecessors of the lane
    nearest to the agent, plus their left / right neighbours) instead of the reference's velocity-dependent DFS: the
    generator only has to produce realistically sized RoIs (~50-110 nodes) for the LaneRCNN benchmark (BASELINE
    config 5) and for tests."""
    g = scene["graph"]
    lane_idcs = np.asarray(g["lane_idcs"], np.int64)
    n_lanes = int(lane_idcs[-1]) + 1
    adj = {k: [[] for _ in range(n_lanes)] for k in ("pre", "suc", "left", "right")}
    for k in adj:
        for i, j in np.asarray(g[k + "_pairs"], np.int64).reshape(-1, 2):
            adj[k][i].append(int(j))
    nodes_of = [np.nonzero(lane_idcs == i)[0] for i in range(n_lanes)]
    subs = []
    for a in range(len(scene["ctrs"])):
        d = np.sqrt(((g["ctrs"] - scene["ctrs"][a]) ** 2).sum(1))
        lanes = [int(lane_idcs[int(d.argmin())])]
        for key, hops in (("suc", suc_hops), ("pre", pre_hops)):
            frontier = [lanes[0]]
            for _ in range(hops):
                frontier = [j for i in frontier for j in adj[key][i]]
                lanes += [j for j in frontier if j not in lanes]
        for i in list(lanes):
            lanes += [j for j in adj["left"][i] + adj["right"][i] if j not in lanes]
        node_mask = np.concatenate([nodes_of[i] for i in lanes])
        if len(node_mask) < 6:
            continue
        local = np.full(len(lane_idcs), -1, np.int64)
        local[node_mask] = np.arange(len(node_mask))

        def restrict(e):
            u, v = local[np.asarray(e["u"], np.int64)], local[np.asarray(e["v"], np.int64)]
            keep = (u >= 0) & (v >= 0)
            u, v = u[keep], v[keep]
            order = np.lexsort((v, u))       # np.nonzero of the dense relation matrix: row-major
            return {"u": u[order], "v": v[order]}

        sg = {"node_mask": node_mask, "num_nodes": len(node_mask)}
        feats = np.zeros((len(node_mask), 8), np.float32)
        feats[:, :2], feats[:, 2:4], feats[:, 4:6] = g["ctrs"][node_mask], g["feats"][node_mask], g["turn"][node_mask]
        feats[:, 6], feats[:, 7] = g["control"][node_mask], g["intersect"][node_mask]
        sg["feats"] = feats
        traj = np.cumsum(scene["feats"][a, :, :2], 0) + scene["ctrs"][a] - np.sum(scene["feats"][a, :, :2], 0)
        sg["agent_feat"] = np.concatenate([traj, scene["feats"][a, :, :2]], -1).reshape(-1).astype(np.float32)
        sg["agent_vel"] = float(np.linalg.norm(scene["feats"][a, -1, :2]) * 10.0 + 0.1)
        vs = np.nonzero(d[node_mask] < near)[0].astype(np.int32)
        sg["a2m"] = {"u": np.zeros(len(vs), np.int32), "v": vs}
        for k1 in ("pre", "suc"):
            sg[k1] = [restrict(e) for e in g[k1]]
        for k1 in ("left", "right"):
            sg[k1] = restrict(g[k1])
        if len(sg["pre"][0]["u"]) == 0 and len(sg["suc"][0]["u"]) == 0:
            continue
        subs.append(sg)
    return subs
