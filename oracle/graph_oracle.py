"""ORACLE — integer/index side of the path, numpy + small pure-Python loops.  Test infrastructure only.

* dilate_smmp      restates scipy's CSR x CSR product ordering (third-party: scipy.sparse `csr_matmat`,
                   SMMP algorithm; reference call site data.py:520-534, unpinned scipy, here 1.18.1)
* merged_csr       the destination-sorted, key-then-edge-ordered CSR the LaneConv gather consumes; this is
                   the summation order CPU `index_add_` realises at lanegcn.py:333-354
* pair_list        lanegcn.py:672-689 in numpy, incl. the empty-scene offset quirk
"""
from __future__ import annotations

import numpy as np

KEY_ORDER = [f"{d}{s}" for s in range(6) for d in ("pre", "suc")] + ["left", "right"]  # lanegcn.py:288-291


def rows_from_edges(u, v, n):
    """csr_matrix((ones,(u,v))) : per row, ascending de-duplicated columns (data.py:521-522)."""
    rows = [[] for _ in range(n)]
    for a, b in sorted(set(zip(np.asarray(u).tolist(), np.asarray(v).tolist()))):
        rows[a].append(b)
    return rows


def dilate_smmp(u, v, n, num_scales=6):
    """Scales 1..num_scales-1 (hops 2,4,..) with scipy's per-row column order: columns come out in
    REVERSE first-discovery order (linked-list head insertion in csr_matmat)."""
    rows = rows_from_edges(u, v, n)
    out = []
    for _ in range(1, num_scales):
        new = []
        for i in range(n):
            seen, order = set(), []
            for j in rows[i]:
                for k in rows[j]:
                    if k not in seen:
                        seen.add(k)
                        order.append(k)
            new.append(order[::-1])
        rows = new
        uu = np.asarray([i for i in range(n) for _ in rows[i]], np.int64)
        vv = np.asarray([k for i in range(n) for k in rows[i]], np.int64)
        out.append({"u": uu, "v": vv})
    return out


def edge_lists(graph) -> list:
    """[(u,v)] in accumulation order pre0,suc0,...,pre5,suc5,left,right from a batched graph dict."""
    out = []
    for s in range(len(graph["pre"])):
        for d in ("pre", "suc"):
            out.append((np.asarray(graph[d][s]["u"]), np.asarray(graph[d][s]["v"])))
    for d in ("left", "right"):
        out.append((np.asarray(graph[d]["u"]), np.asarray(graph[d]["v"])))
    return out


def merged_csr(edges, n, n_keys=None):
    """rowptr int32[n+1], col int32[E] with col = v*(K+1) + (k+1): the 128-float block of the wide
    projection Y[n, (K+1)*128] to add into row u; per row ordered by key k, then by edge-list position
    (stable) — the order CPU index_add_ accumulates in."""
    K = len(edges) if n_keys is None else n_keys
    u_all = np.concatenate([np.asarray(u, np.int64) for u, _ in edges]) if edges else np.zeros(0, np.int64)
    blk = np.concatenate(
        [np.asarray(v, np.int64) * (K + 1) + (k + 1) for k, (_, v) in enumerate(edges)]
    ) if edges else np.zeros(0, np.int64)
    order = np.argsort(u_all, kind="stable")
    rowptr = np.zeros(n + 1, np.int64)
    np.add.at(rowptr, u_all + 1, 1)
    return np.cumsum(rowptr).astype(np.int32), blk[order].astype(np.int32)


def laneconv_plan(edges, n, tile=128):
    """What lgcn_laneconv_plan_build must encode for the aggregate-first LaneConv kernel: for every destination row m
    and key k the ordered list of source rows (edge-list order) — exactly the terms ``temp.index_add_(0, u, W_k x[v])``
    adds into row m (lanegcn.py:333-354).  Returns ``lists[k][m]`` (python lists) and the padded row count."""
    rows = (n + tile - 1) // tile * tile
    lists = []
    for u, v in edges:
        per = [[] for _ in range(n)]
        for uu, vv in zip(np.asarray(u, np.int64).tolist(), np.asarray(v, np.int64).tolist()):
            per[uu].append(vv)
        lists.append(per)
    return lists, rows


def pair_list(agt_ctrs, ctx_ctrs, dist_th, fix_empty_scene_offsets=False):
    """hi,wi int64 for lists of per-scene f32[*,2] centres (lanegcn.py:672-689).  fp32 arithmetic with
    separate roundings: sub, square, add, sqrt, <=."""
    hi, wi, hc, wc = [], [], 0, 0
    th = np.float32(dist_th)
    for a, c in zip(agt_ctrs, ctx_ctrs):
        a = np.asarray(a, np.float32).reshape(-1, 2)
        c = np.asarray(c, np.float32).reshape(-1, 2)
        d = a[:, None, :] - c[None, :, :]
        d = d * d
        dist = np.sqrt(d[:, :, 0] + d[:, :, 1])
        r, q = np.nonzero(dist <= th)
        if len(r) == 0 and not fix_empty_scene_offsets:
            continue
        hi.append(r.astype(np.int64) + hc)
        wi.append(q.astype(np.int64) + wc)
        hc += len(a)
        wc += len(c)
    if not hi:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.concatenate(hi), np.concatenate(wi)


# --------------------------------------------------------------------------- data.py:272-295
def scale0_edges(lane_idcs, pre_pairs, suc_pairs):
    """Scale-0 pre / suc node edges from the lane topology, in the reference's order (data.py:272-295): per lane the
    in-lane chain, then one boundary link per (lane, neighbour) pair; the pairs are the lane-level ``pre_pairs`` /
    ``suc_pairs`` of the same function (data.py:302-317), which list a lane's neighbours in the same order."""
    lane_idcs = np.asarray(lane_idcs, np.int64)
    n_lanes = int(lane_idcs[-1]) + 1 if len(lane_idcs) else 0
    node_idcs = [np.nonzero(lane_idcs == i)[0] for i in range(n_lanes)]
    pre = {"u": [], "v": []}
    suc = {"u": [], "v": []}
    for i in range(n_lanes):
        idcs = node_idcs[i]
        pre["u"] += list(idcs[1:])
        pre["v"] += list(idcs[:-1])
        for a, j in np.asarray(pre_pairs, np.int64).reshape(-1, 2):
            if a == i:
                pre["u"].append(idcs[0])
                pre["v"].append(node_idcs[j][-1])
        suc["u"] += list(idcs[:-1])
        suc["v"] += list(idcs[1:])
        for a, j in np.asarray(suc_pairs, np.int64).reshape(-1, 2):
            if a == i:
                suc["u"].append(idcs[-1])
                suc["v"].append(node_idcs[j][0])
    as64 = lambda d: {k: np.asarray(v, np.int64) for k, v in d.items()}  # noqa: E731
    return as64(pre), as64(suc)


# --------------------------------------------------------------------------- preprocess_data.py:287-392
def side_edges(ctrs, feats, lane_idcs, side_pairs, pre_pairs, suc_pairs, cross_dist):
    """Left (or right) node edges of ``preprocess()`` with cross_angle=None: dense fp32 distance matrix, candidates =
    nodes of the lanes reachable as side neighbour or a predecessor / successor of it, row minimum (first index on
    ties), distance and heading filters.  numpy fp32 restatement of the torch ops, same order."""
    ctrs, feats = np.asarray(ctrs, np.float32), np.asarray(feats, np.float32)
    lane_idcs = np.asarray(lane_idcs, np.int64)
    side_pairs = np.asarray(side_pairs, np.int64).reshape(-1, 2)
    if len(side_pairs) == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    n_lanes = int(lane_idcs[-1]) + 1
    d = ctrs[:, None, :] - ctrs[None, :, :]
    dist = np.sqrt((d * d).sum(2, dtype=np.float32))
    mk = lambda pairs: _dense(pairs, n_lanes)  # noqa: E731
    mat = mk(side_pairs)
    mat = (mat @ mk(pre_pairs) + mat @ mk(suc_pairs) + mat) > 0.5
    masked = np.where(mat[lane_idcs[:, None], lane_idcs[None, :]], dist, np.float32(1e6))
    vi = masked.argmin(1)
    ok = masked[np.arange(len(ctrs)), vi] < np.float32(cross_dist)
    ui = np.nonzero(ok)[0]
    vi = vi[ok]
    t1 = np.arctan2(feats[ui, 1], feats[ui, 0])
    t2 = np.arctan2(feats[vi, 1], feats[vi, 0])
    dt = np.abs(t1 - t2)
    m = dt > np.float32(np.pi)
    dt[m] = np.abs(dt[m] - np.float32(2 * np.pi))
    m = dt < np.float32(0.25 * np.pi)
    return ui[m].astype(np.int64), vi[m].astype(np.int64)


def _dense(pairs, n):
    m = np.zeros((n, n), np.float32)
    pairs = np.asarray(pairs, np.int64).reshape(-1, 2)
    if len(pairs):
        m[pairs[:, 0], pairs[:, 1]] = 1
    return m
