"""ORACLE — test infrastructure only.  Not imported by the product package.

CPU restatement of the LaneGCN forward graph path of leepaul009/LaneGCN-1, written as plain functions
over a flat ``state_dict`` (reference parameter names).  The reference's arithmetic
lives in a third-party dependency that is not under /root/reference: **PyTorch** (README.MD:28,36 pins
"PyTorch>=1.3.1", example 1.5.1; here torch 2.11.0 CPU) — ``mm``, advanced indexing, ``index_add_``,
``group_norm``, ``nonzero``.  Each function below issues the same torch CPU ops in the same order as the
reference call site it cites, so its outputs are bit-identical to the reference run on CPU
(pinned by tests/test_oracle_vs_reference.py when /root/reference is present, and by the committed
goldens in tests/golden/ otherwise).

Parity status: the reference ships no tests, golden vectors or fixtures for this path, so
the pin is "outputs of the reference itself run here": tests/golden/make_golden.py imports the
unmodified reference through oracle/ref_loader.py and writes the fixtures.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
EPS = 1e-5  # nn.GroupNorm default


# --------------------------------------------------------------------------- layers.py:65-87
def gn(x: Tensor, sd, name: str) -> Tensor:
    """nn.GroupNorm(gcd(1, C)=1 group, C) (layers.py:72)."""
    return F.group_norm(x, 1, sd[name + ".weight"], sd[name + ".bias"], EPS)


def linear_gn(x: Tensor, sd, name: str, act: bool = True) -> Tensor:
    """layers.Linear: bias-free Linear -> GroupNorm(1) -> optional ReLU (layers.py:65-87)."""
    out = gn(F.linear(x, sd[name + ".linear.weight"]), sd, name + ".norm")
    return F.relu(out) if act else out


# --------------------------------------------------------------------------- utils.py:88-96
def to_long(x):
    """int16 index tensors -> int64 (utils.py:88-96); everything else untouched."""
    if isinstance(x, dict):
        return {k: to_long(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [to_long(v) for v in x]
    if torch.is_tensor(x) and x.dtype == torch.int16:
        return x.long()
    return x


# --------------------------------------------------------------------------- lanegcn.py:155-168
def actor_gather(actors: List[Tensor]):
    n = [len(a) for a in actors]
    cat = torch.cat([a.transpose(1, 2) for a in actors], 0)
    idcs, count = [], 0
    for k in n:
        idcs.append(torch.arange(count, count + k))
        count += k
    return cat, idcs


# --------------------------------------------------------------------------- lanegcn.py:171-209
def graph_gather(graphs: List[dict]) -> dict:
    offs, count, idcs = [], 0, []
    for g in graphs:
        offs.append(count)
        idcs.append(torch.arange(count, count + g["num_nodes"]))
        count += g["num_nodes"]
    out = {"idcs": idcs, "ctrs": [g["ctrs"] for g in graphs]}
    for key in ("feats", "turn", "control", "intersect"):
        out[key] = torch.cat([g[key] for g in graphs], 0)
    for k1 in ("pre", "suc"):
        out[k1] = []
        for s in range(len(graphs[0]["pre"])):
            out[k1].append(
                {k2: torch.cat([g[k1][s][k2] + o for g, o in zip(graphs, offs)], 0) for k2 in ("u", "v")}
            )
    empty = out["pre"][0]["u"].new_zeros(0)
    for k1 in ("left", "right"):
        out[k1] = {}
        for k2 in ("u", "v"):
            parts = [g[k1][k2] + o for g, o in zip(graphs, offs)]
            out[k1][k2] = torch.cat([p if p.dim() > 0 else empty for p in parts])
    return out


# --------------------------------------------------------------------------- lanegcn.py:331-362 / 448-479
FUSE_EDGE_KEYS = [f"{d}{s}" for s in range(6) for d in ("pre", "suc")]  # ModuleDict order: pre0,suc0,...


def lane_conv_stack(sd, prefix: str, feat: Tensor, graph: dict, num_blocks: int = 4,
                    num_scales: int = 6) -> Tensor:
    """The 4-block LaneConv loop shared by MapNet (lanegcn.py:331-362) and M2M (:448-479)."""
    res = feat
    for i in range(num_blocks):
        temp = F.linear(feat, sd[f"{prefix}.fuse.ctr.{i}.weight"])
        for s in range(num_scales):
            for d in ("pre", "suc"):
                e = graph[d][s]
                temp.index_add_(0, e["u"], F.linear(feat[e["v"]], sd[f"{prefix}.fuse.{d}{s}.{i}.weight"]))
        for d in ("left", "right"):
            e = graph[d]
            if len(e["u"]) > 0:
                temp.index_add_(0, e["u"], F.linear(feat[e["v"]], sd[f"{prefix}.fuse.{d}.{i}.weight"]))
        feat = F.relu(gn(temp, sd, f"{prefix}.fuse.norm.{i}"))
        feat = linear_gn(feat, sd, f"{prefix}.fuse.ctr2.{i}", act=False)
        feat = F.relu(feat + res)
        res = feat
    return feat


def _mlp2(x: Tensor, sd, name: str, act: bool) -> Tensor:
    """nn.Sequential(nn.Linear(2,n), ReLU, layers.Linear(n,n,act=act)) (lanegcn.py:277-286, 644-648)."""
    h = F.relu(F.linear(x, sd[name + ".0.weight"], sd[name + ".0.bias"]))
    return linear_gn(h, sd, name + ".2", act=act)


# --------------------------------------------------------------------------- lanegcn.py:311-363
def map_net(sd, graph: dict, prefix: str = "map_net"):
    ctrs = torch.cat(graph["ctrs"], 0)
    feat = _mlp2(ctrs, sd, prefix + ".input", act=False)
    feat = feat + _mlp2(graph["feats"], sd, prefix + ".seg", act=False)
    feat = F.relu(feat)
    return lane_conv_stack(sd, prefix, feat, graph), graph["idcs"], graph["ctrs"]


# --------------------------------------------------------------------------- lanegcn.py:672-689
def att_pairs(agt_ctrs: List[Tensor], ctx_ctrs: List[Tensor], agt_n: List[int], ctx_n: List[int],
              dist_th: float):
    """Distance-thresholded pair list incl. the empty-scene offset quirk."""
    hi, wi, hc, wc = [], [], 0, 0
    for a, c, na, nc in zip(agt_ctrs, ctx_ctrs, agt_n, ctx_n):
        d = a.view(-1, 1, 2) - c.view(1, -1, 2)
        d = torch.sqrt((d ** 2).sum(2))
        idcs = torch.nonzero(d <= dist_th, as_tuple=False)
        if len(idcs) == 0:
            continue
        hi.append(idcs[:, 0] + hc)
        wi.append(idcs[:, 1] + wc)
        hc += na
        wc += nc
    return torch.cat(hi, 0), torch.cat(wi, 0)


# --------------------------------------------------------------------------- lanegcn.py:662-710
def att(sd, p: str, agts: Tensor, agt_idcs, agt_ctrs, ctx: Tensor, ctx_idcs, ctx_ctrs, dist_th: float):
    res = agts
    if len(ctx) == 0:  # :664-670 — note: no GroupNorm on this path
        a = F.relu(F.linear(agts, sd[p + ".agt.weight"]))
        a = linear_gn(a, sd, p + ".linear", act=False)
        return F.relu(a + res)
    hi, wi = att_pairs(agt_ctrs, ctx_ctrs, [len(x) for x in agt_idcs], [len(x) for x in ctx_idcs], dist_th)
    ac, cc = torch.cat(agt_ctrs, 0), torch.cat(ctx_ctrs, 0)
    dist = _mlp2(ac[hi] - cc[wi], sd, p + ".dist", act=True)
    query = linear_gn(agts[hi], sd, p + ".query", act=True)
    c = torch.cat((dist, query, ctx[wi]), 1)
    c = F.linear(linear_gn(c, sd, p + ".ctx.0", act=True), sd[p + ".ctx.1.weight"])
    a = F.linear(agts, sd[p + ".agt.weight"])
    a.index_add_(0, hi, c)
    a = F.relu(gn(a, sd, p + ".norm"))
    a = linear_gn(a, sd, p + ".linear", act=False)
    return F.relu(a + res)


# --------------------------------------------------------------------------- lanegcn.py:385-407
def a2m(sd, feat, graph, actors, actor_idcs, actor_ctrs, dist_th=7.0):
    meta = torch.cat((graph["turn"], graph["control"].unsqueeze(1), graph["intersect"].unsqueeze(1)), 1)
    feat = linear_gn(torch.cat((feat, meta), 1), sd, "a2m.meta", act=True)
    for i in range(2):
        feat = att(sd, f"a2m.att.{i}", feat, graph["idcs"], graph["ctrs"], actors, actor_idcs, actor_ctrs,
                   dist_th)
    return feat


def m2m(sd, feat, graph):  # lanegcn.py:445-480
    return lane_conv_stack(sd, "m2m", feat, graph)


def m2a(sd, actors, actor_idcs, actor_ctrs, nodes, node_idcs, node_ctrs, dist_th=6.0):  # :502-513
    for i in range(2):
        actors = att(sd, f"m2a.att.{i}", actors, actor_idcs, actor_ctrs, nodes, node_idcs, node_ctrs, dist_th)
    return actors


def a2a(sd, actors, actor_idcs, actor_ctrs, dist_th=100.0):  # :534-545
    for i in range(2):
        actors = att(sd, f"a2a.att.{i}", actors, actor_idcs, actor_ctrs, actors, actor_idcs, actor_ctrs,
                     dist_th)
    return actors


# --------------------------------------------------------------------------- off-path stages (stock torch)
def _conv_gn(x, sd, name, stride=1, act=True):
    """layers.Conv1d (layers.py:40-62): conv k=3 pad 1 no bias -> GN(1) -> optional ReLU."""
    k = sd[name + ".conv.weight"].shape[-1]
    out = F.conv1d(x, sd[name + ".conv.weight"], None, stride, (k - 1) // 2)
    out = gn(out, sd, name + ".norm")
    return F.relu(out) if act else out


def _res1d(x, sd, name, stride=1):
    """layers.Res1d (layers.py:142-190)."""
    out = F.conv1d(x, sd[name + ".conv1.weight"], None, stride, 1)
    out = F.relu(gn(out, sd, name + ".bn1"))
    out = gn(F.conv1d(out, sd[name + ".conv2.weight"], None, 1, 1), sd, name + ".bn2")
    if name + ".downsample.0.weight" in sd:
        x = gn(F.conv1d(x, sd[name + ".downsample.0.weight"], None, stride, 0), sd, name + ".downsample.1")
    return F.relu(out + x)


def actor_net(sd, actors: Tensor) -> Tensor:
    """ActorNet (lanegcn.py:249-263): 3 groups of 2 Res1d, FPN laterals, output Res1d, last step."""
    out, outs = actors, []
    for g in range(3):
        out = _res1d(out, sd, f"actor_net.groups.{g}.0", stride=1 if g == 0 else 2)
        out = _res1d(out, sd, f"actor_net.groups.{g}.1")
        outs.append(out)
    out = _conv_gn(outs[-1], sd, "actor_net.lateral.2", act=False)
    for i in (1, 0):
        out = F.interpolate(out, scale_factor=2, mode="linear", align_corners=False)
        out = out + _conv_gn(outs[i], sd, f"actor_net.lateral.{i}", act=False)
    return _res1d(out, sd, "actor_net.output")[:, :, -1]


def _linear_res(x, sd, name):
    """layers.LinearRes with n_in == n_out (layers.py:193-238)."""
    out = F.relu(gn(F.linear(x, sd[name + ".linear1.weight"]), sd, name + ".norm1"))
    out = gn(F.linear(out, sd[name + ".linear2.weight"]), sd, name + ".norm2")
    return F.relu(out + x)


def pred_net(sd, actors, actor_idcs, actor_ctrs, num_mods=6):
    """PredNet + AttDest (lanegcn.py:602-631, 726-737)."""
    preds = []
    for i in range(num_mods):
        h = _linear_res(actors, sd, f"pred_net.pred.{i}.0")
        preds.append(F.linear(h, sd[f"pred_net.pred.{i}.1.weight"], sd[f"pred_net.pred.{i}.1.bias"]))
    reg = torch.cat([x.unsqueeze(1) for x in preds], 1)
    reg = reg.view(reg.size(0), reg.size(1), -1, 2)
    for idcs, ctrs in zip(actor_idcs, actor_ctrs):
        reg[idcs] = reg[idcs] + ctrs.view(-1, 1, 1, 2)
    dest = reg[:, :, -1].detach()
    agt_ctrs = torch.cat(actor_ctrs, 0)
    n = actors.size(1)
    d = _mlp2((agt_ctrs.unsqueeze(1) - dest).view(-1, 2), sd, "pred_net.att_dest.dist", act=True)
    a = actors.unsqueeze(1).repeat(1, num_mods, 1).view(-1, n)
    feats = linear_gn(torch.cat((d, a), 1), sd, "pred_net.att_dest.agt", act=True)
    h = _linear_res(feats, sd, "pred_net.cls.0")
    cls = F.linear(h, sd["pred_net.cls.1.weight"], sd["pred_net.cls.1.bias"]).view(-1, num_mods)
    cls, sort_idcs = cls.sort(1, descending=True)
    rows = torch.arange(len(sort_idcs)).view(-1, 1).repeat(1, num_mods).view(-1)
    reg = reg[rows, sort_idcs.view(-1)].view(cls.size(0), cls.size(1), -1, 2)
    return {"cls": [cls[i] for i in actor_idcs], "reg": [reg[i] for i in actor_idcs]}


# --------------------------------------------------------------------------- lanegcn.py:127-151
def net_forward(sd: Dict[str, Tensor], data: dict, taps: dict | None = None) -> dict:
    """Full forward on CPU.  ``taps`` (optional dict) receives the per-stage tensors."""
    actors, actor_idcs = actor_gather(data["feats"])
    actor_ctrs = data["ctrs"]
    actors = actor_net(sd, actors)
    graph = graph_gather(to_long(data["graph"]))
    nodes, node_idcs, node_ctrs = map_net(sd, graph)
    t = taps if taps is not None else {}
    t["actor_net"], t["map_net"] = actors, nodes
    nodes = a2m(sd, nodes, graph, actors, actor_idcs, actor_ctrs)
    t["a2m"] = nodes
    nodes = m2m(sd, nodes, graph)
    t["m2m"] = nodes
    actors = m2a(sd, actors, actor_idcs, actor_ctrs, nodes, node_idcs, node_ctrs)
    t["m2a"] = actors
    actors = a2a(sd, actors, actor_idcs, actor_ctrs)
    t["a2a"] = actors
    out = pred_net(sd, actors, actor_idcs, actor_ctrs)
    for i in range(len(out["reg"])):
        out["reg"][i] = torch.matmul(out["reg"][i], data["rot"][i]) + data["orig"][i].view(1, 1, 1, -1)
    return out


# --------------------------------------------------------------------------- LaneRCNN lane-graph layers
def lane_roi(sd, prefix: str, feat: Tensor, graph: dict) -> Tensor:
    """lanercnn.py:388-430 — input Linear+GN+ReLU, then the LaneConv loop (empty edge sets are skipped there by
    `len > 0` guards; index_add_ with empty indices is a no-op, so the shared loop restates it)."""
    feat = linear_gn(feat, sd, prefix + ".input", act=True)
    g = dict(graph)
    for d in ("pre", "suc"):
        g[d] = list(graph[d])
    return lane_conv_stack(sd, prefix, feat, g)


def global_graph_net(sd, prefix: str, feat: Tensor, graph: dict) -> Tensor:
    """lanercnn.py:552-600 — identical to the M2M loop."""
    return lane_conv_stack(sd, prefix, feat, graph)


def rcnn_graph_gather(graphs: List[dict]) -> dict:
    """lanercnn.py:234-277: LaneGCN's graph_gather + num_nodes / counts / pose (= ctrs | feats per scene)."""
    g = graph_gather(graphs)
    g["num_nodes"] = [x["num_nodes"] for x in graphs]
    g["pose"] = [torch.cat([x["ctrs"], x["feats"]], -1) for x in graphs]
    return g


def rcnn_subgraph_gather(subgraphs_in_batch) -> dict:
    """lanercnn.py:122-231 (the keys the forward graph layers read)."""
    flat = [sg for sgs in subgraphs_in_batch for sg in sgs]
    counts, c = [], 0
    for sg in flat:
        counts.append(c)
        c += len(sg["feats"])
    g = {"num_nodes": c, "counts": counts}
    per_scene = lambda f: [torch.cat([f(sg) for sg in sgs], 0) for sgs in subgraphs_in_batch]  # noqa: E731
    g["feats"] = per_scene(lambda sg: sg["feats"])
    g["agent_feat"] = per_scene(lambda sg: sg["agent_feat"].view(1, -1))
    g["ctrs"] = [f[:, :2] for f in g["feats"]]
    g["pose"] = [f[:, :4] for f in g["feats"]]
    g["a2m"] = {"u": torch.cat([sg["a2m"]["u"].long() + i for i, sg in enumerate(flat)]),
                "v": torch.cat([sg["a2m"]["v"].long() + counts[i] for i, sg in enumerate(flat)])}

    def cat_edges(get):
        parts = [get(sg).long() + counts[i] for i, sg in enumerate(flat) if len(get(sg)) > 0]
        return torch.cat(parts) if parts else torch.zeros(0, dtype=torch.long)

    for k1 in ("pre", "suc"):
        g[k1] = [{k2: cat_edges(lambda sg, i=i, k2=k2: sg[k1][i][k2]) for k2 in ("u", "v")} for i in range(6)]
    for k1 in ("left", "right"):
        g[k1] = {k2: cat_edges(lambda sg, k2=k2: sg[k1][k2]) for k2 in ("u", "v")}
    return g


def lane_input(sd, p: str, graph: dict) -> Tensor:
    """lanercnn.py:309-351."""
    m = F.linear(torch.cat(graph["feats"], 0), sd[p + ".map_fc.weight"])
    a = torch.cat(graph["agent_feat"], 0)
    m.index_add_(0, graph["a2m"]["v"], F.linear(a[graph["a2m"]["u"]], sd[p + ".agt_fc.weight"]))
    return F.relu(gn(m, sd, p + ".bn"))


def lane_pooling(sd, p: str, context_feat, context_graph, target_feat, target_graph, dist_th=6.0) -> Tensor:
    """lanercnn.py:463-514 (pairs on centres, relative 4-D pose feature, scatter on the target index wi)."""
    hi, wi = att_pairs(context_graph["ctrs"], target_graph["ctrs"], [len(x) for x in context_graph["ctrs"]],
                       [len(x) for x in target_graph["ctrs"]], dist_th)
    cp, tp = torch.cat(context_graph["pose"], 0), torch.cat(target_graph["pose"], 0)
    d = F.relu(F.linear(cp[hi] - tp[wi], sd[p + ".relpose.0.weight"], sd[p + ".relpose.0.bias"]))
    c = linear_gn(torch.cat([context_feat[hi], d], -1), sd, p + ".ctx.0", act=True)
    c = F.linear(c, sd[p + ".ctx.1.weight"])
    t = F.linear(target_feat, sd[p + ".input.weight"])
    t.index_add_(0, wi, c)
    t = F.relu(gn(t, sd, p + ".norm"))
    t = linear_gn(t, sd, p + ".mlp.0", act=True)
    t = linear_gn(t, sd, p + ".mlp.1", act=False)
    return F.relu(t + target_feat)


def interactor(sd, p: str, graph: dict, subgraph: dict, roi_feat: Tensor) -> Tensor:
    """lanercnn.py:630-642."""
    g_in = _mlp2(torch.cat(graph["ctrs"], 0), sd, p + ".input", act=False)
    g_in = F.relu(g_in + _mlp2(graph["feats"], sd, p + ".seg", act=False))
    g_feat = lane_pooling(sd, p + ".roi2graph", roi_feat, subgraph, g_in, graph)
    g_feat = lane_conv_stack(sd, p + ".global_graph_net", g_feat, graph)
    return lane_pooling(sd, p + ".graph2roi", g_feat, graph, roi_feat, subgraph)
