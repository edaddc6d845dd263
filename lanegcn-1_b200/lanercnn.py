"""LaneRCNN clients of the same kernels (reference ``lanercnn.py``; BASELINE config 5).

Drop-ins with the reference's parameter names for the graph layers of LaneRCNN: ``LaneRoI`` (lanercnn.py:354-430:
Linear+GN+ReLU, then the 4-block LaneConv loop with per-key ``len > 0`` guards), ``GlobalGraphNet`` (:517-600: the
M2M loop), ``LaneInput`` (:280-351), ``LanePooling`` (:433-514), ``Interactor`` (:603-642) and the batching helpers
``graph_gather`` (:234-277) / ``subgraph_gather`` (:122-231).  Everything float runs on the same C-ABI kernels as
LaneGCN.  ``Decode`` (python NMS + trajectory sampling) and the losses are outside the forward graph path.
"""
from __future__ import annotations

from typing import Dict

import torch
from torch import Tensor, nn

from . import _C
from .blocks import Linear
from .lanegcn import (C_, _LaneConvStack, _lin_ws, _as_scene_list, _f32c, _need_cuda, _packed_of, _target_device,
                      build_pair_lists, scene_list)


def as_cat(lst) -> Tensor:
    """Batched tensor of a per-scene list (SceneList keeps it; plain lists are concatenated)."""
    if torch.is_tensor(lst):
        return lst
    return lst.cat if getattr(lst, "cat", None) is not None else torch.cat(list(lst), 0)


class LaneRoI(_LaneConvStack):
    """lanercnn.py:354-430.  Empty edge sets contribute nothing, which is what the ``len(...) > 0`` guards do."""

    def __init__(self, config, input_dim):
        super().__init__()
        if input_dim != C_ or config["n_map"] != C_:
            raise ValueError("lanegcn_b200: LaneRoI kernels are built for input_dim = n_map = 128")
        self.input = Linear(input_dim, config["n_map"], act=True)
        self._init_fuse(config)
        self.relu = nn.ReLU(inplace=True)

    @torch.no_grad()
    def forward(self, feat: Tensor, graph: Dict) -> Tensor:
        _need_cuda(feat, "feat")
        feat = _f32c(feat)
        out = torch.empty_like(feat)
        _C.check(_C.lib().lgcn_linear128_ws(feat.data_ptr(), None, None, None, None, None, 1, None, 0,
                                         self.input.linear.weight.data_ptr(), 1, self.input.norm.weight.data_ptr(),
                                         self.input.norm.bias.data_ptr(), None, _C.EPI_GN | _C.EPI_RELU1,
                                         out.data_ptr(), C_, feat.shape[0], _lin_ws(), _C.stream_ptr()), "linear128(LaneRoI.input)")
        return self._stack(out, _packed_of(graph))


class GlobalGraphNet(_LaneConvStack):
    """lanercnn.py:517-600: identical to LaneGCN's M2M loop."""

    def __init__(self, config):
        super().__init__()
        self._init_fuse(config)
        self.relu = nn.ReLU(inplace=True)

    @torch.no_grad()
    def forward(self, feat: Tensor, graph: Dict) -> Tensor:
        if len(graph["feats"]) == 0 or len(graph["pre"][-1]["u"]) == 0 or len(graph["suc"][-1]["u"]) == 0:
            return (graph["feats"].new_zeros(0),)  # the reference's degenerate branch returns a 1-tuple (:553-563)
        _need_cuda(feat, "feat")
        return self._stack(_f32c(feat).clone(), _packed_of(graph))


# --------------------------------------------------------------------------- batching helpers
def graph_gather(graphs):
    """lanercnn.py:234-277: LaneGCN's graph_gather plus ``num_nodes``, ``counts`` and ``pose`` (= ctrs | feats)."""
    from .lanegcn import graph_gather as _gg

    g = _gg(graphs)
    sizes = [len(x) for x in g["idcs"]]
    g["num_nodes"] = sizes
    g["counts"] = g["ctrs"].off[:-1]
    pose = torch.cat((g["ctrs"].cat, g["feats"]), 1).contiguous()
    g["pose"] = scene_list(pose, sizes, g["ctrs"].off_dev)
    return g


def subgraph_gather(subgraphs_in_batch, device=None):
    """lanercnn.py:122-231: batch the per-agent lane-RoI sub-graphs of every scene (node offsets per RoI, a2m
    offsets per RoI / per node).  Host-side list walking + torch ops (index plumbing of a client module, not on the
    benchmarked path); the result carries the same keys as the reference plus the private CSR (``_packed``)."""
    dev = device or _target_device(subgraphs_in_batch[0][0]["feats"])
    counts, count, spans, start, n_agts, roi_sizes = [], 0, [], 0, [], []
    for sgs in subgraphs_in_batch:
        n_agts.append(len(sgs))
        n_this = 0
        for sg in sgs:
            counts.append(count)
            n = len(sg["feats"])
            roi_sizes.append(n)
            count += n
            n_this += n
        spans.append([start, start + n_this])
        start += n_this
    g = dict()
    g["num_nodes"], g["counts"], g["batch_spans"], g["num_atgs_per_batch"] = count, counts, spans, n_agts
    g["node_idcs"] = torch.arange(count, device=dev)
    ends = counts[1:] + [count]
    g["roi_spans"] = [[a, b] for a, b in zip(counts, ends)]
    first, k = [], 0
    for n in n_agts:
        first.append(k)
        k += n
    g["interest_roi"] = torch.tensor(first, dtype=torch.long)
    flat = [sg for sgs in subgraphs_in_batch for sg in sgs]
    feats = torch.cat([sg["feats"].float() for sg in flat], 0).to(dev)
    agt = torch.stack([sg["agent_feat"].float().reshape(-1) for sg in flat], 0).to(dev)
    batch_nodes = [b - a for a, b in spans]
    g["feats"] = scene_list(feats, batch_nodes)
    g["agent_feat"] = scene_list(agt, n_agts)
    g["ctrs"] = scene_list(feats[:, :2].contiguous(), batch_nodes, g["feats"].off_dev)
    g["dirs"] = scene_list(feats[:, 2:4].contiguous(), batch_nodes, g["feats"].off_dev)
    g["pose"] = scene_list(feats[:, :4].contiguous(), batch_nodes, g["feats"].off_dev)
    g["agent_vel"] = [sg["agent_vel"] for sg in flat]
    g["a2m"] = {
        "u": torch.cat([sg["a2m"]["u"].long() + i for i, sg in enumerate(flat)]).to(dev),
        "v": torch.cat([sg["a2m"]["v"].long() + counts[i] for i, sg in enumerate(flat)]).to(dev),
    }

    def cat_edges(get):
        parts = [get(sg).long() + counts[i] for i, sg in enumerate(flat) if len(get(sg)) > 0]
        return torch.cat(parts).to(dev) if parts else torch.zeros(0, dtype=torch.long, device=dev)

    for k1 in ("pre", "suc"):
        g[k1] = [{k2: cat_edges(lambda sg, i=i, k2=k2: sg[k1][i][k2]) for k2 in ("u", "v")} for i in range(6)]
    for k1 in ("left", "right"):
        g[k1] = {k2: cat_edges(lambda sg, k2=k2: sg[k1][k2]) for k2 in ("u", "v")}
    return g


def _scatter_csr(dst: Tensor, src_index: Tensor, n_rows: int, n_src: int):
    """Stable destination-sorted CSR of a scatter ``out[dst[e]] += rows[src_index[e]]`` (what index_add_ does on an
    UNSORTED destination index, in edge order): rowptr int32[n_rows+1], col int32[E] = src_index in CSR order."""
    lib = _C.lib()
    dst, src_index = dst.long().contiguous(), src_index.long().contiguous()
    e = int(dst.numel())
    rowptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dst.device)
    col = torch.empty(max(e, 1), dtype=torch.int32, device=dst.device)
    err = torch.empty(1, dtype=torch.int32, device=dst.device)
    ws = torch.empty(lib.lgcn_csr_workspace_bytes(n_rows, e), dtype=torch.uint8, device=dst.device)
    _C.check(lib.lgcn_scatter_csr_build(dst.data_ptr(), src_index.data_ptr(), e, n_rows, n_src, rowptr.data_ptr(),
                                        col.data_ptr(), ws.data_ptr(), err.data_ptr(), _C.stream_ptr()), "scatter_csr_build")
    return rowptr, col


def _lin(x, w, gamma=None, beta=None, res=None, flags=0, idx=None):
    out = torch.empty(x.shape[0] if idx is None else idx.shape[0], C_, dtype=torch.float32, device=x.device)
    _C.check(_C.lib().lgcn_linear128_ws(x.data_ptr(), _C.ptr(idx), None, None, None, None, 1, None, 0, w.data_ptr(), 1,
                                     _C.ptr(gamma), _C.ptr(beta), _C.ptr(res), flags, out.data_ptr(), C_, out.shape[0], _lin_ws(),
                                     _C.stream_ptr()), "linear128")
    return out


def _gather_rows(base, blocks, rowptr, col, gamma, beta):
    out = torch.empty_like(base)
    _C.check(_C.lib().lgcn_gather_rows_gn_relu(base.data_ptr(), C_, blocks.data_ptr(), rowptr.data_ptr(), col.data_ptr(),
                                               gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), base.shape[0],
                                               _C.stream_ptr()), "gather_rows_gn_relu")
    return out


def _pad_k(x: Tensor, k: int = C_) -> Tensor:
    """[*, K<128] -> [*, 128] zero-padded (small-K Linears of LaneInput run on the 128-wide GEMM)."""
    return torch.nn.functional.pad(x, (0, k - x.shape[1])).contiguous()


# --------------------------------------------------------------------------- LaneInput / LanePooling / Interactor
class LaneInput(nn.Module):
    """lanercnn.py:280-351: map_fc(feats[.,8]); index_add_(a2m.v, agt_fc(agent_feat[.,80][a2m.u])); GN; ReLU."""

    def __init__(self, config):
        super().__init__()
        n = config["n_map"]
        self.map_fc = nn.Linear(8, n, bias=False)
        self.agt_fc = nn.Linear(80, n, bias=False)
        self.bn = nn.GroupNorm(1, n)
        self.relu = nn.ReLU(inplace=True)

    @torch.no_grad()
    def forward(self, graph):
        feats, agt = as_cat(graph["feats"]), as_cat(graph["agent_feat"])
        _need_cuda(feats, "graph['feats']")
        t = _lin(_pad_k(feats), _pad_k(self.map_fc.weight))              # [nodes,128]
        a = _lin(_pad_k(agt), _pad_k(self.agt_fc.weight))                # [agents,128]; row-wise, so per agent
        rowptr, col = _scatter_csr(graph["a2m"]["v"], graph["a2m"]["u"], t.shape[0], a.shape[0])
        return _gather_rows(t, a, rowptr, col, self.bn.weight, self.bn.bias)


class LanePooling(nn.Module):
    """lanercnn.py:433-514: Att-like pooling between two lane graphs (pairs on centres, 4-D relative pose feature,
    256-wide context MLP, scatter on the TARGET index ``wi``, which is not sorted)."""

    def __init__(self, in_dim: int = 128, out_dim: int = 128) -> None:
        super().__init__()
        d = 128  # the reference ignores its arguments (lanercnn.py:439)
        self.input = nn.Linear(d, d, bias=False)
        self.relpose = nn.Sequential(nn.Linear(4, d), nn.ReLU(inplace=True))
        self.ctx = nn.Sequential(Linear(2 * d, d), nn.Linear(d, d, bias=False))
        self.mlp = nn.Sequential(Linear(d, d), Linear(d, d, act=False))
        self.norm = nn.GroupNorm(1, d)
        self.relu = nn.ReLU(inplace=True)

    @torch.no_grad()
    def forward(self, context_feat, context_graph, target_feat, target_graph, dist_th=6.0, g2r=False):
        lib, st = _C.lib(), _C.stream_ptr()
        context_feat, target_feat = _f32c(context_feat), _f32c(target_feat)
        _need_cuda(context_feat, "context_feat")
        c_ctrs, t_ctrs = _as_scene_list(context_graph["ctrs"]), _as_scene_list(target_graph["ctrs"])
        pairs = build_pair_lists([(c_ctrs, t_ctrs, dist_th)], want_int64=True)[0]   # hi = context, wi = target
        if pairs.n_pairs == 0:
            raise RuntimeError("lanegcn_b200: LanePooling found no pair within dist_th (reference: torch.cat([]) raises)")
        P = pairs.n_pairs
        c_pose, t_pose = as_cat(context_graph["pose"]), as_cat(target_graph["pose"])
        dist = torch.empty(P, C_, dtype=torch.float32, device=context_feat.device)
        _C.check(lib.lgcn_mlp4_in(_f32c(c_pose).data_ptr(), pairs.hi.data_ptr(), _f32c(t_pose).data_ptr(),
                                  pairs.wi.data_ptr(), self.relpose[0].weight.data_ptr(), self.relpose[0].bias.data_ptr(),
                                  dist.data_ptr(), P, st), "mlp4_in")
        c0 = self.ctx[0]
        ctx = torch.empty(P, C_, dtype=torch.float32, device=dist.device)
        _C.check(lib.lgcn_linear128_ws(context_feat.data_ptr(), pairs.hi.data_ptr(), dist.data_ptr(), None, None, None, 2,
                                    None, 0, c0.linear.weight.data_ptr(), 1, c0.norm.weight.data_ptr(),
                                    c0.norm.bias.data_ptr(), None, _C.EPI_GN | _C.EPI_RELU1, ctx.data_ptr(), C_, P, _lin_ws(), st),
                 "linear128(ctx.0)")
        ctx = _lin(ctx, self.ctx[1].weight)
        t = _lin(target_feat, self.input.weight)
        rowptr, col = _scatter_csr(pairs.wi64, torch.arange(P, device=t.device), t.shape[0], P)
        t = _gather_rows(t, ctx, rowptr, col, self.norm.weight, self.norm.bias)
        m0, m1 = self.mlp[0], self.mlp[1]
        t = _lin(t, m0.linear.weight, m0.norm.weight, m0.norm.bias, None, _C.EPI_GN | _C.EPI_RELU1)
        return _lin(t, m1.linear.weight, m1.norm.weight, m1.norm.bias, target_feat, _C.EPI_GN | _C.EPI_RES | _C.EPI_RELU2)


class Interactor(nn.Module):
    """lanercnn.py:603-642: global-graph input MLPs, LanePooling roi->graph, GlobalGraphNet, LanePooling graph->roi."""

    def __init__(self, config):
        super().__init__()
        n = config["n_map"]
        self.input = nn.Sequential(nn.Linear(2, n), nn.ReLU(inplace=True), Linear(n, n, act=False))
        self.seg = nn.Sequential(nn.Linear(2, n), nn.ReLU(inplace=True), Linear(n, n, act=False))
        self.relu = nn.ReLU(inplace=True)
        self.roi2graph = LanePooling(128, 128)
        self.global_graph_net = GlobalGraphNet(config)
        self.graph2roi = LanePooling(128, 128)

    @torch.no_grad()
    def forward(self, graph, subgraph, roi_feat):
        lib, st = _C.lib(), _C.stream_ptr()
        ctrs, feats = as_cat(graph["ctrs"]), _f32c(graph["feats"])
        n = ctrs.shape[0]
        hid = torch.empty(n, C_, dtype=torch.float32, device=ctrs.device)
        a, g_in = torch.empty_like(hid), torch.empty_like(hid)
        for src, mlp, res, flags, out in ((ctrs, self.input, None, _C.EPI_GN, a),
                                          (feats, self.seg, a, _C.EPI_GN | _C.EPI_RES | _C.EPI_RELU2, g_in)):
            _C.check(lib.lgcn_mlp2_in(_f32c(src).data_ptr(), None, None, None, mlp[0].weight.data_ptr(),
                                      mlp[0].bias.data_ptr(), hid.data_ptr(), n, st), "mlp2_in")
            _C.check(lib.lgcn_linear128_ws(hid.data_ptr(), None, None, None, None, None, 1, None, 0,
                                        mlp[2].linear.weight.data_ptr(), 1, mlp[2].norm.weight.data_ptr(),
                                        mlp[2].norm.bias.data_ptr(), _C.ptr(res), flags, out.data_ptr(), C_, n, _lin_ws(), st),
                     "linear128")
        graph_feat = self.roi2graph(roi_feat, subgraph, g_in, graph)
        graph_feat = self.global_graph_net(graph_feat, graph)
        return self.graph2roi(graph_feat, graph, roi_feat, subgraph)


class Net(nn.Module):
    """The forward GRAPH path of lanercnn.Net (lanercnn.py:85-119) up to the RoI features that ``Decode`` consumes:
    graph_gather / subgraph_gather -> LaneInput -> LaneRoI -> Interactor -> LaneRoI.  Sub-module names follow the
    reference (``input``, ``roi_net1``, ``interactor``, ``roi_net2``) so its checkpoints load by key with
    ``strict=False``; ``Decode`` (python NMS + polynomial trajectory sampling) is outside the graph path."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.input = LaneInput(config)
        self.roi_net1 = LaneRoI(config, input_dim=config["n_map"])
        self.interactor = Interactor(config)
        self.roi_net2 = LaneRoI(config, input_dim=config["n_map"])

    @torch.no_grad()
    def forward(self, data: Dict):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("lanegcn_b200: move the model to a CUDA device first (there is no CPU path)")
        with torch.cuda.device(dev):
            graph = graph_gather(data["graph"])
            roi = subgraph_gather(data["subgraphs"], dev)
            return {"roi_feat": self.forward_graphs(graph, roi), "graph": graph, "graphRoI": roi}

    @torch.no_grad()
    def forward_graphs(self, graph: Dict, roi: Dict) -> Tensor:
        """The graph layers on already batched (device-resident) graphs: lanercnn.py:97-112."""
        feat = self.input(roi)
        feat = self.roi_net1(feat, roi)
        feat = self.interactor(graph, roi, feat)
        return self.roi_net2(feat, roi)
