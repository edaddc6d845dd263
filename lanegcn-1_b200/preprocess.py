"""Graph construction right before the forward path, on the GPU.

Same names and results as the reference's host / dense-torch code:
  ``scale0_edges``   the scale-0 ``pre`` / ``suc`` node edges of ``ArgoDataset.get_lane_graph`` (data.py:272-295) from
                     the node->lane map and the lane-level pre / suc pairs (data.py:302-317);
  ``preprocess``     the ``left`` / ``right`` node edges of preprocess_data.py:287-392 (same signature and return value:
                     int16 numpy ``u``, ``v``), without the dense N x N matrices;
  ``build_edges``    scale 0 + the dilated scales 1..S-1 (``lanegcn.dilated_nbrs``, data.py:520-534): a scene's complete
                     ``pre`` / ``suc`` lists from its lane topology, all on the device.
Integer outputs are bit-exact with the reference (tests/test_gpu_preprocess.py against goldens made by the reference).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List

import numpy as np
import torch

from . import _C
from .lanegcn import _target_device, dilated_nbrs


def _dev64(x, dev) -> torch.Tensor:
    return torch.as_tensor(x).to(dev, torch.int64).contiguous()


def scale0_edges(lane_idcs, pre_pairs, suc_pairs, device=None):
    """-> (pre0, suc0), each ``{"u","v"}`` int64 CUDA tensors in the reference's order (data.py:272-295)."""
    lib = _C.lib()
    lane = torch.as_tensor(lane_idcs)
    dev = device or _target_device(lane)
    lane = _dev64(lane, dev)
    n = lane.numel()
    n_lanes = int(lane[-1].item()) + 1 if n else 0
    out = []
    with torch.cuda.device(dev):
        ws = torch.empty(lib.lgcn_scale0_workspace_bytes(n_lanes), dtype=torch.uint8, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        for is_suc, pairs in ((0, pre_pairs), (1, suc_pairs)):
            p = _dev64(np.asarray(pairs).reshape(-1, 2) if not torch.is_tensor(pairs) else pairs.reshape(-1, 2), dev)
            m = max(n - n_lanes, 0) + p.shape[0]
            u = torch.empty(m, dtype=torch.int64, device=dev)
            v = torch.empty(m, dtype=torch.int64, device=dev)
            _C.check(lib.lgcn_scale0_edges(lane.data_ptr(), n, n_lanes, p.data_ptr() if p.numel() else None, p.shape[0],
                                           is_suc, u.data_ptr(), v.data_ptr(), ws.data_ptr(), err.data_ptr(),
                                           _C.stream_ptr()), "scale0_edges")
            if int(err.item()):
                raise RuntimeError("lanegcn_b200: scale0_edges: the lane pairs must be sorted by their first column "
                                   "(the reference appends them lane by lane, data.py:302-317)")
            out.append({"u": u, "v": v})
    return out[0], out[1]


def build_edges(lane_idcs, pre_pairs, suc_pairs, num_scales: int = 6, device=None):
    """-> (pre, suc): lists of ``num_scales`` ``{"u","v"}`` (scale 0 + dilated scales), int64 CUDA tensors."""
    pre0, suc0 = scale0_edges(lane_idcs, pre_pairs, suc_pairs, device)
    n = int(torch.as_tensor(lane_idcs).numel())
    return [pre0] + dilated_nbrs(pre0, n, num_scales), [suc0] + dilated_nbrs(suc0, n, num_scales)


def _side(graph: Dict, key: str, cross_dist: float, dev) -> Dict[str, np.ndarray]:
    lib = _C.lib()
    ctrs = torch.as_tensor(graph["ctrs"]).to(dev, torch.float32).contiguous()
    feats = torch.as_tensor(graph["feats"]).to(dev, torch.float32).contiguous()
    lane = _dev64(graph["lane_idcs"], dev)
    n = lane.numel()
    n_lanes = int(lane[-1].item()) + 1 if n else 0
    pairs = [_dev64(torch.as_tensor(graph[k]).reshape(-1, 2), dev) for k in (key + "_pairs", "pre_pairs", "suc_pairs")]
    u = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    v = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    ws = torch.empty(lib.lgcn_side_edges_workspace_bytes(n, n_lanes), dtype=torch.uint8, device=dev)
    cnt = ctypes.c_int64(0)
    ptr = lambda t: t.data_ptr() if t.numel() else None  # noqa: E731
    _C.check(lib.lgcn_side_edges(ctrs.data_ptr(), feats.data_ptr(), lane.data_ptr(), n, n_lanes, ptr(pairs[0]),
                                 pairs[0].shape[0], ptr(pairs[1]), pairs[1].shape[0], ptr(pairs[2]), pairs[2].shape[0],
                                 float(cross_dist), u.data_ptr(), v.data_ptr(), ws.data_ptr(), ctypes.byref(cnt),
                                 _C.stream_ptr()), "side_edges")
    k = cnt.value
    return {"u": u[:k].cpu().numpy().astype(np.int16), "v": v[:k].cpu().numpy().astype(np.int16)}


def preprocess(graph: Dict, cross_dist: float, cross_angle=None) -> Dict:
    """preprocess_data.py:287-392: ``{"left": {u, v}, "right": {u, v}, "idx": graph["idx"]}`` with int16 numpy arrays.
    ``graph`` needs ``ctrs, feats, lane_idcs, pre_pairs, suc_pairs, left_pairs, right_pairs`` (numpy arrays or tensors on
    any device)."""
    if cross_angle is not None:
        raise NotImplementedError("lanegcn_b200: preprocess(cross_angle=...) — the reference calls it with cross_dist only "
                                  "(preprocess_data.py:250)")
    dev = _target_device(torch.as_tensor(graph["ctrs"]))
    with torch.cuda.device(dev):
        out = {"left": _side(graph, "left", cross_dist, dev), "right": _side(graph, "right", cross_dist, dev)}
    out["idx"] = graph.get("idx")
    return out
