"""bench.py --config 5: LaneRCNN's forward graph layers (LaneInput -> LaneRoI -> Interactor -> LaneRoI, lanercnn.py:97-112)
on synthetic scenes with per-agent lane-RoI sub-graphs (synth.make_lane_rois), through the same kernels as LaneGCN."""
from __future__ import annotations

import json
import os
import time

import torch

from . import _C, synth
from . import lanegcn as L
from . import lanercnn as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, dev, clocks, metric_of, workload_of):
    lib = _C.lib()
    B = args.batch or 32
    metric, unit = metric_of(5, B)
    scenes = synth.make_scenes(B, "argo-1.5k", seed0=0)
    for s in scenes:
        s["subgraphs"] = synth.make_lane_rois(s)
    data = synth.collate(scenes)
    net = R.Net(L.config)
    shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(synth.seeded_state_dict(shapes, 9))
    net = net.to(dev).eval()
    n_roi = sum(len(s["subgraphs"]) for s in scenes)
    n_roi_nodes = sum(sg["num_nodes"] for s in scenes for sg in s["subgraphs"])
    n_nodes = sum(int(s["graph"]["num_nodes"]) for s in scenes)
    with torch.cuda.device(dev):
        graph = R.graph_gather(data["graph"])
        roi = R.subgraph_gather(data["subgraphs"], dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        W = max(args.warmup, 3)
        for _ in range(W):
            net.forward_graphs(graph, roi)
        torch.cuda.synchronize()
        if clocks:
            clocks.wait_first()
            clocks.mark_start()
        l0 = lib.lgcn_launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in ev:
            flush.zero_()
            a.record()
            net.forward_graphs(graph, roi)
            b.record()
        torch.cuda.synchronize()
        if clocks:
            clocks.mark_end()
        launches = lib.lgcn_launch_count() - l0
        ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
        clk = clocks.stop() if clocks else None
        n_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            host = net(data)["roi_feat"].cpu()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
    h2d = sum(t.numel() * t.element_size() for s in (data["graph"], [sg for sgs in data["subgraphs"] for sg in sgs])
              for d in s for t in _tensors(d))
    return {
        "metric": metric, "value": round(B / (ms / 1e3), 2), "unit": unit, "n_gpus": 1, "steps": args.steps, "warmup": W,
        "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_of(5, B), "global_batch": B, "graph_nodes": n_nodes, "lane_rois": n_roi,
                   "roi_nodes": n_roi_nodes, "l2": "256 MiB flush between timed steps",
                   "path": "module path (LaneInput, 2 x LaneRoI = 8 aggregate-first LaneConv blocks on the RoI graph, "
                           "Interactor = 2 LanePooling + GlobalGraphNet's 4 blocks on the scene graph)"},
        "e2e": {"value": round(B / (e2e_ms / 1e3), 2), "unit": unit, "ms_per_step": round(e2e_ms, 3),
                "how": "Net.forward(host dict): graph_gather + subgraph_gather (host walk + H2D) + graph layers + D2H of the "
                       "RoI features, per step, not overlapped",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(host.numel() * 4)},
        "gpu_launches": int(launches), "clocks": clk,
    }


def _tensors(d):
    for v in d.values():
        if torch.is_tensor(v):
            yield v
        elif isinstance(v, dict):
            yield from _tensors(v)
        elif isinstance(v, list):
            for e in v:
                if isinstance(e, dict):
                    yield from _tensors(e)
