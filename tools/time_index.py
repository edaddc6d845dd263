"""Stand-alone device times of the index kernels (CSR build, gather plan, pair lists) at batch B — nothing else on the
GPU (no ActorNet beside them), torch.profiler / CUPTI durations, median of 5 runs."""
import json, os, sys, statistics
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanegcn_b200 import synth
from lanegcn_b200 import lanegcn as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
data = synth.collate(synth.make_scenes(B, "argo-1.5k"))
graphs = data["graph"]
times = {}
for it in range(7):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        g = L.graph_gather(graphs)
        g["_packed"].plan()
        torch.cuda.synchronize()
    if it < 2:
        continue
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            times.setdefault(e.name[:60], []).append(e.time_range.end - e.time_range.start)
for k, v in sorted(times.items(), key=lambda kv: -sum(kv[1])):
    n = len(v) // 5
    print(f"{statistics.median(v):9.1f} us  x{n:2d} per run  {k}")
