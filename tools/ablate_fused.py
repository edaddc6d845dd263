"""Ablation timing of the aggregate-first LaneConv kernel (laneconv_fused.cu) on the GPU box, on the real synthetic
batch-128 lane graph: one block (n_blocks = 1) under the lgcn_debug_flags switches
(1 no stores, 4 no MMAs, 8 no loads, 32 no A conversion / tcgen05.st, 64 no accumulator flushes, 128 no weight TMA)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanegcn_b200 import _C, synth
from lanegcn_b200 import lanegcn as L
lib = _C.lib()
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
data = synth.collate(synth.make_scenes(B, "argo-1.5k"))
graph = L.graph_gather(data["graph"])
pg = graph["_packed"]
M, K = pg.n_nodes, pg.n_keys
g = torch.Generator().manual_seed(0)
wpack = (torch.randn(lib.lgcn_laneconv_wpack_floats(K), generator=g) / 11).to(dev)
X = torch.randn(M, 128, generator=g).to(dev)
feat = X.clone()
ws = torch.empty(lib.lgcn_laneconv_planned_workspace_bytes(M, pg.n_edges, K), dtype=torch.uint8, device=dev)
plan = pg.plan()
sp = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(flags, n=6):
    lib.lgcn_debug_flags(flags)
    best = 1e9
    for it in range(n):
        feat.copy_(X)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _C.check(lib.lgcn_laneconv_stack_planned(feat.data_ptr(), plan.data_ptr(), pg.n_edges, K, 1, wpack.data_ptr(), M,
                                                 ws.data_ptr(), sp))
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, e0.elapsed_time(e1) * 1e3)
    lib.lgcn_debug_flags(0)
    return best


print(f"nodes {M} edges {pg.n_edges}; us per block (split + multi_sum + fused kernel + final copy)")
for name, fl in [("full", 0), ("no stores", 1), ("no flushes", 64), ("no loads", 8), ("no loads, no conversion", 8 | 32),
                 ("no MMAs", 4), ("no MMAs, no loads", 4 | 8), ("no MMAs, no loads, no conversion", 4 | 8 | 32),
                 ("skeleton: + no flushes, no stores", 4 | 8 | 32 | 64 | 1), ("no conversion only", 32), ("no weight TMA", 128), ("MMAs only: no weight TMA, no loads, no conversion", 128 | 8 | 32),
                 ("A feed only: no weight TMA, no MMAs", 128 | 4), ("barriers only", 128 | 4 | 8 | 32 | 64 | 1)]:
    print(f"{name:45s} {t(fl):8.1f}")
