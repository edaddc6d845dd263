"""Timing of one aggregate-first LaneConv block under arbitrary lgcn_debug_flags combinations (argv[1:] = flag values):
quick what-if probes next to tools/ablate_fused.py.  us per call of lgcn_laneconv_stack_planned with n_blocks = 2
(no final copy), divided by 2."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanegcn_b200 import _C, synth
from lanegcn_b200 import lanegcn as L
lib = _C.lib()
dev = torch.device("cuda", 0)
B = int(os.environ.get("PROBE_B", "128"))
data = synth.collate(synth.make_scenes(B, "argo-1.5k"))
pg = L.graph_gather(data["graph"])["_packed"]
M, K = pg.n_nodes, pg.n_keys
g = torch.Generator().manual_seed(0)
wpack = (torch.randn(2 * lib.lgcn_laneconv_wpack_floats(K), generator=g) / 11).to(dev)
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
        _C.check(lib.lgcn_laneconv_stack_planned(feat.data_ptr(), plan.data_ptr(), pg.n_edges, K, 2, wpack.data_ptr(), M,
                                                 ws.data_ptr(), sp))
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, e0.elapsed_time(e1) * 1e3)
    lib.lgcn_debug_flags(0)
    return best / 2


print(f"nodes {M}; us per block (multi_sum + fused kernel; + half a weight split)")
for fl in [int(a) for a in sys.argv[1:]] or [0]:
    print(f"flags {fl:6d}  {t(fl):8.1f}")
