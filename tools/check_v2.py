"""k_laneconv_v2 against the first-generation kernel (debug flag 2048) on the same inputs: max abs difference of one
LaneConv block, plus a per-column-group breakdown to localise layout bugs."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanegcn_b200 import _C, synth
from lanegcn_b200 import lanegcn as L
lib = _C.lib()
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
data = synth.collate(synth.make_scenes(B, "argo-1.5k"))
pg = L.graph_gather(data["graph"])["_packed"]
M, K = pg.n_nodes, pg.n_keys
g = torch.Generator().manual_seed(0)
wpack = (torch.randn(lib.lgcn_laneconv_wpack_floats(K), generator=g) / 11).to(dev)
X = torch.randn(M, 128, generator=g).to(dev)
ws = torch.empty(lib.lgcn_laneconv_planned_workspace_bytes(M, pg.n_edges, K), dtype=torch.uint8, device=dev)
plan = pg.plan()
sp = torch.cuda.current_stream().cuda_stream


def run(flags):
    feat = X.clone()
    lib.lgcn_debug_flags(flags)
    _C.check(lib.lgcn_laneconv_stack_planned(feat.data_ptr(), plan.data_ptr(), pg.n_edges, K, 1, wpack.data_ptr(), M, ws.data_ptr(), sp))
    torch.cuda.synchronize()
    lib.lgcn_debug_flags(0)
    return feat


a, b = run(2048), run(0)
d = (a - b).abs()
print(f"rows {M}: max abs diff {d.max().item():.3e} (old kernel output max {a.abs().max().item():.3f})")
if d.max().item() > 1e-4:
    print("per 16-column group max diff:", [round(d[:, c:c + 16].max().item(), 4) for c in range(0, 128, 16)])
    print("per 128-row tile (first 6):", [round(d[t * 128:(t + 1) * 128].max().item(), 4) for t in range(6)])
    bad_rows = (d.max(dim=1).values > 1e-4).nonzero().flatten()
    print("bad rows:", bad_rows.numel(), "first:", bad_rows[:16].tolist())
    r = int(bad_rows[0])
    print("row", r, "old:", [round(v, 3) for v in a[r, :20].tolist()])
    print("row", r, "new:", [round(v, 3) for v in b[r, :20].tolist()])
    # is the new row a column permutation of the old one?
    sa, sb = a[r].sort().values, b[r].sort().values
    print("sorted-row diff (permutation check):", (sa - sb).abs().max().item())
