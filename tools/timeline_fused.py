"""Per-stage timeline of CTA 0 of the aggregate-first LaneConv kernel (debug flag 256): where the MMA warp and one
producer warp spend their cycles.  Prints deltas between consecutive stamps, averaged over steady-state stages."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanegcn_b200 import _C, synth
from lanegcn_b200 import lanegcn as L
lib = _C.lib()
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
extra = int(sys.argv[2]) if len(sys.argv) > 2 else 0
data = synth.collate(synth.make_scenes(B, "argo-1.5k"))
pg = L.graph_gather(data["graph"])["_packed"]
M, K = pg.n_nodes, pg.n_keys
g = torch.Generator().manual_seed(0)
wpack = (torch.randn(lib.lgcn_laneconv_wpack_floats(K), generator=g) / 11).to(dev)
feat = torch.randn(M, 128, generator=g).to(dev)
ws = torch.empty(lib.lgcn_laneconv_planned_workspace_bytes(M, pg.n_edges, K), dtype=torch.uint8, device=dev)
tl = torch.zeros(1024, 8, dtype=torch.int64, device=dev)
lib.lgcn_debug_timeline(tl.data_ptr())
sp = torch.cuda.current_stream().cuda_stream
for it in range(2):
    lib.lgcn_debug_flags(256 | extra)
    _C.check(lib.lgcn_laneconv_stack_planned(feat.data_ptr(), pg.plan().data_ptr(), pg.n_edges, K, 1, wpack.data_ptr(), M,
                                             ws.data_ptr(), sp))
    torch.cuda.synchronize()
lib.lgcn_debug_flags(0)
t = tl.cpu().numpy().astype(np.int64)
t0 = t[0, 3]
names = ["mma: weights ready", "mma: A ready", "mma: issued", "prod: stage start", "prod: chunk in regs", "prod: A slot free", "prod: published"]
print("first 24 stages, cycles since the producer's first stamp")
print("stage " + " ".join(f"{n[:18]:>18s}" for n in names))
for s in range(70, 94):
    print(f"{s:5d} " + " ".join(f"{t[s, c] - t0:18d}" for c in range(7)))
a, b = 200, 1000
print("\nsteady state (stages %d..%d): mean cycles per stage" % (a, b))
print(" mma loop period            ", np.diff(t[a:b, 2]).mean())
print(" mma: wait weights          ", (t[a:b, 0] - t[a - 1:b - 1, 2]).mean())
print(" mma: wait A                ", (t[a:b, 1] - t[a:b, 0]).mean())
print(" mma: issue                 ", (t[a:b, 2] - t[a:b, 1]).mean())
print(" prod loop period           ", np.diff(t[a:b, 6]).mean())
print(" prod: take (wait + lds)    ", (t[a:b, 4] - t[a:b, 3]).mean())
print(" prod: issue + wait A slot  ", (t[a:b, 5] - t[a:b, 4]).mean())
print(" prod: convert+st+publish   ", (t[a:b, 6] - t[a:b, 5]).mean())
print(" prod: rest (flush/epilogue)", (t[a + 1:b, 3] - t[a:b - 1, 6]).mean())
print(" A lead: publish(s) -> mma A ready(s)", (t[a:b, 1] - t[a:b, 6]).mean())

print("\nMMA warp: per-stage period (cycles) for two tiles, 16 stages per line (64 MMA stages per tile incl. ctr2)")
d = np.diff(t[:704, 2])
for s0 in range(64, 192, 16):
    print(f"{s0:4d}: " + " ".join(f"{int(x):5d}" for x in d[s0:s0 + 16]))
print("tile totals:", [int(t[64 * (i + 1), 2] - t[64 * i, 2]) for i in range(1, 9)])

nz = np.nonzero(t[:1000, 7])[0]
if len(nz) >= 8:
    k = nz[4]
    print("\nsecond epilogue of one tile (producer warp 4), cycles: start -> drained+normed -> residual added -> stored")
    print("  ", [int(t[k + i, 7] - t[k, 7]) for i in range(4)], " | last tile: acc_full seen -> accumulators released:", int(t[1017, 7] - t[1016, 7]),
          " finish start -> acc_full seen:", int(t[1016, 7] - t[nz[-4], 7]))

print("\nMMA thread, stages 76..92: issue (12 MMAs + commit) | wait for the next stage | total")
for s_ in range(76, 92):
    print(f"  {s_:4d} kc={s_ % 4}: {int(t[s_, 1] - t[s_, 0]):6d} {int(t[s_, 2] - t[s_, 1]):6d} {int(t[s_ + 1, 0] - t[s_, 0]):6d}")
