"""Stage timeline of k_laneconv_v2, CTA 0 (build with LGCN_NVCC_EXTRA=-DLGCN_TIMELINE2): MMA thread and convert warp 8."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanegcn_b200 import _C, synth
from lanegcn_b200 import lanegcn as L
lib = _C.lib()
dev = torch.device("cuda", 0)
data = synth.collate(synth.make_scenes(128, "argo-1.5k"))
pg = L.graph_gather(data["graph"])["_packed"]
M, K = pg.n_nodes, pg.n_keys
g = torch.Generator().manual_seed(0)
wpack = (torch.randn(lib.lgcn_laneconv_wpack_floats(K), generator=g) / 11).to(dev)
feat = torch.randn(M, 128, generator=g).to(dev)
ws = torch.empty(lib.lgcn_laneconv_planned_workspace_bytes(M, pg.n_edges, K), dtype=torch.uint8, device=dev)
tl = torch.zeros(1024, 8, dtype=torch.int64, device=dev)
lib.lgcn_debug_timeline(tl.data_ptr())
sp = torch.cuda.current_stream().cuda_stream
for extra in [int(a) for a in sys.argv[1:]] or [0]:
    for it in range(2):
        tl.zero_()
        lib.lgcn_debug_flags(256 | extra)
        _C.check(lib.lgcn_laneconv_stack_planned(feat.data_ptr(), pg.plan().data_ptr(), pg.n_edges, K, 1, wpack.data_ptr(), M,
                                                 ws.data_ptr(), sp))
        torch.cuda.synchronize()
    lib.lgcn_debug_flags(0)
    t = tl.cpu().numpy().view(np.uint32).reshape(2, 1024, 8).astype(np.int64)
    mma, cv = t[0], t[1]
    n = int((mma[:, 0] != 0).sum())
    print(f"\n=== flags {extra}: MMA stages recorded {n}")
    per = (mma[1:n, 0] - mma[:n - 1, 0]) % (1 << 32)
    print("MMA thread: stage period, 16 per line, tiles 1-2 (64 stages per tile: 60 projection + 4 ctr2)")
    for s0 in range(64, 192, 16):
        print(f"{s0:4d}: " + " ".join(f"{int(x):5d}" for x in per[s0:s0 + 16]))
    print("tile totals:", [int((mma[64 * (i + 1), 0] - mma[64 * i, 0]) % (1 << 32)) for i in range(1, min(9, n // 64 - 1))])
    sid = mma[:n - 1, 3]
    for kc in range(4):
        m = (sid % 4 == kc) & (sid >= 4) & (sid < 56)
        m[:64] = False
        print(f"  steady kc={kc}: period {per[m].mean():7.0f}")
    m = (sid >= 4) & (sid < 56)
    m[:64] = False
    print(f"  steady all : period {per[m].mean():7.0f}   (768 = 12 MMAs x 64 cycles)")
    nc = int((cv[:, 0] != 0).sum())
    names = ["x-full wait+lds", "a-empty wait", "convert+st", "publish", "rest"]
    d = np.zeros((nc - 1, 5))
    for j in range(4):
        d[:, j] = (cv[:nc - 1, j + 1] - cv[:nc - 1, j]) % (1 << 32)
    d[:, 4] = (cv[1:nc, 0] - cv[:nc - 1, 4]) % (1 << 32)
    pc = (cv[1:nc, 0] - cv[:nc - 1, 0]) % (1 << 32)
    cs = cv[:nc - 1, 5]
    print("convert warp 8 (per projection stage): " + " | ".join(names) + " | period")
    for kc in range(4):
        m = (cs % 4 == kc) & (cs >= 4) & (cs < 56)
        m[:60] = False
        print(f"  steady kc={kc}: " + " ".join(f"{d[m, j].mean():8.0f}" for j in range(5)) + f" {pc[m].mean():8.0f}")
    for s0 in (56, 57, 58, 59, 0, 1, 2, 3, 4):
        m = cs == s0
        m[:60] = False
        print(f"  stage {s0:2d}:   " + " ".join(f"{d[m, j].mean():8.0f}" for j in range(5)) + f" {pc[m].mean():8.0f}")
