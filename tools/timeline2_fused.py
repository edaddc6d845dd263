"""Fine-grained producer timeline (build with LGCN_NVCC_EXTRA=-DLGCN_TIMELINE2): per-stage clock stamps of producer
warp 4 of CTA 0.  Prints mean deltas between consecutive stamps, by K-chunk index kc, over the steady-state stages of
the first tiles.  argv: debug flag combinations to run (256 is added)."""
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
names = ["top->peek", "peek->cpwait", "cpwait->lds", "lds->slot free", "slot->put16", "put16->issue", "issue->wait::st",
         "wait::st->arrive", "arrive->extras", "extras->flush", "flush->next top"]
for extra in [int(a) for a in sys.argv[1:]] or [0]:
    for it in range(2):
        tl.zero_()
        lib.lgcn_debug_flags(256 | extra)
        _C.check(lib.lgcn_laneconv_stack_planned(feat.data_ptr(), pg.plan().data_ptr(), pg.n_edges, K, 1, wpack.data_ptr(), M,
                                                 ws.data_ptr(), sp))
        torch.cuda.synchronize()
    lib.lgcn_debug_flags(0)
    t = tl.cpu().numpy().view(np.uint32).reshape(1024, 16).astype(np.int64)
    n = int((t[:, 0] != 0).sum())
    print(f"\n=== flags {extra}: {n} stages recorded")
    idx = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10]
    rows = np.arange(64, min(n, 600) - 1)
    sid = t[rows, 11]
    d = np.zeros((len(rows), 11))
    for j in range(10):
        d[:, j] = (t[rows, idx[j + 1]] - t[rows, idx[j]]) % (1 << 32)
    d[:, 10] = (t[rows + 1, 0] - t[rows, 10]) % (1 << 32)
    period = (t[rows + 1, 0] - t[rows, 0]) % (1 << 32)
    print("kc    " + " ".join(f"{x[:14]:>14s}" for x in names) + "         period")
    for kc in range(4):
        m = (sid % 4 == kc) & (sid >= 4) & (sid < 56)   # steady keys 1..13
        print(f"{kc}     " + " ".join(f"{d[m, j].mean():14.0f}" for j in range(11)) + f" {period[m].mean():14.0f}")
    m = (sid >= 4) & (sid < 56)
    print("all   " + " ".join(f"{d[m, j].mean():14.0f}" for j in range(11)) + f" {period[m].mean():14.0f}")
    print("median" + " ".join(f"{np.median(d[m, j]):14.0f}" for j in range(11)) + f" {np.median(period[m]):14.0f}")
    # tile-boundary stages
    for s0 in (56, 57, 58, 59, 0, 1, 2, 3):
        m = sid == s0
        print(f"s{s0:<4d} " + " ".join(f"{d[m, j].mean():14.0f}" for j in range(11)) + f" {period[m].mean():14.0f}")
