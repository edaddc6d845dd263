"""Error of the aggregate-first LaneConv stack (tcgen05 3xTF32) against the fp32 SIMT split stack on a real lane graph,
for different accumulator-flush intervals (debug flags: 0 = every 3 keys, 512 = every 5 keys, 64 = never)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanegcn_b200 import _C, synth
from lanegcn_b200 import lanegcn as L
lib = _C.lib()
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pg = L.graph_gather(synth.collate(synth.make_scenes(B, "argo-1.5k"))["graph"])["_packed"]
M, K = pg.n_nodes, pg.n_keys
sp = torch.cuda.current_stream().cuda_stream
for seed in range(3):
    g = torch.Generator().manual_seed(seed)
    per = lib.lgcn_laneconv_wpack_floats(K)
    wp = torch.empty(4 * per)
    for i in range(4):
        o = i * per
        wp[o:o + (K + 2) * 16384] = torch.randn((K + 2) * 16384, generator=g) * (0.7 / 128 ** 0.5)
        wp[o + (K + 2) * 16384:o + per] = torch.randn(512, generator=g) * 0.3 + torch.tensor([1.0, 0.0, 1.0, 0.0]).repeat_interleave(128)
    wp = wp.to(dev)
    x = torch.randn(M, 128, generator=g).to(dev)
    lib.lgcn_set_gemm_engine(0)
    want = x.clone()
    ws = torch.empty(lib.lgcn_laneconv_workspace_bytes(M, K), dtype=torch.uint8, device=dev)
    _C.check(lib.lgcn_laneconv_stack(want.data_ptr(), pg.rowptr.data_ptr(), pg.col.data_ptr(), K, 4, wp.data_ptr(), M, ws.data_ptr(), sp))
    lib.lgcn_set_gemm_engine(1)
    split = x.clone()
    _C.check(lib.lgcn_laneconv_stack(split.data_ptr(), pg.rowptr.data_ptr(), pg.col.data_ptr(), K, 4, wp.data_ptr(), M, ws.data_ptr(), sp))
    ws2 = torch.empty(lib.lgcn_laneconv_planned_workspace_bytes(M, pg.n_edges, K), dtype=torch.uint8, device=dev)
    res = {"split tcgen05": split}
    for name, fl in [("fused, flush 3", 0), ("fused, flush 5", 512), ("fused, no flush", 64)]:
        lib.lgcn_debug_flags(fl)
        got = x.clone()
        _C.check(lib.lgcn_laneconv_stack_planned(got.data_ptr(), pg.plan().data_ptr(), pg.n_edges, K, 4, wp.data_ptr(), M, ws2.data_ptr(), sp))
        lib.lgcn_debug_flags(0)
        res[name] = got
    torch.cuda.synchronize()
    for name, got in res.items():
        err = (got.double() - want.double()).abs()
        tol = 1e-5 + 1e-4 * want.double().abs()
        print(f"seed {seed} {name:18s} max err/tol {float((err / tol).max()):6.3f}  rms err {float(err.pow(2).mean().sqrt()):.3e}  (out rms {float(want.pow(2).mean().sqrt()):.3f})")
