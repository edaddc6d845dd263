"""Ablation timing of the tcgen05 GEMM on the GPU box: wide projection (15 output blocks) and ctr2-style linear,
under the lgcn_debug_flags switches."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanegcn_b200 import _C
lib = _C.lib()
dev = torch.device("cuda", 0)
M = 193536
g = torch.Generator(device="cpu").manual_seed(0)
X = torch.randn(M, 128, generator=g).to(dev)
Ww = (torch.randn(1920, 128, generator=g) / 11).to(dev)
W1 = (torch.randn(128, 128, generator=g) / 11).to(dev)
gam, bet = torch.ones(128, device=dev), torch.zeros(128, device=dev)
Y = torch.empty(M, 1920, device=dev)
O = torch.empty(M, 128, device=dev)
R = torch.randn(M, 128, device=dev)
sp = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def run(kind):
    if kind == "wide":
        _C.check(lib.lgcn_linear128(X.data_ptr(), None, None, None, None, None, 1, None, 0, Ww.data_ptr(), 15, None, None, None, 0, Y.data_ptr(), 1920, M, sp))
    elif kind == "ctr2":
        _C.check(lib.lgcn_linear128(X.data_ptr(), None, None, None, None, None, 1, None, 0, W1.data_ptr(), 1, gam.data_ptr(), bet.data_ptr(), R.data_ptr(), 13, O.data_ptr(), 128, M, sp))
    else:
        _C.check(lib.lgcn_linear128(X.data_ptr(), None, None, None, None, None, 1, None, 0, W1.data_ptr(), 1, None, None, None, 0, O.data_ptr(), 128, M, sp))

def timeit(kind, n=5):
    for _ in range(2):
        run(kind)
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(kind); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)

for flags in (0, 1, 2, 3, 4, 8, 12, 13, 15):
    lib.lgcn_debug_flags(flags)
    print(f"dbg={flags:2d}  wide {timeit('wide')*1e3:8.1f} us   ctr2 {timeit('ctr2')*1e3:8.1f} us   plain128 {timeit('plain')*1e3:8.1f} us", flush=True)
lib.lgcn_debug_flags(0)
