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

import ctypes
rng = torch.Generator(device="cpu").manual_seed(1)
K = 14
wpack = (torch.randn(lib.lgcn_laneconv_wpack_floats(K), generator=rng) / 11).to(dev)
E = 12 * M
rowptr = torch.arange(0, M + 1, dtype=torch.int32, device=dev) * 12
col = (torch.randint(0, M, (E,), generator=rng, dtype=torch.int32) * 15 + torch.randint(1, 15, (E,), generator=rng, dtype=torch.int32)).to(dev)
ws = torch.empty(lib.lgcn_laneconv_workspace_bytes(M, K), dtype=torch.uint8, device=dev)
feat = X.clone()


def stack_times(n=5):
    """(wide, gather, ctr2) us per launch inside lgcn_laneconv_stack (1 block), via the library's event profiler"""
    best = None
    for it in range(n + 1):
        feat.copy_(X)
        flush.zero_()
        lib.lgcn_prof_enable(1)
        _C.check(lib.lgcn_laneconv_stack(feat.data_ptr(), rowptr.data_ptr(), col.data_ptr(), K, 1, wpack.data_ptr(), M, ws.data_ptr(), sp))
        lib.lgcn_prof_enable(0)
        ms, cnt = (ctypes.c_double * 8)(), (ctypes.c_int64 * 8)()
        _C.check(lib.lgcn_prof_collect(ms, cnt))
        cur = [ms[0] * 1e3, ms[1] * 1e3, ms[2] * 1e3]
        if it:
            best = cur if best is None else [min(a, b) for a, b in zip(best, cur)]
    return best


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

for flags in (0, 1, 4, 8, 12, 13, 15, 16):
    lib.lgcn_debug_flags(flags)
    st = stack_times()
    print(f"dbg={flags:2d}  stack: wide {st[0]:7.1f} gather {st[1]:6.1f} ctr2 {st[2]:6.1f} us | direct: wide15 {timeit('wide')*1e3:7.1f} ctr2 {timeit('ctr2')*1e3:6.1f} plain128 {timeit('plain')*1e3:6.1f} us", flush=True)
lib.lgcn_debug_flags(0)
