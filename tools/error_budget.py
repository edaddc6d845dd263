"""Error budget on the GPU box: (1) one GEMM, both engines, against fp64; (2) per-stage error of the full forward
for both engines against the fp32 oracle and the fp64 oracle (how much of the tolerance the reference's own fp32
rounding already uses)."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import golden_scenes, weights  # noqa: E402
from lanegcn_b200 import _C, synth  # noqa: E402
from lanegcn_b200 import lanegcn as L  # noqa: E402
from oracle import lanegcn_oracle as O  # noqa: E402

dev = torch.device("cuda", 0)
lib = _C.lib()
sp = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731

g = torch.Generator().manual_seed(0)
for scale in (1.0, 8.0):
    A = (torch.randn(4096, 128, generator=g) * scale)
    W = torch.randn(128, 128, generator=g) / 11.3
    ref = (A.double() @ W.double().t())
    cpu32 = (A @ W.t()).double()
    for eng in (0, 1):
        lib.lgcn_set_gemm_engine(eng)
        out = torch.empty(4096, 128, device=dev)
        dA, dW = A.to(dev), W.to(dev)
        _C.check(lib.lgcn_linear128(dA.data_ptr(), None, None, None, None, None, 1, None, 0, dW.data_ptr(), 1, None, None,
                                    None, 0, out.data_ptr(), 128, 4096, sp()))
        e = out.cpu().double() - ref
        print(f"scale {scale} engine {eng}: max abs {e.abs().max():.3e} rms {e.pow(2).mean().sqrt():.3e} "
              f"mean signed {e.mean():.3e} mean signed*sign(ref) {(e * ref.sign()).mean():.3e}   (ref rms {ref.pow(2).mean().sqrt():.3f})")
    e = cpu32 - ref
    print(f"scale {scale} cpu fp32: max abs {e.abs().max():.3e} rms {e.pow(2).mean().sqrt():.3e}")

for name in ("tiny_b3", "argo_b1"):
    sd = weights()
    sd64 = {k: v.double() for k, v in sd.items()}
    batch = synth.collate(golden_scenes(name))

    def to64(x):
        if isinstance(x, dict):
            return {k: to64(v) for k, v in x.items()}
        if isinstance(x, list):
            return [to64(v) for v in x]
        if torch.is_tensor(x) and x.dtype == torch.float32:
            return x.double()
        return x

    t32, t64 = {}, {}
    with torch.no_grad():
        o32 = O.net_forward(sd, batch, t32)
        o64 = O.net_forward(sd64, to64(synth.collate(golden_scenes(name))), t64)
    t32["cls"], t32["reg"] = torch.cat(o32["cls"]), torch.cat(o32["reg"])
    t64["cls"], t64["reg"] = torch.cat(o64["cls"]), torch.cat(o64["reg"])
    net = L.Net(L.config)
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    net.use_cuda_graphs = False   # forward hooks copy to the host: not capturable
    for eng, fused in ((0, False), (1, False), (1, True)):
        lib.lgcn_set_gemm_engine(eng)
        L.LANECONV_FUSED = fused
        taps = {}
        hooks = [getattr(net, s).register_forward_hook(
            lambda m, i, o, s=s: taps.__setitem__(s, (o[0] if isinstance(o, tuple) else o).detach().cpu()))
            for s in ("actor_net", "map_net", "a2m", "m2m", "m2a", "a2a")]
        out = net(synth.collate(golden_scenes(name)))
        for h in hooks:
            h.remove()
        taps["cls"], taps["reg"] = torch.cat(out["cls"]).cpu(), torch.cat(out["reg"]).cpu()
        print(f"--- {name} engine {eng} ({'aggregate-first LaneConv' if fused else 'split LaneConv path'})")
        for s in ("actor_net", "map_net", "a2m", "m2m", "m2a", "a2a", "cls", "reg"):
            ref32, ref64, got = t32[s].double(), t64[s], taps[s].double()
            tol = 1e-5 + 1e-4 * ref32.abs()
            e_ours, e_ref = (got - ref32).abs(), (ref32 - ref64).abs()
            print(f"{s:10s} ours-vs-fp32ref: max {e_ours.max():.2e} max err/tol {(e_ours / tol).max():.2f} viol {(e_ours > tol).sum().item()}"
                  f" | fp32ref-vs-fp64: max {e_ref.max():.2e} max err/tol {(e_ref / tol).max():.2f} | ours-vs-fp64 max {(got - ref64).abs().max():.2e}")
