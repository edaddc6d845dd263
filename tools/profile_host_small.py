"""Host-side time of forward_device at a small shard (16 scenes = the per-GPU shard at 8 GPUs)."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanegcn_b200 import synth
from lanegcn_b200 import lanegcn as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_shapes.json")))
net = L.Net(L.config); net.load_state_dict(synth.seeded_state_dict(shapes, 0)); net = net.to(dev).eval()
data = synth.collate(synth.make_scenes(B, "argo-1.5k"))
b = net.stage(data)
for _ in range(5):
    net.forward_device(b)
torch.cuda.synchronize()
import lanegcn_b200.lanegcn as M
marks = []
def mark(name):
    marks.append((name, time.perf_counter()))
# monkeypatch a few functions to time them on the host (no sync)
def wrap(obj, name, label):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); acc[label] = acc.get(label, 0.0) + time.perf_counter() - t0; return r
    setattr(obj, name, g)
acc = {}
wrap(M, "finish_graph", "finish_graph"); wrap(M, "count_pair_lists", "pairs.count"); wrap(M, "fill_pair_lists", "pairs.fill(+sync)")
for name in ("map_net", "a2m", "m2m", "m2a", "a2a"):
    mod = getattr(net, name); f = mod.forward
    def g(*a, _f=f, _n=name, **k):
        t0 = time.perf_counter(); r = _f(*a, **k); acc[_n] = acc.get(_n, 0.0) + time.perf_counter() - t0; return r
    mod.forward = g
for gname in ("_g_actor", "_g_pred"):
    gobj = getattr(net, gname); f = gobj.__call__
    class W:
        def __init__(s, o, n): s.o, s.n = o, n
        def __call__(s, *a):
            t0 = time.perf_counter(); r = s.o(*a); acc[s.n] = acc.get(s.n, 0.0) + time.perf_counter() - t0; return r
    setattr(net, gname, W(gobj, gname))
n = 20
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(n):
    net.forward_device(b)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"B={B}: forward_device host {1e3*(t1-t0)/n:.3f} ms, host+device {1e3*(t2-t0)/n:.3f} ms per call")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print(f"  {k:20s} {1e3*v/n:.3f} ms")
