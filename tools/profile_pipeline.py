"""Timeline of the prefetching e2e loop (host timestamps) on the GPU box."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanegcn_b200 import synth
from lanegcn_b200 import lanegcn as L
dev = torch.device("cuda", 0)
shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_shapes.json")))
net = L.Net(L.config); net.load_state_dict(synth.seeded_state_dict(shapes, 0)); net = net.to(dev).eval()
data = synth.collate(synth.make_scenes(128, "argo-1.5k"))
for _ in range(4):
    out = net(data); torch.cat(out["reg"]).cpu()
torch.cuda.synchronize()
T = []
staged = net.stage(data)
for i in range(8):
    t0 = time.perf_counter()
    out = net.forward_device(staged)
    t1 = time.perf_counter()
    staged = net.stage(data)
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    c, r = L._cat_of(out["cls"]).cpu(), L._cat_of(out["reg"]).cpu()
    t4 = time.perf_counter()
    T.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
for row in T:
    print("forward_device host %.2f | stage %.2f | wait gpu %.2f | D2H %.2f | total %.2f ms" % tuple(1e3 * x for x in row + (sum(row),)))
# inside forward_device: where does the host time go?
import cProfile, io, pstats
pr = cProfile.Profile()
for i in range(5):
    torch.cuda.synchronize()
    pr.enable(); out = net.forward_device(staged); pr.disable()
torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(12); print(s.getvalue()[:3000])

def finish(out):
    return L._cat_of(out["cls"]).cpu(), L._cat_of(out["reg"]).cpu()

def run_e2e(n):
    for out in L.prefetch_forward(net, (data for _ in range(n))):
        res = finish(out)
    return res

run_e2e(3)
torch.cuda.synchronize()
t0 = time.perf_counter(); run_e2e(10); torch.cuda.synchronize()
print("prefetch_forward generator: %.2f ms/step" % (1e3 * (time.perf_counter() - t0) / 10))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
run_e2e(3)
torch.cuda.synchronize()
t0 = time.perf_counter(); run_e2e(10); torch.cuda.synchronize()
print("  ... with a 256 MiB buffer allocated: %.2f ms/step" % (1e3 * (time.perf_counter() - t0) / 10))
