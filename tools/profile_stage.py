"""Host-side cost of Net.stage / forward_device launch on the GPU box."""
import cProfile, io, json, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanegcn_b200 import synth
from lanegcn_b200 import lanegcn as L
dev = torch.device("cuda", 0)
shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_shapes.json")))
net = L.Net(L.config); net.load_state_dict(synth.seeded_state_dict(shapes, 0)); net = net.to(dev).eval()
data = synth.collate(synth.make_scenes(128, "argo-1.5k"))
for _ in range(5):
    b = net.stage(data); net.forward_device(b)
torch.cuda.synchronize()
def wall(f, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return 1e3 * (t1 - t0) / n, 1e3 * (t2 - t0) / n, r
print("stage: host ms %.2f, host+device ms %.2f" % wall(lambda: net.stage(data))[:2])
b = net.stage(data)
print("forward_device: host ms %.2f, host+device ms %.2f" % wall(lambda: net.forward_device(b))[:2])
print("stage_graphs: host ms %.2f, host+device ms %.2f" % wall(lambda: L.stage_graphs(data["graph"]))[:2])
pr = cProfile.Profile(); pr.enable()
for _ in range(5): net.stage(data)
pr.disable(); s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
pr = cProfile.Profile(); pr.enable()
for _ in range(5): net.forward_device(b)
torch.cuda.synchronize()
pr.disable(); s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
