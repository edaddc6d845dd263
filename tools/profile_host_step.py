"""cProfile of the host side of one end-to-end step (Net.stage from packed scenes + Net.forward_device + D2H hand-off)
at a given batch size: where the per-step host time goes once the device work is one graph replay."""
import cProfile, json, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanegcn_b200 import synth
from lanegcn_b200 import lanegcn as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_shapes.json")))
net = L.Net(L.config); net.load_state_dict(synth.seeded_state_dict(shapes, 0)); net = net.to(dev).eval()
data = L.pack_batch(synth.collate(synth.make_scenes(B, "argo-1.5k")))
for _ in L.prefetch_forward(net, (data for _ in range(10)), to_host=True):
    pass
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 200
for _ in L.prefetch_forward(net, (data for _ in range(n)), to_host=True):
    pass
torch.cuda.synchronize()
print(f"B={B}: {1e3 * (time.perf_counter() - t0) / n:.3f} ms per step end to end")
t0 = time.perf_counter()
for _ in range(n):
    net.stage(data)
print(f"stage alone: {1e3 * (time.perf_counter() - t0) / n:.3f} ms")
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in L.prefetch_forward(net, (data for _ in range(n)), to_host=True):
    pass
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(28)
