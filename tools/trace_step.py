"""GPU timeline of one forward_device step (torch.profiler / CUPTI): kernel list with start offsets, durations and
the idle gaps between consecutive kernels per stream."""
import json, os, sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanegcn_b200 import synth
from lanegcn_b200 import lanegcn as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_shapes.json")))
net = L.Net(L.config); net.load_state_dict(synth.seeded_state_dict(shapes, 0)); net = net.to(dev).eval()
data = synth.collate(synth.make_scenes(B, "argo-1.5k"))
if len(sys.argv) > 3:
    from lanegcn_b200 import _C
    _C.lib().lgcn_debug_flags(int(sys.argv[3]))
staged = net.stage(data)
for _ in range(5):
    net.forward_device(staged)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    net.forward_device(staged)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
end = max(e.time_range.end for e in ev)
print(f"step span {1e-3 * (end - t0):.3f} ms, {len(ev)} device activities")
busy = 0.0
cur_end = t0
gaps = []
for e in ev:
    s, t = e.time_range.start, e.time_range.end
    if s > cur_end:
        gaps.append((s - cur_end, e.name[:60], 1e-3 * (s - t0)))
    if t > cur_end:
        busy += t - max(s, cur_end)
        cur_end = t
print(f"device busy (union over streams) {1e-3 * busy:.3f} ms, idle {1e-3 * (end - t0 - busy):.3f} ms in {len(gaps)} gaps")
gaps.sort(reverse=True)
print("largest gaps (us, before kernel, at ms):")
for g in gaps[:15]:
    print(f"  {g[0]:8.1f}  {g[1]:60s} @ {g[2]:.3f}")
agg = {}
for e in ev:
    a = agg.setdefault(e.name[:70], [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
print("top kernels by total time:")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"  {1e-3 * v[1]:7.3f} ms  x{v[0]:3d}  {k}")

if len(sys.argv) > 2:   # full timeline: start offset, duration, stream, name
    with open(sys.argv[2], "w") as f:
        for e in ev:
            f.write(f"{1e-3 * (e.time_range.start - t0):9.3f} {e.time_range.end - e.time_range.start:8.1f}us s{getattr(e, 'device_resource_id', '?')} {e.name[:90]}\n")
