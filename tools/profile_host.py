"""Host-side profile of Net.forward_device / Net.forward on the GPU box (cProfile + synchronised section timers)."""
import cProfile
import io
import json
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanegcn_b200 import synth  # noqa: E402
from lanegcn_b200 import lanegcn as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_shapes.json")))
net = L.Net(L.config)
net.load_state_dict(synth.seeded_state_dict(shapes, 0))
net = net.to(dev).eval()
data = synth.collate(synth.make_scenes(B, "argo-1.5k"))
staged = net.stage(data)
for _ in range(3):
    net.forward_device(staged)
torch.cuda.synchronize()


def timed(fn, n=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n, r


print("forward_device ms", timed(lambda: net.forward_device(staged))[0])
print("stage ms", timed(lambda: net.stage(data))[0])
print("forward (e2e, no D2H) ms", timed(lambda: net(data))[0])
ms, graph = timed(lambda: L.finish_graph(staged.graphs))
print("finish_graph ms", ms)
actor_ctrs = staged.actor_ctrs
print("pair lists ms", timed(lambda: L.build_pair_lists([(graph["ctrs"], actor_ctrs, 7.0), (actor_ctrs, graph["ctrs"], 6.0), (actor_ctrs, actor_ctrs, 100.0)]))[0])
ms, actors = timed(lambda: net.actor_net(staged.actors.transpose(1, 2).contiguous()))
print("actor_net ms", ms)
ms, (nodes, idcs, ctrs) = timed(lambda: net.map_net(graph))
print("map_net ms", ms)
sizes = [len(x) for x in actor_ctrs]
actor_idcs = L.scene_list(torch.arange(sum(sizes), device=dev), sizes, actor_ctrs.off_dev)
ms, n2 = timed(lambda: net.a2m(nodes, graph, actors, actor_idcs, actor_ctrs))
print("a2m ms", ms)
print("m2m ms", timed(lambda: net.m2m(n2, graph))[0])
ms, a2 = timed(lambda: net.m2a(actors, actor_idcs, actor_ctrs, n2, idcs, ctrs))
print("m2a ms", ms)
print("a2a ms", timed(lambda: net.a2a(a2, actor_idcs, actor_ctrs))[0])
print("pred_net ms", timed(lambda: net.pred_net(a2, actor_idcs, actor_ctrs))[0])

pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    net.forward_device(staged)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
