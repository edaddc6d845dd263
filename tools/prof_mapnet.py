"""Small driver for ncu: MapNet forward (input MLPs + 4 LaneConv blocks) on a B-scene batch."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanegcn_b200 import synth
from lanegcn_b200 import lanegcn as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_shapes.json")))
net = L.Net(L.config); net.load_state_dict(synth.seeded_state_dict(shapes, 0)); net = net.to(dev).eval()
data = synth.collate(synth.make_scenes(B, "argo-1.5k"))
graph = L.graph_gather(data["graph"])
for _ in range(2):
    net.map_net(graph)
torch.cuda.synchronize()
print("ok")
