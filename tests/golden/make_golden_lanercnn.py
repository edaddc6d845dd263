"""Golden fixtures for the LaneRCNN lane-graph layers from the UNMODIFIED reference lanercnn.py (needs
/root/reference).  Writes tests/golden/lanercnn_shapes.json (parameter names/shapes of LaneRoI and GlobalGraphNet)
and lanercnn_tiny.npz (their outputs on the batched tiny_b3 graph, seeded weights and input; LaneRoI with the `left`
edge set emptied so the per-key `len > 0` guards of lanercnn.py:397-415 are exercised).

    python tests/golden/make_golden_lanercnn.py
"""
import copy
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_loader  # noqa: E402

ref_lanegcn, ref_data = ref_loader.load()
import lanercnn as ref_rcnn  # noqa: E402  (reference module, found through the path ref_loader set up)

from helpers import golden_scenes  # noqa: E402
from lanegcn_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEED_W, SEED_X = 5, 11


def main():
    cfg = dict(ref_rcnn.config)
    roi, ggn = ref_rcnn.LaneRoI(cfg, 128).eval(), ref_rcnn.GlobalGraphNet(cfg).eval()
    shapes = {"roi." + k: list(v.shape) for k, v in roi.state_dict().items()}
    shapes.update({"ggn." + k: list(v.shape) for k, v in ggn.state_dict().items()})
    sd = synth.seeded_state_dict(shapes, SEED_W)
    roi.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("roi.")})
    ggn.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("ggn.")})
    scenes = golden_scenes("tiny_b3")
    graph = ref_lanegcn.graph_gather(ref_lanegcn.to_long(ref_data.collate_fn(copy.deepcopy(scenes))["graph"]))
    graph_roi = dict(graph)
    graph_roi["left"] = {"u": graph["left"]["u"][:0], "v": graph["left"]["v"][:0]}
    n = graph["feats"].shape[0]
    feat = torch.from_numpy(np.random.default_rng(SEED_X).standard_normal((n, 128)).astype(np.float32))
    with torch.no_grad():
        y_roi, y_ggn = roi(feat.clone(), graph_roi), ggn(feat.clone(), graph)
    json.dump(shapes, open(os.path.join(HERE, "lanercnn_shapes.json"), "w"), indent=0, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "lanercnn_tiny.npz"), roi=y_roi.numpy(), ggn=y_ggn.numpy())
    print("roi", tuple(y_roi.shape), "ggn", tuple(y_ggn.shape))


if __name__ == "__main__":
    main()
