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


# --------------------------------------------------------------------------- LaneInput / LanePooling / Interactor
def roi_scenes():
    """tiny scenes whose actors move at ~5 m/s along their nearest lane (generate_lane_roi skips agents with zero
    velocity or a heading far from every nearby lane, data_lrcnn.py:757-779), with the `obs_trajs` key it reads."""
    scenes = synth.make_scenes(2, "tiny", seed0=300)
    for s in scenes:
        g = s["graph"]
        rng = np.random.default_rng(int(s["idx"]) + 7)
        for a in range(len(s["ctrs"])):
            near = np.argmin(((g["ctrs"] - s["ctrs"][a]) ** 2).sum(1))
            d = g["feats"][near] / np.linalg.norm(g["feats"][near])
            s["ctrs"][a] = g["ctrs"][near] + rng.normal(0, 0.3, 2).astype(np.float32)
            s["feats"][a, :, :2] = (0.5 * d)[None] + rng.normal(0, 0.01, (20, 2))
        t = np.arange(-19, 1, dtype=np.float32)[None, :, None]
        s["obs_trajs"] = (s["ctrs"][:, None, :] + t * s["feats"][:, -1:, :2]).astype(np.float32)
    return scenes


def dump_subgraphs(rec, scenes):
    rec["n_scenes"] = np.asarray(len(scenes))
    for b, s in enumerate(scenes):
        rec[f"n_roi_{b}"] = np.asarray(len(s["subgraphs"]))
        for a, sg in enumerate(s["subgraphs"]):
            p = f"sg_{b}_{a}_"
            rec[p + "feats"], rec[p + "agent_feat"] = sg["feats"], sg["agent_feat"].astype(np.float32)
            rec[p + "agent_vel"] = np.asarray(sg["agent_vel"], np.float32)
            rec[p + "a2m_u"], rec[p + "a2m_v"] = sg["a2m"]["u"], sg["a2m"]["v"]
            for k1 in ("pre", "suc"):
                for i in range(6):
                    rec[p + f"{k1}{i}_u"], rec[p + f"{k1}{i}_v"] = sg[k1][i]["u"], sg[k1][i]["v"]
            for k1 in ("left", "right"):
                rec[p + f"{k1}_u"], rec[p + f"{k1}_v"] = sg[k1]["u"], sg[k1]["v"]


def main_roi():
    import data_lrcnn as ref_dl  # noqa: E402 (reference module)

    ref_rcnn.gpu = lambda x: x
    cfg = dict(ref_rcnn.config)
    scenes = [ref_dl.generate_lane_roi(s) for s in roi_scenes()]
    assert all(len(s["subgraphs"]) >= 2 for s in scenes), [len(s["subgraphs"]) for s in scenes]
    rec = {}
    dump_subgraphs(rec, scenes)
    batch = ref_dl.collate_fn(copy.deepcopy(scenes))
    graph = ref_rcnn.graph_gather(ref_rcnn.to_long(batch["graph"]))
    roi = ref_rcnn.subgraph_gather(ref_rcnn.to_long(batch["subgraphs"]))
    mods = {"inp": ref_rcnn.LaneInput(cfg).eval(), "pool": ref_rcnn.LanePooling(128, 128).eval(),
            "inter": ref_rcnn.Interactor(cfg).eval()}
    shapes = {f"{n}.{k}": list(v.shape) for n, m in mods.items() for k, v in m.state_dict().items()}
    sd = synth.seeded_state_dict(shapes, SEED_W + 1)
    for n, m in mods.items():
        m.load_state_dict({k[len(n) + 1:]: v for k, v in sd.items() if k.startswith(n + ".")})
    rng = np.random.default_rng(SEED_X + 1)
    roi_feat = torch.from_numpy(rng.standard_normal((roi["num_nodes"], 128)).astype(np.float32))
    g_feat = torch.from_numpy(rng.standard_normal((len(graph["feats"]), 128)).astype(np.float32))
    with torch.no_grad():
        rec["out_input"] = mods["inp"](roi).numpy()
        rec["out_pool_r2g"] = mods["pool"](roi_feat.clone(), roi, g_feat.clone(), graph).numpy()
        rec["out_pool_g2r"] = mods["pool"](g_feat.clone(), graph, roi_feat.clone(), roi).numpy()
        rec["out_interactor"] = mods["inter"](graph, roi, roi_feat.clone()).numpy()
    rec["a2m_u"], rec["a2m_v"] = roi["a2m"]["u"].numpy(), roi["a2m"]["v"].numpy()
    rec["roi_pre3_v"], rec["roi_left_u"] = roi["pre"][3]["v"].numpy(), roi["left"]["u"].numpy()
    json.dump(shapes, open(os.path.join(HERE, "lanercnn_roi_shapes.json"), "w"), indent=0, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "lanercnn_roi.npz"), **rec)
    print("rois per scene", [len(s["subgraphs"]) for s in scenes], "roi nodes", roi["num_nodes"],
          {k: v.shape for k, v in rec.items() if k.startswith("out_")})


if __name__ == "__main__":
    main_roi()
