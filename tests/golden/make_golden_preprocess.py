"""Golden vectors for the left / right edge builder: outputs of the UNMODIFIED reference ``preprocess()``
(preprocess_data.py:287-392) run on CPU on seeded synthetic scenes.  Run in the authoring container
(needs /root/reference):  python tests/golden/make_golden_preprocess.py
Writes tests/golden/preprocess_lr.npz: per scene the reference's left/right u, v (int16)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from lanegcn_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

SCENES = [("tiny", 0, 0), ("tiny", 5, 1), ("small", 11, 0), ("small", 12, 1), ("argo-1.5k", 3, 0), ("argo-1.5k", 4, 1)]


def scene_graph(preset, seed, hard=0):
    """Imported by tests/test_gpu_preprocess.py and tests/test_oracle.py.  hard=1: lanes stretched to ~5.6 m spacing with
    0.4 m noise (the 6 m distance filter bites) and headings rotated by up to +-60 degrees (the 45-degree filter bites)."""
    g = dict(synth.make_scene(seed, preset)["graph"])
    rng = np.random.default_rng(1000 + seed)
    if hard:
        g["ctrs"] = (g["ctrs"] * np.float32(1.6) + rng.normal(0, 0.4, g["ctrs"].shape)).astype(np.float32)
        ang = rng.uniform(-np.pi / 3, np.pi / 3, len(g["feats"]))
        c, s = np.cos(ang), np.sin(ang)
        f = g["feats"]
        g["feats"] = np.stack([c * f[:, 0] - s * f[:, 1], s * f[:, 0] + c * f[:, 1]], 1).astype(np.float32)
    else:
        g["ctrs"] = (g["ctrs"] + rng.normal(0, 0.15, g["ctrs"].shape)).astype(np.float32)   # break the exact lane symmetry
    return g


def main():
    ref_loader.load()                      # shim + sys.path for the reference's bare-name imports
    sys.argv = sys.argv[:1]
    import preprocess_data as ref_pp       # the reference module (argparse parser is only built, not parsed)

    out = {}
    for k, (preset, seed, hard) in enumerate(SCENES):
        g = scene_graph(preset, seed, hard)
        tg = {key: torch.from_numpy(np.asarray(g[key]).astype(np.int64 if "pairs" in key or key == "lane_idcs" else np.float32))
              for key in ("ctrs", "feats", "lane_idcs", "pre_pairs", "suc_pairs", "left_pairs", "right_pairs")}
        tg["idx"] = k
        res = ref_pp.preprocess(tg, 6)
        for side in ("left", "right"):
            out[f"{k}_{side}_u"] = res[side]["u"]
            out[f"{k}_{side}_v"] = res[side]["v"]
        print(preset, seed, hard, g["num_nodes"], {s: len(res[s]["u"]) for s in ("left", "right")})
    np.savez_compressed(os.path.join(HERE, "preprocess_lr.npz"), **out)


if __name__ == "__main__":
    main()
