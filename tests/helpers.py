"""Shared test helpers: golden inputs, seeded weights, tolerances."""
import json
import os

import numpy as np
import torch

from lanegcn_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# north_star tolerance for float outputs: within 1e-4 relative / 1e-5 absolute of the reference fp32 forward
RTOL, ATOL = 1e-4, 1e-5
STAGES = ["actor_net", "map_net", "a2m", "m2m", "m2a", "a2a"]


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def shapes():
    return json.load(open(os.path.join(GOLDEN, "state_dict_shapes.json")))


def weights(seed=0):
    return synth.seeded_state_dict(shapes(), seed)


def golden_scenes(name):
    """Must mirror tests/golden/make_golden.py::golden_scenes (the seed is recorded in the fixture)."""
    seed0 = int(golden(name)["seed0"])
    if name == "tiny_b3":
        scenes = synth.make_scenes(3, "tiny", seed0=seed0)
        scenes[1]["ctrs"] = scenes[1]["ctrs"] + np.float32(200.0)
        return scenes
    if name == "argo_b1":
        return synth.make_scenes(1, "argo-1.5k", seed0=seed0)
    raise KeyError(name)


def assert_close(got, want, what="", rtol=RTOL, atol=ATOL):
    got = torch.as_tensor(got).detach().cpu().double()
    want = torch.as_tensor(want).detach().cpu().double()
    assert got.shape == want.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    bad = err > tol
    assert not bool(bad.any()), (
        f"{what}: {int(bad.sum())}/{bad.numel()} elements outside rtol={rtol} atol={atol}; "
        f"max abs err {err.max().item():.3e}, max err/tol {(err / tol).max().item():.2f}"
    )


def roi_scenes():
    """Must mirror tests/golden/make_golden_lanercnn.py::roi_scenes; the per-agent lane-RoI sub-graphs (built by the
    reference's data_lrcnn.generate_lane_roi, host numpy, out of scope) are read back from the fixture."""
    scenes = synth.make_scenes(2, "tiny", seed0=300)
    for s in scenes:
        g = s["graph"]
        rng = np.random.default_rng(int(s["idx"]) + 7)
        for a in range(len(s["ctrs"])):
            near = np.argmin(((g["ctrs"] - s["ctrs"][a]) ** 2).sum(1))
            d = g["feats"][near] / np.linalg.norm(g["feats"][near])
            s["ctrs"][a] = g["ctrs"][near] + rng.normal(0, 0.3, 2).astype(np.float32)
            s["feats"][a, :, :2] = (0.5 * d)[None] + rng.normal(0, 0.01, (20, 2))
    fx = golden("lanercnn_roi")
    for b, s in enumerate(scenes):
        sgs = []
        for a in range(int(fx[f"n_roi_{b}"])):
            p = f"sg_{b}_{a}_"
            sg = {"feats": fx[p + "feats"], "agent_feat": fx[p + "agent_feat"], "agent_vel": float(fx[p + "agent_vel"]),
                  "a2m": {"u": fx[p + "a2m_u"], "v": fx[p + "a2m_v"]}}
            for k1 in ("pre", "suc"):
                sg[k1] = [{"u": fx[p + f"{k1}{i}_u"], "v": fx[p + f"{k1}{i}_v"]} for i in range(6)]
            for k1 in ("left", "right"):
                sg[k1] = {"u": fx[p + f"{k1}_u"], "v": fx[p + f"{k1}_v"]}
            sgs.append(sg)
        s["subgraphs"] = sgs
    return scenes
