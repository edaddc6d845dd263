"""GPU (B200): graph construction before the path: scale-0 edges, full multi-scale edge build, and the
left / right edge builder — bit-exact against the reference's outputs (goldens) and the numpy oracle."""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import golden
from lanegcn_b200 import preprocess as PP
from lanegcn_b200 import synth
from oracle import graph_oracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_preprocess import SCENES, scene_graph  # noqa: E402

pytestmark = pytest.mark.gpu


def test_left_right_bit_exact_vs_reference_golden(cuda, lib):
    fx = golden("preprocess_lr")
    for k, (preset, seed, hard) in enumerate(SCENES):
        g = scene_graph(preset, seed, hard)
        g["idx"] = k
        out = PP.preprocess(g, 6)
        assert out["idx"] == k
        for side in ("left", "right"):
            for key in ("u", "v"):
                got = out[side][key]
                assert got.dtype == np.int16, "the reference stores int16 (preprocess_data.py:348-352)"
                assert np.array_equal(got, fx[f"{k}_{side}_{key}"]), (preset, seed, hard, side, key)


def test_left_right_random_graphs_vs_oracle(cuda, lib):
    """Random lane topologies incl. lanes without a side neighbour, empty pair sets, self-reachable lanes."""
    rng = np.random.default_rng(7)
    for trial in range(6):
        n_lanes = int(rng.integers(3, 40))
        lens = rng.integers(1, 12, n_lanes)
        lane_idcs = np.repeat(np.arange(n_lanes), lens)
        n = len(lane_idcs)
        g = {"ctrs": rng.uniform(-15, 15, (n, 2)).astype(np.float32), "feats": rng.normal(0, 1, (n, 2)).astype(np.float32),
             "lane_idcs": lane_idcs}
        for key, m in (("pre_pairs", n_lanes), ("suc_pairs", n_lanes), ("left_pairs", n_lanes // 2), ("right_pairs", 0 if trial == 0 else n_lanes // 2)):
            p = rng.integers(0, n_lanes, (m, 2))
            g[key] = p[np.argsort(p[:, 0], kind="stable")].astype(np.int64)
        out = PP.preprocess(g, 6)
        for side in ("left", "right"):
            u, v = graph_oracle.side_edges(g["ctrs"], g["feats"], g["lane_idcs"], g[side + "_pairs"], g["pre_pairs"],
                                           g["suc_pairs"], 6)
            assert np.array_equal(out[side]["u"].astype(np.int64), u) and np.array_equal(out[side]["v"].astype(np.int64), v)


@pytest.mark.parametrize("preset,seed", [("tiny", 2), ("small", 9), ("argo-1.5k", 1)])
def test_scale0_and_full_edge_build_bit_exact(cuda, lib, preset, seed):
    """scale 0 from the lane topology (data.py:272-295) + dilation (data.py:520-534) == the lists the generator made with
    the reference's scipy path, all six scales, both directions."""
    g = synth.make_scene(seed, preset)["graph"]
    pre, suc = PP.build_edges(g["lane_idcs"], g["pre_pairs"], g["suc_pairs"], 6)
    assert len(pre) == 6 and len(suc) == 6
    for got, want in ((pre, g["pre"]), (suc, g["suc"])):
        for s in range(6):
            for k in ("u", "v"):
                assert got[s][k].dtype == torch.int64
                assert np.array_equal(got[s][k].cpu().numpy(), want[s][k].astype(np.int64)), (s, k)


def test_scale0_rejects_unsorted_pairs(cuda, lib):
    g = synth.make_scene(0, "tiny")["graph"]
    bad = g["pre_pairs"][::-1].copy()
    if len(bad) > 1 and bad[0, 0] != bad[-1, 0]:
        with pytest.raises(RuntimeError, match="sorted"):
            PP.scale0_edges(g["lane_idcs"], bad, g["suc_pairs"])
