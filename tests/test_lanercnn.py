"""LaneRCNN lane-graph layers (BASELINE config 5): parameter contract + oracle vs the reference's golden on CPU,
drop-in modules vs the golden on the GPU."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, assert_close, golden, golden_scenes
from lanegcn_b200 import lanercnn as R
from lanegcn_b200 import lanegcn as L
from lanegcn_b200 import synth
from oracle import lanegcn_oracle as O

SEED_W, SEED_X = 5, 11


def _shapes():
    return json.load(open(os.path.join(GOLDEN, "lanercnn_shapes.json")))


def _inputs():
    batch = synth.collate(golden_scenes("tiny_b3"))
    graph = O.graph_gather(O.to_long(batch["graph"]))
    n = graph["feats"].shape[0]
    feat = torch.from_numpy(np.random.default_rng(SEED_X).standard_normal((n, 128)).astype(np.float32))
    return batch, graph, feat


def test_parameter_names_match_reference():
    want = _shapes()
    got = {"roi." + k: list(v.shape) for k, v in R.LaneRoI(L.config, 128).state_dict().items()}
    got.update({"ggn." + k: list(v.shape) for k, v in R.GlobalGraphNet(L.config).state_dict().items()})
    assert got == want


def test_oracle_matches_reference_golden():
    g = golden("lanercnn_tiny")
    sd = synth.seeded_state_dict(_shapes(), SEED_W)
    _, graph, feat = _inputs()
    graph_roi = dict(graph)
    graph_roi["left"] = {"u": graph["left"]["u"][:0], "v": graph["left"]["v"][:0]}
    with torch.no_grad():
        assert_close(O.lane_roi(sd, "roi", feat.clone(), graph_roi), g["roi"], "LaneRoI", rtol=1e-6, atol=1e-6)
        assert_close(O.global_graph_net(sd, "ggn", feat.clone(), graph), g["ggn"], "GlobalGraphNet", rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
def test_dropins_match_reference_golden(cuda, lib):
    g = golden("lanercnn_tiny")
    sd = synth.seeded_state_dict(_shapes(), SEED_W)
    roi, ggn = R.LaneRoI(L.config, 128), R.GlobalGraphNet(L.config)
    roi.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("roi.")})
    ggn.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("ggn.")})
    roi, ggn = roi.to(cuda).eval(), ggn.to(cuda).eval()
    batch, _, feat = _inputs()
    graph = L.graph_gather(batch["graph"])
    # reference-style dict with the `left` set emptied (no _packed: the CSR is rebuilt from the dict)
    graph_roi = {k: v for k, v in graph.items() if k != "_packed"}
    graph_roi["left"] = {"u": graph["left"]["u"][:0], "v": graph["left"]["v"][:0]}
    assert_close(roi(feat.to(cuda), graph_roi), g["roi"], "LaneRoI")
    assert_close(ggn(feat.to(cuda), graph), g["ggn"], "GlobalGraphNet")
