"""LaneRCNN lane-graph layers (BASELINE config 5): parameter contract + oracle vs the reference's golden on CPU,
drop-in modules vs the golden on the GPU."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, assert_close, golden, golden_scenes, roi_scenes
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


# --------------------------------------------------------------------------- LaneInput / LanePooling / Interactor
def _roi_shapes():
    return json.load(open(os.path.join(GOLDEN, "lanercnn_roi_shapes.json")))


def _roi_inputs():
    batch = synth.collate(roi_scenes())
    fx = golden("lanercnn_roi")
    rng = np.random.default_rng(SEED_X + 1)
    n_roi = sum(len(sg["feats"]) for sgs in batch["subgraphs"] for sg in sgs)
    n_g = sum(int(g["num_nodes"]) for g in batch["graph"])
    roi_feat = torch.from_numpy(rng.standard_normal((n_roi, 128)).astype(np.float32))
    g_feat = torch.from_numpy(rng.standard_normal((n_g, 128)).astype(np.float32))
    return batch, fx, roi_feat, g_feat


def test_roi_module_parameter_names_match_reference():
    want = _roi_shapes()
    mods = {"inp": R.LaneInput(L.config), "pool": R.LanePooling(128, 128), "inter": R.Interactor(L.config)}
    got = {f"{n}.{k}": list(v.shape) for n, m in mods.items() for k, v in m.state_dict().items()}
    assert got == want


def test_roi_oracle_matches_reference_golden():
    batch, fx, roi_feat, g_feat = _roi_inputs()
    sd = synth.seeded_state_dict(_roi_shapes(), SEED_W + 1)
    graph = O.rcnn_graph_gather(O.to_long(batch["graph"]))
    roi = O.rcnn_subgraph_gather(O.to_long(batch["subgraphs"]))
    assert np.array_equal(roi["a2m"]["u"].numpy(), fx["a2m_u"]) and np.array_equal(roi["a2m"]["v"].numpy(), fx["a2m_v"])
    assert np.array_equal(roi["pre"][3]["v"].numpy(), fx["roi_pre3_v"]) and np.array_equal(roi["left"]["u"].numpy(), fx["roi_left_u"])
    with torch.no_grad():
        assert_close(O.lane_input(sd, "inp", roi), fx["out_input"], "LaneInput", rtol=1e-6, atol=1e-6)
        assert_close(O.lane_pooling(sd, "pool", roi_feat, roi, g_feat, graph), fx["out_pool_r2g"], "pool r2g", rtol=1e-6, atol=2e-6)
        assert_close(O.lane_pooling(sd, "pool", g_feat, graph, roi_feat, roi), fx["out_pool_g2r"], "pool g2r", rtol=1e-6, atol=2e-6)
        assert_close(O.interactor(sd, "inter", graph, roi, roi_feat), fx["out_interactor"], "Interactor", rtol=1e-6, atol=2e-6)


@pytest.mark.gpu
def test_roi_dropins_match_reference_golden(cuda, lib):
    batch, fx, roi_feat, g_feat = _roi_inputs()
    sd = synth.seeded_state_dict(_roi_shapes(), SEED_W + 1)
    mods = {"inp": R.LaneInput(L.config), "pool": R.LanePooling(128, 128), "inter": R.Interactor(L.config)}
    for n, m in mods.items():
        m.load_state_dict({k[len(n) + 1:]: v for k, v in sd.items() if k.startswith(n + ".")})
        m.to(cuda).eval()
    graph = R.graph_gather(batch["graph"])
    roi = R.subgraph_gather(batch["subgraphs"], cuda)
    assert np.array_equal(roi["a2m"]["v"].cpu().numpy(), fx["a2m_v"]) and np.array_equal(roi["pre"][3]["v"].cpu().numpy(), fx["roi_pre3_v"])
    assert_close(mods["inp"](roi), fx["out_input"], "LaneInput")
    assert_close(mods["pool"](roi_feat.to(cuda), roi, g_feat.to(cuda), graph), fx["out_pool_r2g"], "LanePooling roi->graph")
    assert_close(mods["pool"](g_feat.to(cuda), graph, roi_feat.to(cuda), roi), fx["out_pool_g2r"], "LanePooling graph->roi")
    assert_close(mods["inter"](graph, roi, roi_feat.to(cuda)), fx["out_interactor"], "Interactor")


@pytest.mark.gpu
def test_lanercnn_graph_path_end_to_end_vs_oracle(cuda, lib):
    """BASELINE config 5: LaneRCNN's forward graph layers chained (input -> roi_net1 -> interactor -> roi_net2) on
    synthetic scenes, against the oracle composition of the same reference call sites (lanercnn.py:97-112)."""
    net = R.Net(L.config)
    shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    sd = synth.seeded_state_dict(shapes, 9)
    net.load_state_dict(sd)
    net = net.to(cuda).eval()
    batch = synth.collate(roi_scenes())
    graph = O.rcnn_graph_gather(O.to_long(batch["graph"]))
    roi = O.rcnn_subgraph_gather(O.to_long(batch["subgraphs"]))
    with torch.no_grad():
        f = O.lane_input(sd, "input", roi)
        f = O.lane_roi(sd, "roi_net1", f, roi)
        f = O.interactor(sd, "interactor", graph, roi, f)
        want = O.lane_roi(sd, "roi_net2", f, roi)
    got = net(synth.collate(roi_scenes()))["roi_feat"]
    assert_close(got, want, "LaneRCNN roi_feat")
