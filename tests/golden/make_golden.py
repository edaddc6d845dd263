"""Golden fixtures from the UNMODIFIED reference (leepaul009/LaneGCN-1), run on CPU fp32 in the authoring
container.  Needs /root/reference; the GPU box and the driver's CPU test run only read the committed files.

    python tests/golden/make_golden.py

Writes (tests/golden/):
  state_dict_shapes.json   name -> shape of the reference Net's 405 state_dict entries (App. A.6 contract)
  tiny_b3.npz             batch of 3 "tiny" scenes (N=180 each; the middle scene's actors moved 200 m away so
                           A2M/M2A see an EMPTY scene -> exercises the offset quirk):
                           per-stage outputs (forward hooks), cls/reg, graph_gather outputs, every pair list
  argo_b1.npz              config 1 (one argo-1.5k scene): cls/reg + per-stage checksums + pair lists
  dilate_tiny.npz         reference data.dilated_nbrs output for scene 0's pre/suc scale-0 edges
Inputs and weights are NOT stored: both sides regenerate them from seeds (synth.make_scenes / seeded_state_dict;
the scene seed is recorded in each file as `seed0`).  Fixtures are required to be WELL CONDITIONED: the reference's
own fp32-vs-fp64 rounding noise must stay under 0.5 of the 1e-4/1e-5 tolerance on every stage, otherwise the
next seed is taken.
"""
import copy
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from lanegcn_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
STAGES = ["actor_net", "map_net", "a2m", "m2m", "m2a", "a2a"]


SEEDS = {"tiny_b3": 100, "argo_b1": 0}  # first seed0 >= this whose fixture is well conditioned (see main)


def golden_scenes(name, seed0=None):
    seed0 = SEEDS[name] if seed0 is None else seed0
    if name == "tiny_b3":
        scenes = synth.make_scenes(3, "tiny", seed0=seed0)
        scenes[1]["ctrs"] = scenes[1]["ctrs"] + np.float32(200.0)  # no node within 7 m / 6 m of any actor
        return scenes
    if name == "argo_b1":
        return synth.make_scenes(1, "argo-1.5k", seed0=seed0)
    raise KeyError(name)


def run_reference(net, ref_lanegcn, ref_data, scenes):
    batch = ref_data.collate_fn(copy.deepcopy(scenes))
    taps, pairs = {}, []
    hooks = []
    for s in STAGES:
        hooks.append(getattr(net, s).register_forward_hook(
            lambda m, i, o, s=s: taps.__setitem__(s, (o[0] if isinstance(o, tuple) else o).detach().clone())))
    orig_index_add = torch.Tensor.index_add_

    def spy(self, dim, index, src, *a, **k):
        pairs.append(index.detach().clone())
        return orig_index_add(self, dim, index, src, *a, **k)

    torch.Tensor.index_add_ = spy
    try:
        with torch.no_grad():
            out = net(batch)
    finally:
        torch.Tensor.index_add_ = orig_index_add
        for h in hooks:
            h.remove()
    graph = ref_lanegcn.graph_gather(ref_lanegcn.to_long(ref_data.collate_fn(copy.deepcopy(scenes))["graph"]))
    return out, taps, pairs, graph


def main():
    ref_lanegcn, ref_data = ref_loader.load()
    torch.manual_seed(0)
    net = ref_lanegcn.Net(ref_lanegcn.config).eval()
    shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    json.dump(shapes, open(os.path.join(HERE, "state_dict_shapes.json"), "w"), indent=0, sort_keys=True)
    net.load_state_dict(synth.seeded_state_dict(shapes, seed=0))

    from oracle import lanegcn_oracle as O

    def conditioning(scenes):
        """max over outputs of |fp32 - fp64| / (1e-5 + 1e-4 |fp32|) for the oracle restatement (bit-identical to
        the reference in fp32).  A fixture where the reference's OWN rounding noise already uses a large part of
        the tolerance (e.g. two PredNet scores nearly tied, so the descending sort flips) cannot pin parity."""
        sd = synth.seeded_state_dict(shapes, seed=0)
        sd64 = {k: v.double() for k, v in sd.items()}

        def to64(x):
            if isinstance(x, dict):
                return {k: to64(v) for k, v in x.items()}
            if isinstance(x, list):
                return [to64(v) for v in x]
            return x.double() if torch.is_tensor(x) and x.dtype == torch.float32 else x

        with torch.no_grad():
            t32, t64 = {}, {}
            o32 = O.net_forward(sd, synth.collate(scenes), t32)
            o64 = O.net_forward(sd64, to64(synth.collate(scenes)), t64)
        worst = 0.0
        for a, b in [(torch.cat(o32[k]), torch.cat(o64[k])) for k in ("cls", "reg")] + [(t32[s], t64[s]) for s in STAGES]:
            worst = max(worst, float(((a.double() - b).abs() / (1e-5 + 1e-4 * a.double().abs())).max()))
        return worst

    for name in ("tiny_b3", "argo_b1"):
        seed0 = SEEDS[name]
        while True:
            scenes = golden_scenes(name, seed0)
            c = conditioning(scenes)
            print(name, "seed0", seed0, "reference fp32-vs-fp64 noise / tolerance:", round(c, 3))
            if c < 0.5:
                break
            seed0 += 1
            assert seed0 < SEEDS[name] + 20, "no well-conditioned fixture in 20 seeds: look at the generator"
        out, taps, idx_log, graph = run_reference(net, ref_lanegcn, ref_data, scenes)
        rec = {"cls": torch.cat(out["cls"]).numpy(), "reg": torch.cat(out["reg"]).numpy(),
               "seed0": np.asarray(seed0), "ref_noise_over_tol": np.asarray(c)}
        # Att scatter indices (hi) in call order: 14 index_add_ per LaneConv block x 4 blocks for MapNet, then
        # A2M's two Att layers, M2M's 56, M2A's two, A2A's two  (lanegcn.py:702-703)
        n_lc = 14 * 4
        att_hi = [idx_log[n_lc], idx_log[n_lc + 2 + n_lc], idx_log[n_lc + 2 + n_lc + 2]]
        for tag, hi in zip(("a2m", "m2a", "a2a"), att_hi):
            rec[f"hi_{tag}"] = hi.numpy()
        # wi is not passed to index_add_: recompute both with the reference's own expression sequence
        batch = ref_data.collate_fn(copy.deepcopy(scenes))
        nctr, actr = [g["ctrs"] for g in batch["graph"]], batch["ctrs"]
        for tag, (a, c, th) in {"a2m": (nctr, actr, 7.0), "m2a": (actr, nctr, 6.0), "a2a": (actr, actr, 100.0)}.items():
            hi, wi, hc, wc = [], [], 0, 0
            for x, y in zip(a, c):
                d = x.view(-1, 1, 2) - y.view(1, -1, 2)
                d = torch.sqrt((d ** 2).sum(2))
                idcs = torch.nonzero(d <= th, as_tuple=False)
                if len(idcs) == 0:
                    continue
                hi.append(idcs[:, 0] + hc); wi.append(idcs[:, 1] + wc)
                hc += len(x); wc += len(y)
            assert torch.equal(torch.cat(hi), torch.from_numpy(rec[f"hi_{tag}"])), tag
            rec[f"wi_{tag}"] = torch.cat(wi).numpy()
        if name == "tiny_b3":
            for s in STAGES:
                rec[f"stage_{s}"] = taps[s].numpy()
            for k1 in ("pre", "suc"):
                for i in range(6):
                    for k2 in ("u", "v"):
                        rec[f"g_{k1}{i}_{k2}"] = graph[k1][i][k2].numpy()
            for k1 in ("left", "right"):
                for k2 in ("u", "v"):
                    rec[f"g_{k1}_{k2}"] = graph[k1][k2].numpy()
            for k in ("feats", "turn", "control", "intersect"):
                rec[f"g_{k}"] = graph[k].numpy()
        else:
            for s in STAGES:  # full tensors would be 0.8 MB each: keep a strided sample + float64 sums
                t = taps[s].double()
                rec[f"sum_{s}"] = np.asarray([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()])
                rec[f"rows_{s}"] = taps[s][::37].numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
        print(name, {k: v.shape for k, v in rec.items() if not k.startswith("g_")})

    g = golden_scenes("tiny_b3", int(np.load(os.path.join(HERE, "tiny_b3.npz"))["seed0"]))[0]["graph"]
    rec = {}
    for d in ("pre", "suc"):
        ref = ref_data.dilated_nbrs({"u": g[d][0]["u"].astype(np.int64), "v": g[d][0]["v"].astype(np.int64)},
                                    g["num_nodes"], 6)
        for i, e in enumerate(ref):
            rec[f"{d}{i + 1}_u"], rec[f"{d}{i + 1}_v"] = e["u"], e["v"]
    np.savez_compressed(os.path.join(HERE, "dilate_tiny.npz"), **rec)


if __name__ == "__main__":
    main()
