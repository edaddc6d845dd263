#!/usr/bin/env python
"""bench.py — LaneGCN forward scenes/sec (batch 128) on 1/2/4/8 B200 + LaneConv rooflines.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1..5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Default workload = BASELINE.json configs[2] (--config 3): one global batch of 128 synthetic Argoverse-shaped scenes
("argo-1.5k": 1,512 lane nodes, 20 actors, ~18.8k edges per scene), cut into contiguous scene shards across the N
ranks (strong scaling).  A step = one full Net forward of the rank's shard (+ the final NCCL gather of cls/reg at
N > 1).  ``value`` times the step from the staged (HBM-resident) batch with CUDA events: one CUDA-graph replay per
step on the one-call path.  ``e2e`` times collated HOST dicts -> cls/reg back on the host (pack into pinned memory,
H2D, forward, D2H inside the timed region).  ``roofline*`` = per-kernel CUDA-event timings of an eager pass over the
same staged inputs right after the timed region.  ``cpu_baseline`` / ``--impl reference`` time the UNMODIFIED reference
(oracle/_ref, built by oracle/make_ref.py) on the host cores; ``gpu_eager_reference`` times the same reference modules
under stock PyTorch eager on this GPU.  Other configs: 1 = batch 1, 2 = batch 32, 4 = MapNet only on one 100,800-node
graph (both LaneConv paths), 5 = LaneRCNN graph layers.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import copy
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PRESET = "argo-1.5k"
CONFIG_BATCH = {1: 1, 2: 32, 3: 128}
ONE_THREAD_SCENES = 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--batch", type=int, default=None, help="override the batch size of configs 1-3")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-reference", action="store_true")
    ap.add_argument("--debug-flags", type=int, default=0,
                    help="A/B switches of include/lgcn_debug.h that keep results right (e.g. 16384 no PDL, 65536 / 131072 unfused Att pieces)")
    return ap.parse_args()


def metric_of(cfg: int, batch: int):
    if cfg == 4:
        return "LaneGCN MapNet-only fwd graphs/sec (100,800-node graph)", "graphs/s"
    if cfg == 5:
        return f"LaneRCNN graph-layer fwd scenes/sec (batch {batch})", "scenes/s"
    return f"LaneGCN fwd scenes/sec (batch {batch})", "scenes/s"


def workload_of(cfg: int, batch: int, world: int = 1):
    if cfg == 4:
        return "MapNet only on one synthetic city-scale lane graph (100,800 nodes, 4 fuse blocks, 6 scales) (BASELINE configs[3])"
    if cfg == 5:
        return f"LaneRCNN graph layers on {batch} synthetic scenes with lane-RoI sub-graphs (BASELINE configs[4])"
    return (f"LaneGCN forward, batch {batch} synthetic {PRESET} scenes (BASELINE configs[{cfg - 1}])"
            + (f", scene-sharded over {world} GPU(s)" if cfg == 3 else ""))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), float(j["bf16_tflops"]), "MEASURED_PEAKS.json (hbm_gbs, bf16_tflops burst)"
    return 6650.0, 1590.0, "fallback of B200_PROFILING.md (6.65 TB/s, 1.59 PFLOP/s bf16)"


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def weights():
    from lanegcn_b200 import synth
    shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_shapes.json")))
    return synth.seeded_state_dict(shapes, 0)


# ----------------------------------------------------------------------------- the reference on the host CPU
def reference_cpu_rate(scenes, steps: int, warmup: int, threads: int):
    """(scenes/s, ms per forward, threads, kind) of the reference forward on CPU: the unmodified reference modules from
    oracle/_ref (or /root/reference where it exists); the oracle port only if neither is there."""
    import torch

    from lanegcn_b200 import synth
    from oracle import ref_loader

    torch.set_num_threads(max(1, threads))
    sd = weights()
    if ref_loader.available():
        ref, ref_data = ref_loader.load(keep_gpu=False)
        net = ref.Net(ref.config).eval()
        net.load_state_dict(sd)
        data = ref_data.collate_fn(copy.deepcopy(scenes))
        fwd, kind = (lambda: net(data)), "reference"
    else:
        from oracle import lanegcn_oracle as O
        data = synth.collate(scenes)
        fwd, kind = (lambda: O.net_forward(sd, data)), "port"
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            fwd()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    return len(scenes) / (ms / 1e3), ms, torch.get_num_threads(), kind


def run_reference(args):
    """--impl reference: the reference's own CPU forward of the SAME workload (all scenes of the batch per step), all
    host threads, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from lanegcn_b200 import synth

    cfg = args.config
    B = args.batch or CONFIG_BATCH.get(cfg, 128)
    metric, unit = metric_of(cfg, B)
    if cfg in (4, 5):
        print(json.dumps({"impl": "reference", "metric": metric, "unavailable":
                          "the reference arm is implemented for configs 1-3 (the full LaneGCN forward)"}))
        return
    steps, warmup = max(1, min(args.steps, 50)), max(1, min(args.warmup, 10))
    scenes = synth.make_scenes(B, PRESET, seed0=0)
    v, ms, cores, kind = reference_cpu_rate(scenes, steps, warmup, host_threads())
    sample = (f"all {B} scenes of the batch per step, {warmup} warm-up + {steps} timed forwards of "
              + ("the unmodified reference lanegcn.Net (oracle/_ref, utils.gpu -> identity)" if kind == "reference"
                 else "the oracle port (oracle/_ref missing)") + ", fp32, torch CPU")
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": round(v, 3), "unit": unit, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_of(cfg, B, args.gpus), "global_batch": B},
        "cpu_baseline": {"value": round(v, 3), "unit": unit, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(v, 3), "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def reference_gpu_eager(scenes, dev, steps=5, warmup=2):
    """The unmodified reference modules under stock PyTorch eager on this GPU (utils.gpu intact: one H2D per tensor,
    cudnn.allow_tf32 = False, default stream) — the same-box bar of BASELINE.md.  Host dict in, cls/reg left on the GPU."""
    import torch

    from oracle import ref_loader

    if not ref_loader.available():
        return None
    ref, ref_data = ref_loader.load(keep_gpu=True)
    try:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        with torch.cuda.device(dev):
            net = ref.Net(ref.config)
            net.load_state_dict(weights())
            net = net.cuda().eval()
            data = ref_data.collate_fn(copy.deepcopy(scenes))
            ev = []
            with torch.no_grad():
                for i in range(warmup + steps):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    out = net(data)
                    b.record()
                    if i >= warmup:
                        ev.append((a, b))
                torch.cuda.synchronize()
            ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
            del net, out
        return {"value": round(len(scenes) / (ms / 1e3), 2), "unit": "scenes/s", "ms_per_step": round(ms, 3),
                "how": f"unmodified reference lanegcn.Net (oracle/_ref), stock PyTorch {torch.__version__} eager on this GPU, "
                       f"utils.gpu intact (host dict in), allow_tf32=False, {warmup} warm-up + {steps} timed forwards, CUDA events"}
    finally:
        ref_loader.load(keep_gpu=False)


# ----------------------------------------------------------------------------- clocks sampler
class Clocks:
    """nvidia-smi sampler (every 50 ms) running from before the warm-up to the end of the timed region.  Samples are
    time-stamped by a reader thread; the summary uses those inside the timed window when there are any, otherwise all
    samples taken under load (warm-up + timed steps: a 10-step timed region can be shorter than one sampling period)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        import threading
        self.p, self.rows, self.t0, self.t1 = None, [], None, None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return

        def reader():
            for line in self.p.stdout:
                self.rows.append((time.perf_counter(), line))
        self.th = threading.Thread(target=reader, daemon=True)
        self.th.start()

    def wait_first(self, timeout=4.0):
        t = time.perf_counter()
        while self.p is not None and not self.rows and time.perf_counter() - t < timeout:
            time.sleep(0.02)

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.th.join(timeout=2)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def summarise(rows):
            sm, mx, reasons = [], [], set()
            for _, line in rows:
                f = [x.strip() for x in line.split(",")]
                if len(f) < 6 or not f[0].isdigit():
                    continue
                sm.append(int(f[0]))
                mx.append(int(f[1]))
                reasons |= {n for n, x in zip(names, f[2:6]) if x.lower().startswith("active")}
            sm.sort()
            return sm, mx, reasons
        inside = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or 1e30)]
        window = "timed region"
        sm, mx, reasons = summarise(inside)
        if not sm:
            window = "warm-up + timed region (the timed region is shorter than a sampling period)"
            sm, mx, reasons = summarise(self.rows[1:] if len(self.rows) > 1 else self.rows)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def traffic_of(name: str, n_nodes: int):
    """DRAM bytes per launch from the committed `ncu --set full` capture of that kernel (profiles/<name>.json), scaled
    by the row count when this run's shard differs from the captured one (traffic is linear in the rows)."""
    tp = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(tp):
        return None, None
    tj = json.load(open(tp))
    if not tj.get("n_nodes") or not tj.get("dram_bytes_per_launch"):
        return None, None
    scale = n_nodes / float(tj["n_nodes"])
    src = f"ncu --set full capture in profiles/{name} at {tj['n_nodes']} rows" + ("" if scale == 1.0 else f", scaled x{scale:.4f} by rows")
    return int(tj["dram_bytes_per_launch"] * scale), src


def collect_prof(lib):
    from lanegcn_b200 import _C
    ms, n = (ctypes.c_double * 8)(), (ctypes.c_int64 * 8)()
    _C.check(lib.lgcn_prof_collect(ms, n), "prof_collect")
    return list(ms), list(n)


def laneconv_rooflines(ms_f, n_f, ms_s, n_s, n_nodes, n_edges, how=None):
    """roofline dicts of the dominant kernel (aggregate-first LaneConv block) and of the split path's gather / GEMM."""
    hbm, bf16, src = peaks()
    roof = roof_g = roof_w = None
    kind = 5 if n_f[5] else 4          # 5: the block's main kernel alone; 4: block incl. the multi-source pre-pass
    if n_f[kind]:
        f_ms = ms_f[kind] / n_f[kind]
        useful = 2.0 * n_nodes * 128 * (16 * 128) / (f_ms * 1e-3) / 1e12
        traffic, tsrc = traffic_of("fused_traffic.json", n_nodes)
        roof = {"kernel": "k_laneconv_v2 (one LaneConv block: neighbour gather -> 15 projections -> GN+ReLU -> ctr2 -> "
                          "GN + residual + ReLU, tcgen05 3xTF32)" + ("" if kind == 5 else " incl. its multi-source pre-pass"),
                "bound": "tensor", "achieved": round(3 * useful, 1), "peak": round(bf16 / 2, 1), "unit": "TFLOP/s",
                "frac": round(3 * useful / (bf16 / 2), 4), "traffic": traffic, "traffic_source": tsrc,
                "useful_fp32_tflops": round(useful, 1), "useful_frac_of_tf32_peak": round(useful / (bf16 / 2), 4),
                "avg_launch_ms": round(f_ms, 5), "launches_timed": int(n_f[kind]),
                "block_ms_incl_prepass": round(ms_f[4] / n_f[4], 5) if n_f[4] else None,
                "flops_per_launch": 3 * 2 * n_nodes * 128 * 16 * 128,
                "peak_source": f"tf32 dense peak taken as half of the cuBLAS bf16 burst of {src}; executed flops = 3 x 2*N*128*2048",
                "hbm_floor_ms": round((3 * 512 + 60) * n_nodes / (hbm * 1e9) * 1e3, 4),
                "how": how or "CUDA events around every launch"}
    if n_s[1]:
        gather_bytes = 4 * 128 * (n_edges + 2 * n_nodes) + 4 * n_edges + 4 * (n_nodes + 1)
        g_ms = ms_s[1] / n_s[1]
        ach = gather_bytes / (g_ms * 1e-3) / 1e9
        traffic, tsrc = traffic_of("gather_traffic.json", n_nodes)
        roof_g = {"kernel": "k_gather_gn_relu (LaneConv CSR gather + GN + ReLU of the split path)", "bound": "hbm",
                  "achieved": round(ach, 1), "peak": round(hbm, 1), "unit": "GB/s", "frac": round(ach / hbm, 4),
                  "traffic": traffic, "traffic_source": tsrc, "peak_source": src, "algorithmic_bytes_per_launch": gather_bytes,
                  "avg_launch_ms": round(g_ms, 5), "launches_timed": int(n_s[1]),
                  "how": "separate eager pass with the split LaneConv path (LGCN_LANECONV=split) after the timed region"}
    if n_s[0]:
        w_ms = ms_s[0] / n_s[0]
        useful_w = 2.0 * n_nodes * 128 * 1920 / (w_ms * 1e-3) / 1e12
        roof_w = {"kernel": "k_wide_tc (split path: wide projection [N,128]x[128,1920], tcgen05 3xTF32)", "bound": "tensor",
                  "achieved": round(3 * useful_w, 1), "peak": round(bf16 / 2, 1), "unit": "TFLOP/s",
                  "frac": round(3 * useful_w / (bf16 / 2), 4), "avg_launch_ms": round(w_ms, 5), "launches_timed": int(n_s[0]),
                  "split_path_ms_per_block": {"wide": round(w_ms, 4), "gather": round(ms_s[1] / max(1, n_s[1]), 4),
                                              "ctr2": round(ms_s[2] / max(1, n_s[2]), 4)}}
    return roof, roof_g, roof_w


def edge_count(scenes):
    return sum(sum(len(e["u"]) for e in s["graph"]["pre"] + s["graph"]["suc"])
               + len(s["graph"]["left"]["u"]) + len(s["graph"]["right"]["u"]) for s in scenes)


# ----------------------------------------------------------------------------- our arm: full forward (configs 1-3)
def run_forward(args):
    import torch
    import torch.distributed as dist

    from lanegcn_b200 import _C, build, shard, synth
    from lanegcn_b200 import lanegcn as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if not os.path.exists(_C.LIB_PATH):
        if rank == 0:
            build.build()
        if world > 1:
            dist.barrier()
    lib = _C.lib()
    if args.debug_flags:
        lib.lgcn_debug_flags(args.debug_flags)

    # ---- workload: global batch, this rank's contiguous shard
    cfg = args.config
    B = args.batch or CONFIG_BATCH[cfg]
    metric, unit = metric_of(cfg, B)
    mine = shard.partition([1512] * B, world)[rank]   # every argo-1.5k scene has 1,512 nodes: equal counts per rank
    scenes = [synth.make_scene(i, PRESET) for i in mine]
    data = synth.collate(scenes)
    net = L.Net(L.config)
    net.load_state_dict(weights())
    net = net.to(dev).eval()
    plan = shard.make_plan([len(s["ctrs"]) for s in scenes]) if world > 1 else None  # host metadata, once per batch

    def step_device(staged):
        out = net.forward_device(staged)
        return shard.gather_outputs(out, plan, lazy=True) if world > 1 else out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    staged = net.stage(data)
    onecall = staged.slot is not None
    step_device(staged)   # first call: lazy initialisation, graph capture
    if net.check(staged):                       # pair capacities learned from this batch: stage into the right bucket
        staged = net.stage(data)
        step_device(staged)
        assert not net.check(staged)
    sync_all()
    clocks = Clocks(local) if rank == 0 else None
    if clocks:
        clocks.wait_first()
    W = max(args.warmup, 3)
    for _ in range(W):
        step_device(staged)
    sync_all()

    # ---- timed: K steps from HBM-resident inputs, CUDA events per step, L2 flushed between steps
    n_nodes = sum(int(s["graph"]["num_nodes"]) for s in scenes)
    n_edges = edge_count(scenes)
    launches0, replayed0 = lib.lgcn_launch_count(), getattr(net, "replayed_kernels", 0)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    if clocks:
        clocks.mark_start()
    for a, b in ev:
        flush.zero_()
        a.record()
        out_last = step_device(staged)
        b.record()
    sync_all()
    if clocks:
        clocks.mark_end()
    launches = (lib.lgcn_launch_count() - launches0) + (getattr(net, "replayed_kernels", 0) - replayed0)
    staged.checked = None
    assert not net.check(staged), "pair capacity overflow inside the timed region"
    clk = clocks.stop() if clocks else None
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    lt = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    ms_per_step = float(t.item()) / args.steps
    value = B / (ms_per_step / 1e3)

    # ---- N > 1: the NCCL-gathered result against ONE rank's forward of the whole batch
    gather_verified = None
    if world > 1:
        got_cls, got_reg = out_last["cls"].cat.clone(), out_last["reg"].cat.clone()
        sync_all()
        if rank == 0:
            full = synth.collate([synth.make_scene(i, PRESET) for i in range(B)])
            ref_out = net(full)
            rc, rr = torch.cat(ref_out["cls"]), torch.cat(ref_out["reg"])
            ok = got_cls.shape == rc.shape and got_reg.shape == rr.shape
            dc = float((got_cls - rc).abs().max()) if ok else float("inf")
            dr = float(((got_reg - rr).abs() / (1e-5 + 1e-5 * rr.abs())).max()) if ok else float("inf")
            gather_verified = {"ok": bool(ok and dc <= 1e-4 and dr <= 1.0), "max_abs_diff_cls": dc, "max_err_over_tol_reg(1e-5,1e-5)": dr,
                               "how": f"all_gather result of {world} ranks vs a single-rank forward of the whole batch on rank 0"}
            del ref_out
        sync_all()

    # ---- end to end: collated HOST dicts -> cls/reg on the host.  Every step packs its batch into pinned memory,
    # copies it H2D, runs the forward and reads the result back (D2H); the staging of step i+1 is overlapped with the
    # device work of step i (lanegcn.prefetch_forward — a DataLoader-style prefetch).  Wall clock, max over ranks.
    # At N > 1 every rank takes part in the gather; rank 0 alone reads the gathered result back.
    data_packed = L.pack_batch(synth.collate(scenes))   # one host blob per sample, made once (the dataset's job)

    def run_e2e(n, batch):
        gather = (lambda o: shard.gather_outputs(o, plan, lazy=True)) if world > 1 else None
        mode = True if rank == 0 else "defer"
        res = None
        for out in L.prefetch_forward(net, (batch for _ in range(n)), to_host=mode, post=gather, lazy_lists=True):
            res = out
        return res

    def time_e2e(batch):
        run_e2e(max(args.warmup, 8), batch)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(5):
            net.stage(batch)
        host_ms = 1e3 * (time.perf_counter() - t0) / 5
        sync_all()
        t0 = time.perf_counter()
        res = run_e2e(args.steps, batch)
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0) / args.steps
        sync_all()
        return ms, host_ms, res

    e2e_ms, stage_host_ms, res = time_e2e(data_packed)
    e2e_dict_ms, stage_dict_ms, _ = time_e2e(data)
    lat = []
    for _ in range(3):   # un-overlapped latency of one public call + read-back
        t0 = time.perf_counter()
        o = step_device(net.stage(data_packed)) if world > 1 else net(data_packed)
        if rank == 0:
            L._cat_of(o["cls"]).cpu(), L._cat_of(o["reg"]).cpu()
        torch.cuda.synchronize()
        lat.append(1e3 * (time.perf_counter() - t0))
        sync_all()
    lat_ms = sum(lat) / len(lat)
    te = torch.tensor([e2e_ms, lat_ms, stage_host_ms, e2e_dict_ms, stage_dict_ms], dtype=torch.float64, device=dev)
    hb = torch.tensor([float(net.stage(data_packed).h2d_bytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(hb, op=dist.ReduceOp.SUM)
    e2e_ms, lat_ms, stage_host_ms, e2e_dict_ms, stage_dict_ms = (float(x) for x in te.tolist())
    d2h = int(sum(t.numel() for t in res["cls"]) * 4 + sum(t.numel() for t in res["reg"]) * 4) if rank == 0 else 0
    sync_all()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel timings (rank 0's shard; the other ranks have left).  The bucket's graph is captured once more with
    # event-record nodes around the LaneConv / Att launches (lgcn_prof_enable(2)) and replayed like the timed steps (L2
    # flushed): same kernels, same stream priorities and branch overlap as the timed region, read after every replay.
    n_prof = max(3, min(args.steps, 10))
    ms_f, n_f = [0.0] * 8, [0] * 8
    if onecall:
        slot = staged.slot
        saved_graph = slot.graph
        slot.graph = None
        lib.lgcn_prof_enable(2)
        net.forward_device(staged)            # capture with event nodes (+ one replay)
        torch.cuda.synchronize()
        for _ in range(n_prof):
            flush.zero_()
            net.forward_device(staged)
            torch.cuda.synchronize()
            ms_i, n_i = (ctypes.c_double * 8)(), (ctypes.c_int64 * 8)()
            _C.check(lib.lgcn_prof_peek(ms_i, n_i), "prof_peek")
            ms_f = [x + y for x, y in zip(ms_f, ms_i)]
            n_f = [x + y for x, y in zip(n_f, n_i)]
        lib.lgcn_prof_enable(0)
        collect_prof(lib)                     # forget the events
        slot.graph = saved_graph
        prof_how = ("event-record nodes around every launch inside the step's CUDA graph (lgcn_prof_enable(2)), replayed "
                    f"{n_prof} times with the L2 flush right after the timed region")
    net.use_cuda_graphs = False
    if not onecall:
        for _ in range(2):
            net.forward_device(staged)
        torch.cuda.synchronize()
        lib.lgcn_prof_enable(1)
        for _ in range(n_prof):
            flush.zero_()
            net.forward_device(staged)
        torch.cuda.synchronize()
        lib.lgcn_prof_enable(0)
        ms_f, n_f = collect_prof(lib)
        prof_how = "CUDA events around every launch in an eager pass over the same staged inputs right after the timed region"
    # ... and the prescribed pair (wide projection + CSR gather + ctr2): its gather is the HBM-bound kernel the north
    # star asks to hold >= 60 % of the copy bandwidth
    L.LANECONV_FUSED = False
    staged_split = net.stage(data)
    for _ in range(2):
        net.forward_device(staged_split)
    torch.cuda.synchronize()
    lib.lgcn_prof_enable(1)
    for _ in range(3):
        flush.zero_()
        net.forward_device(staged_split)
    torch.cuda.synchronize()
    lib.lgcn_prof_enable(0)
    ms_s, n_s = collect_prof(lib)
    L.LANECONV_FUSED = True
    net.use_cuda_graphs = True
    roof, roof_g, roof_w = laneconv_rooflines(ms_f, n_f, ms_s, n_s, n_nodes, n_edges, prof_how)
    kernel_ms = {k: round(ms_f[i] / n_prof, 4) for i, k in [(5, "k_laneconv_v2 (8 launches)"), (4, "LaneConv blocks incl. multi-source pre-pass (8)"),
                                                            (3, "att (6 layers)")] if n_f[i]}

    line = {
        "metric": metric, "value": round(value, 2), "unit": unit, "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_of(cfg, B, world),
                   "global_batch": B, "scenes_per_gpu": len(scenes), "nodes_per_gpu": n_nodes, "edges_per_gpu": n_edges,
                   "gemm_engine": ["simt-fp32", "tcgen05-3xtf32"][lib.lgcn_get_gemm_engine()],
                   "forward": "one CUDA-graph replay per step (lgcn_forward + ActorNet/PredNet), device-side sizes, no host sync"
                              if onecall else "module path",
                   "l2": "256 MiB flush between timed steps; per-step working set (two feature buffers "
                         f"{2 * n_nodes * 512 / 1e6:.0f} MB + plan + aux rows, re-read by 8 LaneConv blocks) exceeds the 126 MB L2",
                   "kernel_ms_per_step": kernel_ms},
        "e2e": {"value": round(B / (e2e_ms / 1e3), 2), "unit": unit, "ms_per_step": round(e2e_ms, 4),
                "single_call_latency_ms": round(lat_ms, 4), "stage_host_ms": round(stage_host_ms, 4),
                "from_dicts": {"value": round(B / (e2e_dict_ms / 1e3), 2), "ms_per_step": round(e2e_dict_ms, 4),
                               "stage_host_ms": round(stage_dict_ms, 4),
                               "how": "same pipeline fed with the reference's dict-of-lists batch (~40 tensors per scene "
                                      "walked in Python) instead of packed scenes"},
                "how": "host inputs = one packed blob per scene (lanegcn.pack_scene, made once per sample like the reference's "
                       "preprocessed pickles); every step: Net.stage (lgcn_stage_scenes assembles the batch in pinned "
                       "memory, 4 H2D copies into the bucket's static buffers) + Net.forward_device (one graph replay) + "
                       "D2H of cls/reg; staging of step i+1 overlapped with the device work of step i (prefetch_forward "
                       "with to_host=True: D2H on its own stream, results handed out one batch later)" + ("; every rank joins the NCCL gather, rank 0 alone reads the gathered result back" if world > 1 else ""),
                "h2d_bytes_per_step": int(hb.item()), "d2h_bytes_per_step": d2h},
        "gpu_launches": int(lt.item()),
        "roofline": roof if roof else roof_g,
        "roofline_gather": roof_g,
        "roofline_gemm": roof_w,
        "clocks": clk,
    }
    if gather_verified is not None:
        line["gather_verified"] = gather_verified
    if world == 1 and not args.no_cpu_baseline:
        full = synth.make_scenes(B, PRESET, seed0=0)
        v, ms, cores, kind = reference_cpu_rate(full, 3, 1, host_threads())
        line["cpu_baseline"] = {"value": round(v, 3), "unit": unit, "cores": cores, "kind": kind, "ms_per_step": round(ms, 2),
                                "sample": f"all {B} scenes of the batch, 1 warm-up + 3 timed forwards of "
                                          + ("the unmodified reference (oracle/_ref)" if kind == "reference" else "the oracle port")
                                          + " on the host CPU, fp32"}
        n1 = min(B, ONE_THREAD_SCENES)
        v1, ms1, c1, kind1 = reference_cpu_rate(full[:n1], 2, 1, 1)
        line["cpu_baseline_1thread"] = {"value": round(v1, 3), "unit": unit, "cores": c1, "kind": kind1,
                                        "sample": f"{n1} of the {B} scenes, 1 warm-up + 2 timed forwards, torch.set_num_threads(1) "
                                                  "(the reference scripts pin OMP/MKL threads to 1: train.py:8-10)"}
    if world == 1 and not args.no_eager_reference:
        del flush
        torch.cuda.empty_cache()
        ge = reference_gpu_eager(synth.make_scenes(B, PRESET, seed0=0), dev)
        if ge:
            line["gpu_eager_reference"] = ge
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- config 4: MapNet only, one 100k-node graph
def run_mapnet_city(args):
    import torch

    from lanegcn_b200 import _C, synth
    from lanegcn_b200 import lanegcn as L

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _C.lib()
    metric, unit = metric_of(4, 1)
    scene = synth.make_scene(0, "city-100k")
    batch = synth.collate([scene])
    net = L.Net(L.config)
    net.load_state_dict(weights())
    net = net.to(dev).eval()
    n_nodes, n_edges = int(scene["graph"]["num_nodes"]), edge_count([scene])
    graph = L.graph_gather(batch["graph"])       # HBM-resident batched graph + CSR; the plan is built on first use
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    W = max(args.warmup, 3)

    def timed(n):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in ev:
            flush.zero_()
            a.record()
            net.map_net(graph)
            b.record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev) / n

    for _ in range(W):
        net.map_net(graph)
    torch.cuda.synchronize()
    clocks = Clocks(local) if rank == 0 else None
    if clocks:
        clocks.wait_first()
        clocks.mark_start()
    l0 = lib.lgcn_launch_count()
    lib.lgcn_prof_enable(1)
    ms_step = timed(args.steps)
    lib.lgcn_prof_enable(0)
    launches = lib.lgcn_launch_count() - l0
    if clocks:
        clocks.mark_end()
    ms_f, n_f = collect_prof(lib)
    clk = clocks.stop() if clocks else None
    # split path
    L.LANECONV_FUSED = False
    for _ in range(2):
        net.map_net(graph)
    lib.lgcn_prof_enable(1)
    ms_split = timed(max(3, min(args.steps, 10)))
    lib.lgcn_prof_enable(0)
    ms_s, n_s = collect_prof(lib)
    L.LANECONV_FUSED = True
    # end to end: host dict -> graph_gather (pack, H2D, CSR, plan) -> MapNet -> features back on the host
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 10))
    h2d = 0
    for _ in range(n_e2e):
        g = L.graph_gather(batch["graph"])
        feat, _, _ = net.map_net(g)
        host = feat.cpu()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
    idx_bytes = sum(e[k].numel() * e[k].element_size() for e in batch["graph"][0]["pre"] + batch["graph"][0]["suc"]
                    + [batch["graph"][0]["left"], batch["graph"][0]["right"]] for k in ("u", "v"))
    h2d = 8 * n_nodes * 4 + idx_bytes
    if rank != 0:
        return
    roof, roof_g, roof_w = laneconv_rooflines(ms_f, n_f, ms_s, n_s, n_nodes, n_edges)
    print(json.dumps({
        "metric": metric, "value": round(world * 1e3 / ms_step, 2), "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_of(4, 1), "nodes": n_nodes, "edges": n_edges,
                   "parallelism": "replicas only (one graph does not shard)" if world > 1 else "1 GPU",
                   "laneconv": "aggregate-first single kernel (default); split path timed beside it",
                   "split_path_ms_per_step": round(ms_split, 4),
                   "l2": "256 MiB flush between timed steps",
                   "nodes_per_s": round(n_nodes * 1e3 / ms_step)},
        "e2e": {"value": round(world * 1e3 / e2e_ms, 2), "unit": unit, "ms_per_step": round(e2e_ms, 3),
                "how": "graph_gather(host dict: pack + H2D + CSR + plan) + MapNet + D2H of the [N,128] features, per step, not overlapped",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(host.numel() * 4)},
        "gpu_launches": int(launches) * world,
        "roofline": roof, "roofline_gather": roof_g, "roofline_gemm": roof_w, "clocks": clk,
    }))


# ----------------------------------------------------------------------------- config 5: LaneRCNN graph layers
def run_lanercnn(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    try:
        from lanegcn_b200 import lanercnn_bench
    except ImportError:
        if rank == 0:
            print(json.dumps({"metric": metric_of(5, 32)[0], "unavailable": "config 5 bench is not built yet"}))
        return

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    line = lanercnn_bench.run(args, torch.device("cuda", local), Clocks(local) if rank == 0 else None, metric_of, workload_of)
    if rank == 0:
        print(json.dumps(line))


def main():
    args = parse()
    if args.gpus > 1 and "RANK" not in os.environ:  # convenience: re-launch ourselves under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    elif args.config == 4:
        run_mapnet_city(args)
    elif args.config == 5:
        run_lanercnn(args)
    else:
        run_forward(args)


if __name__ == "__main__":
    main()
