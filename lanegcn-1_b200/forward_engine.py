"""Capacity buckets for the one-call forward (``lgcn_forward``, include/lgcn.h).

The reference's forward pays one host synchronisation per scene per Att layer (``lanegcn.py:680-681``) and ~50 kernel
launches per LaneConv block from Python.  Here a whole forward is ONE CUDA graph launch: every batch-dependent row
count (nodes, actors, pairs) lives in device memory, all buffers are sized by CAPACITIES, and batches are sorted into
capacity buckets (12.5 % geometric steps), so a graph captured for a bucket is replayed for every later batch that
fits it.  A bucket owns ``N_SLOTS`` slots (static input / output buffers + workspace + one graph each): batch i+1 is
staged (host pack -> pinned -> H2D on the copy stream) into the other slot while batch i computes.

Pair-list capacities cannot be known before the distances are evaluated on the device.  They start from a cheap
bound (the dense count, capped at ``PAIRS_PER_ACTOR`` per actor), the kernels clamp and raise a flag in ``status``,
and the caller re-runs an overflowed batch with the exact counts (``Net.forward`` does this transparently).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _C

C_ = 128
N_SLOTS = 2
PAIRS_PER_ACTOR = 64
ST_OVERFLOW, ST_EMPTY = 0b000111, 0b111000


def round_cap(x: int, lo: int = 64) -> int:
    """Smallest value of the form m * 2^k (m in 8..15) that is >= max(x, lo): at most 12.5 % above x."""
    x = max(int(x), lo)
    k = max(x.bit_length() - 4, 0)
    m = -(-x >> k) if k else x
    if m > 15:
        k, m = k + 1, -(-x >> (k + 1))
    return m << k


class Caps(tuple):
    """(nodes, actors, index, scenes, pairs_a2m, pairs_m2a, pairs_a2a, idx_bytes, n_scales)"""
    __slots__ = ()
    nodes = property(lambda s: s[0])
    actors = property(lambda s: s[1])
    index = property(lambda s: s[2])
    scenes = property(lambda s: s[3])
    pairs = property(lambda s: s[4:7])
    idx_bytes = property(lambda s: s[7])
    n_scales = property(lambda s: s[8])


_IDX_DTYPE = {2: torch.int16, 4: torch.int32, 8: torch.int64}


class Slot:
    """Static device buffers of ONE in-flight batch of a bucket (+ the CUDA graph that works on them)."""

    def __init__(self, caps: Caps, dev, cfg: dict, keep_quirk: bool):
        lib = _C.lib()
        self.caps, self.dev = caps, dev
        N, A, B = caps.nodes, caps.actors, caps.scenes
        f32, i32 = torch.float32, torch.int32
        # ---- inputs: one float arena with capacity-based region offsets, the local indices, two small tables
        self.fl_off = np.cumsum([0, 2 * N, 2 * N, 2 * N, N, N, 60 * A, 2 * A, 4 * B, 2 * B]).tolist()
        self.fl = torch.zeros(self.fl_off[-1], dtype=f32, device=dev)
        o = self.fl_off
        self.node_ctrs, self.node_feats = self.fl[o[0]:o[1]].view(N, 2), self.fl[o[1]:o[2]].view(N, 2)
        self.turn, self.control, self.intersect = self.fl[o[2]:o[3]].view(N, 2), self.fl[o[3]:o[4]], self.fl[o[4]:o[5]]
        self.actor_feats, self.actor_ctrs = self.fl[o[5]:o[6]].view(A, 20, 3), self.fl[o[6]:o[7]].view(A, 2)
        self.rot, self.orig = self.fl[o[7]:o[8]].view(B, 2, 2), self.fl[o[8]:o[9]].view(B, 2)
        self.local = torch.zeros(max(caps.index, 1), dtype=_IDX_DTYPE[caps.idx_bytes], device=dev)
        self.n_seg = 2 * (2 * caps.n_scales + 2) * B
        self.t64 = torch.zeros(2 * self.n_seg + 1, dtype=torch.int64, device=dev)          # seg_start | seg_add
        self.t32 = torch.zeros(2 * (B + 1) + 4, dtype=i32, device=dev)                      # node_off | actor_off | dims
        self.node_off, self.actor_off, self.dims = self.t32[:B + 1], self.t32[B + 1:2 * B + 2], self.t32[2 * B + 2:]
        # ---- state and outputs
        self.nodes = torch.zeros(N, C_, dtype=f32, device=dev)
        self.actors = torch.zeros(A, C_, dtype=f32, device=dev)
        self.actors_t = torch.zeros(A, 3, 20, dtype=f32, device=dev)
        K, T = cfg["num_mods"], cfg["num_preds"]
        self.cls = torch.zeros(A, K, dtype=f32, device=dev)
        self.reg = torch.zeros(A, K, T, 2, dtype=f32, device=dev)
        self.status = torch.zeros(8, dtype=i32, device=dev)
        self.status_host = torch.zeros(8, dtype=i32).pin_memory()
        a = self.args = _C.ForwardArgs()
        a.cap_nodes, a.cap_actors, a.cap_index, a.cap_scenes = N, A, caps.index, B
        a.cap_pairs = (ctypes.c_int64 * 3)(*caps.pairs)
        a.n_scales, a.idx_bytes, a.keep_pair_quirk = caps.n_scales, caps.idx_bytes, int(keep_quirk)
        a.dist_th = (ctypes.c_float * 3)(cfg["actor2map_dist"], cfg["map2actor_dist"], cfg["actor2actor_dist"])
        a.stages = _C.STAGE_ALL
        for name in ("dims", "node_off", "actor_off", "node_ctrs", "node_feats", "turn", "control", "intersect",
                     "actor_ctrs", "nodes", "actors", "status"):
            setattr(a, name, getattr(self, name).data_ptr())
        a.local_idx, a.segs = self.local.data_ptr(), self.t64.data_ptr()
        nbytes = lib.lgcn_forward_workspace_bytes(ctypes.byref(a))
        if nbytes < 0:
            raise RuntimeError("lgcn forward_workspace_bytes: " + lib.lgcn_last_error().decode())
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        a.workspace = self.ws.data_ptr()
        self.actor_ws = torch.zeros(lib.lgcn_actor_net_tc_workspace_bytes(A), dtype=torch.uint8, device=dev)
        # ---- execution state
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graph_kernels = 0                       # lgcn kernels per replay (counted at capture)
        self.weights_version = None                  # the graph bakes the weight-pack pointers of this version
        self.done: Optional[torch.cuda.Event] = None       # last forward on this slot has finished reading the inputs
        self.d2h_done: Optional[torch.cuda.Event] = None   # last read-back of cls / reg / status has finished
        self.status_ev: Optional[torch.cuda.Event] = None

    def buffer(self, which: int, dtype, n: int) -> torch.Tensor:
        """A typed COPY of one of the workspace's index arrays (lgcn_forward_buffer; parity tests)."""
        p = _C.lib().lgcn_forward_buffer(ctypes.byref(self.args), which)
        if not p:
            raise RuntimeError("lgcn_forward_buffer: unknown selector")
        off = p - self.ws.data_ptr()
        item = torch.empty(0, dtype=dtype).element_size()
        return self.ws[off: off + n * item].view(dtype).clone()


class Bucket:
    def __init__(self, caps: Caps, dev, cfg: dict, keep_quirk: bool):
        self.caps = caps
        self.slots = [Slot(caps, dev, cfg, keep_quirk) for _ in range(N_SLOTS)]
        self.turn = 0

    def next_slot(self) -> Slot:
        self.turn = (self.turn + 1) % len(self.slots)
        return self.slots[self.turn]


def pair_caps(node_sizes: List[int], actor_sizes: List[int], learned: Optional[List[int]], cap_actors: int) -> Tuple[int, int, int]:
    """Capacities of the A2M / M2A / A2A pair lists: the dense count when it is small, otherwise PAIRS_PER_ACTOR per
    actor, raised to what earlier batches needed (+25 %)."""
    dense_na = sum(n * a for n, a in zip(node_sizes, actor_sizes))
    dense_aa = sum(a * a for a in actor_sizes)
    guess = max(4096, PAIRS_PER_ACTOR * cap_actors)
    caps = [min(dense_na, guess), min(dense_na, guess), min(dense_aa, guess)]
    dense = [dense_na, dense_na, dense_aa]
    if learned is not None:
        caps = [min(d, max(c, int(1.25 * s))) for c, s, d in zip(caps, learned, dense)]
    return tuple(round_cap(c) for c in caps)
