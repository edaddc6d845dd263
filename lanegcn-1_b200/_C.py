"""ctypes binding of the C-ABI CUDA library (include/lgcn.h).

The library is REQUIRED: if ``liblgcn_b200.so`` is missing or a symbol does not resolve this module raises —
there is no CPU / PyTorch fallback for the hot path.  Build it with ``python -m lanegcn_b200.build`` (or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblgcn_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "lgcn.h")
DEBUG_HEADER_PATH = os.path.join(_HERE, "..", "include", "lgcn_debug.h")

EPI_GN, EPI_RELU1, EPI_RES, EPI_RELU2 = 1, 2, 4, 8

_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every function include/lgcn.h declares (tests check this)
_SIGS = {
    "lgcn_version": (_i32, []),
    "lgcn_last_error": (C.c_char_p, []),
    "lgcn_get_gemm_engine": (_i32, []),
    "lgcn_set_gemm_engine": (_i32, [_i32]),
    "lgcn_debug_flags": (_i32, [_i32]),
    "lgcn_launch_count": (_i64, []),
    "lgcn_prof_enable": (_i32, [_i32]),
    "lgcn_prof_collect": (_i32, [_vp, _vp]),
    "lgcn_prof_peek": (_i32, [_vp, _vp]),
    "lgcn_pack_host": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32]),
    "lgcn_stage_scenes": (_i32, [_vp, _i32, _i64, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32]),
    "lgcn_offset_indices": (_i32, [_vp, _i32, _vp, _vp, _i32, _i64, _vp, _vp]),
    "lgcn_pack_meta": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "lgcn_csr_workspace_bytes": (_i64, [_i64, _i64]),
    "lgcn_csr_build": (_i32, [_vp, _vp, _vp, _i32, _i64, _vp, _vp, _vp, _vp, _vp]),
    "lgcn_scatter_csr_build": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "lgcn_scale0_workspace_bytes": (_i64, [_i64]),
    "lgcn_scale0_edges": (_i32, [_vp, _i64, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "lgcn_side_edges_workspace_bytes": (_i64, [_i64, _i64]),
    "lgcn_side_edges": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _f32, _vp, _vp, _vp, _vp, _vp]),
    "lgcn_dilate_workspace_bytes": (_i64, [_i64, _i64]),
    "lgcn_dilate_csr0": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "lgcn_dilate_bound": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "lgcn_dilate_square": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lgcn_linear128": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _i32,
                              _vp, _i64, _i64, _vp]),
    "lgcn_linear128_workspace_bytes": (_i64, []),
    "lgcn_linear128_ws": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _i32,
                                 _vp, _i64, _i64, _vp, _vp]),
    "lgcn_mlp2_in": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "lgcn_mlp4_in": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "lgcn_gather_rows_gn_relu": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "lgcn_laneconv_gather_gn_relu": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "lgcn_segsum_gn_relu": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "lgcn_pairs_workspace_bytes": (_i64, [_i64, _i32]),
    "lgcn_pairs_count": (_i32, [_vp, _vp, _vp, _vp, _i32, _i64, _f32, _i32, _vp, _vp, _vp, _vp]),
    "lgcn_pairs_fill": (_i32, [_vp, _vp, _vp, _vp, _i32, _i64, _f32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lgcn_laneconv_wpack_floats": (_i64, [_i32]),
    "lgcn_laneconv_workspace_bytes": (_i64, [_i64, _i32]),
    "lgcn_laneconv_stack": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _i64, _vp, _vp]),
    "lgcn_debug_timeline": (_i32, [_vp]),
    "lgcn_laneconv_plan_bytes": (_i64, [_i64, _i64, _i32]),
    "lgcn_laneconv_plan_build": (_i32, [_vp, _vp, _i32, _i64, _i64, _vp, _vp]),
    "lgcn_laneconv_planned_workspace_bytes": (_i64, [_i64, _i64, _i32]),
    "lgcn_laneconv_stack_planned": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _i64, _vp, _vp]),
    "lgcn_att_wpack_floats": (_i64, []),
    "lgcn_att_workspace_bytes": (_i64, [_i64, _i64]),
    "lgcn_actor_gather": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp]),
    "lgcn_actor_net_wpack_floats": (_i64, []),
    "lgcn_actor_net_pack": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "lgcn_actor_net": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "lgcn_actor_net_tc_workspace_bytes": (_i64, [_i64]),
    "lgcn_actor_net_tc": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "lgcn_pred_net_wpack_floats": (_i64, []),
    "lgcn_pred_net_pack": (_i32, [_vp, _vp, _vp]),
    "lgcn_pred_net": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "lgcn_world_transform": (_i32, [_vp, _vp, _i32, _vp, _vp, _i64, _vp, _i32, _vp]),
    "lgcn_forward_prepared_bytes": (_i64, [_i32]),
    "lgcn_forward_prepare": (_i32, [_vp, _i32, _vp]),
    "lgcn_forward_workspace_bytes": (_i64, [_vp]),
    "lgcn_forward": (_i32, [_vp, _vp]),
    "lgcn_forward_buffer": (_vp, [_vp, _i32]),
    "lgcn_att_forward": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
}



class ForwardWeights(C.Structure):
    """LgcnForwardWeights (include/lgcn.h)."""
    _fields_ = [("map_input", _vp), ("map_seg", _vp), ("map_fuse", _vp), ("a2m_meta", _vp), ("att", _vp * 6),
                ("m2m_fuse", _vp), ("prepared", _vp)]


class ForwardArgs(C.Structure):
    """LgcnForwardArgs (include/lgcn.h)."""
    _fields_ = [("cap_nodes", _i64), ("cap_actors", _i64), ("cap_index", _i64), ("cap_pairs", _i64 * 3),
                ("cap_scenes", C.c_int32), ("n_scales", C.c_int32), ("idx_bytes", C.c_int32),
                ("keep_pair_quirk", C.c_int32), ("dist_th", _f32 * 3), ("stages", C.c_int32),
                ("dims", _vp), ("node_off", _vp), ("actor_off", _vp), ("node_ctrs", _vp), ("node_feats", _vp),
                ("turn", _vp), ("control", _vp), ("intersect", _vp), ("actor_ctrs", _vp), ("local_idx", _vp),
                ("segs", _vp), ("nodes", _vp), ("actors", _vp), ("status", _vp), ("w", ForwardWeights),
                ("workspace", _vp), ("aux_streams", _vp * 2)]


STAGE_GRAPH, STAGE_MAPNET, STAGE_A2M, STAGE_M2M, STAGE_M2A, STAGE_A2A, STAGE_ALL = 1, 2, 4, 8, 16, 32, 63

_lib = None


def header_symbols() -> list:
    """Function names declared in include/lgcn.h and include/lgcn_debug.h (comments stripped)."""
    src = open(HEADER_PATH).read() + open(DEBUG_HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lgcn_[a-z0-9_]+)\s*\(", src)))


def lib():
    """The loaded library (loads on first use).  Raises if it is missing — never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA library is required (no CPU fallback). "
                "Build it with `python -m lanegcn_b200.build`."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError if the symbol is missing
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"lgcn {what} failed ({rc}): {lib().lgcn_last_error().decode()}")


def ptr(t):
    """Device (or host) address of a torch tensor / None."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch

    return torch.cuda.current_stream().cuda_stream
