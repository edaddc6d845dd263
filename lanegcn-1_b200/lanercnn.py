"""LaneRCNN clients of the same kernels (reference ``lanercnn.py``; BASELINE config 5).

Built so far: the lane-graph layers — ``LaneRoI`` (lanercnn.py:354-430: Linear+GN+ReLU, then the 4-block LaneConv
loop with per-key ``len > 0`` guards) and ``GlobalGraphNet`` (lanercnn.py:517-600: the M2M loop) — as drop-ins with
the reference's parameter names, running on ``lgcn_linear128`` + ``lgcn_laneconv_stack``.  Not built yet (listed in
DESIGN.md §7): ``LaneInput``, ``LanePooling``, ``Interactor``, ``subgraph_gather``, ``Decode``.
"""
from __future__ import annotations

from typing import Dict

import torch
from torch import Tensor, nn

from . import _C
from .blocks import Linear
from .lanegcn import C_, _LaneConvStack, _f32c, _need_cuda, _packed_of


class LaneRoI(_LaneConvStack):
    """lanercnn.py:354-430.  Empty edge sets contribute nothing, which is what the ``len(...) > 0`` guards do."""

    def __init__(self, config, input_dim):
        super().__init__()
        if input_dim != C_ or config["n_map"] != C_:
            raise ValueError("lanegcn_b200: LaneRoI kernels are built for input_dim = n_map = 128")
        self.input = Linear(input_dim, config["n_map"], act=True)
        self._init_fuse(config)
        self.relu = nn.ReLU(inplace=True)

    @torch.no_grad()
    def forward(self, feat: Tensor, graph: Dict) -> Tensor:
        _need_cuda(feat, "feat")
        feat = _f32c(feat)
        out = torch.empty_like(feat)
        _C.check(_C.lib().lgcn_linear128(feat.data_ptr(), None, None, None, None, None, 1, None, 0,
                                         self.input.linear.weight.data_ptr(), 1, self.input.norm.weight.data_ptr(),
                                         self.input.norm.bias.data_ptr(), None, _C.EPI_GN | _C.EPI_RELU1,
                                         out.data_ptr(), C_, feat.shape[0], _C.stream_ptr()), "linear128(LaneRoI.input)")
        return self._stack(out, _packed_of(graph))


class GlobalGraphNet(_LaneConvStack):
    """lanercnn.py:517-600: identical to LaneGCN's M2M loop."""

    def __init__(self, config):
        super().__init__()
        self._init_fuse(config)
        self.relu = nn.ReLU(inplace=True)

    @torch.no_grad()
    def forward(self, feat: Tensor, graph: Dict) -> Tensor:
        if len(graph["feats"]) == 0 or len(graph["pre"][-1]["u"]) == 0 or len(graph["suc"][-1]["u"]) == 0:
            return (graph["feats"].new_zeros(0),)  # the reference's degenerate branch returns a 1-tuple (:553-563)
        _need_cuda(feat, "feat")
        return self._stack(_f32c(feat).clone(), _packed_of(graph))
