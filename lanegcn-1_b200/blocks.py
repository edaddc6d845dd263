"""Stock-PyTorch building blocks for the two stages that stay OFF the hot path (ActorNet, PredNet).

Written fresh; only the parameter names/shapes follow the reference (layers.py:40-62 Conv1d, :65-87 Linear,
:142-190 Res1d, :193-238 LinearRes; lanegcn.py:212-263 ActorNet, :575-631 PredNet, :713-737 AttDest) so that
reference checkpoints load by key.  All norms are GroupNorm with ONE group (gcd(ng=1, C) = 1).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


class _GN1(nn.GroupNorm):
    """GroupNorm with ONE group = LayerNorm over all non-batch dims with a per-channel affine.  On the GPU it is
    spelled that way: ATen's GroupNorm path (RowwiseMoments + fused-params + elementwise kernels) was 0.8 ms of
    ActorNet's 1.9 ms per batch-128 forward; the vectorised LayerNorm kernel does the same reduction in one launch.
    Same parameters (``weight``, ``bias``), same state_dict keys."""

    def forward(self, x):
        if not x.is_cuda:
            return super().forward(x)
        if x.dim() == 2:
            return F.layer_norm(x, x.shape[1:], self.weight, self.bias, self.eps)
        y = F.layer_norm(x, x.shape[1:], None, None, self.eps)
        shape = (1, -1) + (1,) * (x.dim() - 2)
        return torch.addcmul(self.bias.view(shape), y, self.weight.view(shape))


def _gn(c: int) -> nn.GroupNorm:
    return _GN1(1, c)


class Linear(nn.Module):
    """bias-free Linear -> GN(1) -> optional ReLU.  Parameter holder for the hot-path layers too."""

    def __init__(self, n_in, n_out, act=True):
        super().__init__()
        self.linear = nn.Linear(n_in, n_out, bias=False)
        self.norm = _gn(n_out)
        self.act = act

    def forward(self, x):
        y = self.norm(self.linear(x))
        return F.relu(y) if self.act else y


class Conv1d(nn.Module):
    def __init__(self, n_in, n_out, kernel_size=3, stride=1, act=True):
        super().__init__()
        self.conv = nn.Conv1d(n_in, n_out, kernel_size, stride, (kernel_size - 1) // 2, bias=False)
        self.norm = _gn(n_out)
        self.act = act

    def forward(self, x):
        y = self.norm(self.conv(x))
        return F.relu(y) if self.act else y


class Res1d(nn.Module):
    def __init__(self, n_in, n_out, stride=1):
        super().__init__()
        self.conv1 = nn.Conv1d(n_in, n_out, 3, stride, 1, bias=False)
        self.conv2 = nn.Conv1d(n_out, n_out, 3, 1, 1, bias=False)
        self.bn1, self.bn2 = _gn(n_out), _gn(n_out)
        self.downsample = None
        if stride != 1 or n_in != n_out:
            self.downsample = nn.Sequential(nn.Conv1d(n_in, n_out, 1, stride, bias=False), _gn(n_out))

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return F.relu(y + (x if self.downsample is None else self.downsample(x)))


class LinearRes(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.linear1 = nn.Linear(n, n, bias=False)
        self.linear2 = nn.Linear(n, n, bias=False)
        self.norm1, self.norm2 = _gn(n), _gn(n)

    def forward(self, x):
        y = F.relu(self.norm1(self.linear1(x)))
        return F.relu(self.norm2(self.linear2(y)) + x)


def _upsample2_linear(x):
    """F.interpolate(x, scale_factor=2, mode="linear", align_corners=False) on [A,C,L] spelled with slices:
    out[2i] = .25 x[i-1] + .75 x[i], out[2i+1] = .75 x[i] + .25 x[i+1], edges clamped.  (ATen's
    upsample_linear1d kernel takes ~65 ms on B200 for [2560,128,10]; this is a handful of tiny launches.)"""
    left = torch.cat((x[..., :1], x[..., :-1]), -1)
    right = torch.cat((x[..., 1:], x[..., -1:]), -1)
    return torch.stack((0.25 * left + 0.75 * x, 0.75 * x + 0.25 * right), -1).flatten(-2)


class ActorNet(nn.Module):
    """1-D conv FPN over the 20 history steps -> [sum A, n_actor] (lanegcn.py:212-263).  Off the hot path."""

    def __init__(self, config):
        super().__init__()
        widths, n_in, groups = [32, 64, 128], 3, []
        for i, w in enumerate(widths):
            groups.append(nn.Sequential(Res1d(n_in, w, stride=1 if i == 0 else 2), Res1d(w, w)))
            n_in = w
        self.groups = nn.ModuleList(groups)
        n = config["n_actor"]
        self.lateral = nn.ModuleList([Conv1d(w, n, act=False) for w in widths])
        self.output = Res1d(n, n)

    def forward(self, actors):
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):  # keep the convs in true fp32
            feats, x = [], actors
            for g in self.groups:
                x = g(x)
                feats.append(x)
            x = self.lateral[-1](feats[-1])
            for i in (1, 0):
                x = _upsample2_linear(x)
                x = x + self.lateral[i](feats[i])
            return self.output(x)[:, :, -1]


class AttDest(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.dist = nn.Sequential(nn.Linear(2, n), nn.ReLU(inplace=True), Linear(n, n))
        self.agt = Linear(2 * n, n)

    def forward(self, agts, agt_ctrs, dest_ctrs):
        m = dest_ctrs.size(1)
        d = self.dist((agt_ctrs.unsqueeze(1) - dest_ctrs).reshape(-1, 2))
        a = agts.unsqueeze(1).expand(-1, m, -1).reshape(-1, agts.size(1))
        return self.agt(torch.cat((d, a), 1))


class PredNet(nn.Module):
    """K regression heads + destination attention + score sort (lanegcn.py:575-631).  Off the hot path.
    Same outputs as the reference; the per-scene python loops are replaced by batched ops + views."""

    def __init__(self, config):
        super().__init__()
        n, self.num_mods = config["n_actor"], config["num_mods"]
        self.pred = nn.ModuleList(
            [nn.Sequential(LinearRes(n), nn.Linear(n, 2 * config["num_preds"])) for _ in range(self.num_mods)]
        )
        self.att_dest = AttDest(n)
        self.cls = nn.Sequential(LinearRes(n), nn.Linear(n, 1))

    def forward(self, actors, actor_idcs, actor_ctrs):
        ctrs = actor_ctrs.cat if hasattr(actor_ctrs, "cat") else torch.cat(list(actor_ctrs), 0)
        cls, reg = self.core(actors, ctrs)
        sizes = [len(i) for i in actor_idcs]
        return {"cls": list(torch.split(cls, sizes)), "reg": list(torch.split(reg, sizes))}

    def core(self, actors, ctrs):
        """Per-actor part (no host-side sizes): actors [A,n], ctrs [A,2] -> cls [A,K], reg [A,K,T,2]."""
        reg = torch.stack([head(actors) for head in self.pred], 1)
        reg = reg.view(reg.size(0), reg.size(1), -1, 2) + ctrs.view(-1, 1, 1, 2)
        feats = self.att_dest(actors, ctrs, reg[:, :, -1].detach())
        cls = self.cls(feats).view(-1, self.num_mods)
        cls, order = cls.sort(1, descending=True)
        reg = torch.gather(reg, 1, order.view(-1, self.num_mods, 1, 1).expand(-1, -1, reg.size(2), 2))
        return cls, reg
