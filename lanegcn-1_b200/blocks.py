"""Parameter holders of ActorNet / PredNet (+ the small layer types the hot-path modules reuse).

Only the parameter names/shapes follow the reference (layers.py:40-62 Conv1d, :65-87 Linear, :142-190 Res1d,
:193-238 LinearRes; lanegcn.py:212-263 ActorNet, :575-631 PredNet, :713-737 AttDest) so that reference checkpoints
load by key.  All norms are GroupNorm with ONE group (gcd(ng=1, C) = 1).  On a CUDA device ``ActorNet.forward`` and
``PredNet.core`` are ONE kernel each (csrc/actor_net.cu, csrc/pred_net.cu through the C ABI); the stock-PyTorch
spelling written below (``forward_torch`` / ``core_torch``) is kept as the fp32 reference the tests compare the
kernels with (LGCN_TORCH_BLOCKS=1 selects it at run time).
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.nn.functional as F
from torch import nn

from . import _C


def _use_torch_blocks() -> bool:
    return os.environ.get("LGCN_TORCH_BLOCKS", "0") == "1"


class _ParamPack:
    """Device weight pack written by a C packing routine from a fixed list of parameters; refreshed IN PLACE when a
    parameter changes (data pointer or version counter), so addresses captured in CUDA graphs stay valid."""

    def __init__(self):
        self.key, self.buf, self.version = None, None, 0

    def get(self, params, n_floats: int, write) -> torch.Tensor:
        key = tuple((p.data_ptr(), p._version) for p in params)
        if key != self.key:
            dev = params[0].device
            if self.buf is None or self.buf.device != dev or self.buf.numel() != n_floats:
                self.buf = torch.empty(n_floats, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                write([p.detach() if (p.dtype == torch.float32 and p.is_contiguous()) else p.detach().float().contiguous()
                       for p in params], self.buf)
            self.key = key
            self.version += 1
        return self.buf

    def invalidate(self):
        self.key = None


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
    """1-D conv FPN over the 20 history steps -> [sum A, n_actor] (lanegcn.py:212-263)."""

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
        self._pp = _ParamPack()

    def _layers(self):
        """(conv, norm) of the 20 layers in the order lgcn_actor_net_pack expects."""
        out = []
        for g in self.groups:
            out += [(g[0].conv1, g[0].bn1), (g[0].conv2, g[0].bn2), (g[0].downsample[0], g[0].downsample[1]),
                    (g[1].conv1, g[1].bn1), (g[1].conv2, g[1].bn2)]
        out += [(lat.conv, lat.norm) for lat in self.lateral]
        return out + [(self.output.conv1, self.output.bn1), (self.output.conv2, self.output.bn2)]

    def wpack(self) -> torch.Tensor:
        lib = _C.lib()
        layers = self._layers()
        params = [c.weight for c, _ in layers] + [n.weight for _, n in layers] + [n.bias for _, n in layers]

        def write(ps, buf):
            arr = ctypes.c_void_p * 20
            w, g, b = (arr(*[t.data_ptr() for t in ps[i * 20:(i + 1) * 20]]) for i in range(3))
            _C.check(lib.lgcn_actor_net_pack(w, g, b, buf.data_ptr(), _C.stream_ptr()), "actor_net_pack")
            self._keep = ps   # converted copies must outlive the asynchronous packing kernels
        return self._pp.get(params, lib.lgcn_actor_net_wpack_floats(), write)

    @staticmethod
    def tensor_core_path() -> bool:
        """The output Res1d on the tensor core (lgcn_actor_net_tc) unless LGCN_ACTOR=fp32 or the SIMT engine is selected."""
        return os.environ.get("LGCN_ACTOR", "tc") != "fp32" and _C.lib().lgcn_get_gemm_engine() == 1

    def forward_ntc(self, feats, out=None, n_dev=None, ws=None):
        """feats [A, 20, 3] (step-major, as the dataset stores the histories) -> [A, n_actor].  One fp32 kernel, or (default
        on the tcgen05 engine) the fp32 kernel up to the feature pyramid + the output Res1d as two tensor-core Linears;
        ``ws``: workspace of lgcn_actor_net_tc_workspace_bytes(A) (static buffers of a graph-captured slot)."""
        if self.output.conv1.weight.shape[0] != 128 or feats.shape[1:] != (20, 3):
            raise RuntimeError("lanegcn_b200: the ActorNet kernel is built for n_actor = 128 and [A, 20, 3] inputs")
        feats = feats if (feats.dtype == torch.float32 and feats.is_contiguous()) else feats.float().contiguous()
        lib, n = _C.lib(), feats.shape[0]
        if out is None:
            out = torch.empty(n, 128, dtype=torch.float32, device=feats.device)
        with torch.cuda.device(feats.device):
            if self.tensor_core_path():
                if ws is None:
                    ws = torch.empty(lib.lgcn_actor_net_tc_workspace_bytes(n), dtype=torch.uint8, device=feats.device)
                _C.check(lib.lgcn_actor_net_tc(feats.data_ptr(), self.wpack().data_ptr(), out.data_ptr(), n, _C.ptr(n_dev),
                                               ws.data_ptr(), _C.stream_ptr()), "actor_net_tc")
            else:
                _C.check(lib.lgcn_actor_net(feats.data_ptr(), self.wpack().data_ptr(), out.data_ptr(), n, _C.ptr(n_dev),
                                            _C.stream_ptr()), "actor_net")
        return out

    def forward(self, actors):
        """actors [A, 3, 20] (what actor_gather returns, lanegcn.py:155-168) -> [A, n_actor]."""
        if not actors.is_cuda or _use_torch_blocks():
            return self.forward_torch(actors)
        with torch.no_grad():
            return self.forward_ntc(actors.transpose(1, 2))

    def forward_torch(self, actors):
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
    """K regression heads + destination attention + score sort (lanegcn.py:575-631).  Same outputs as the reference;
    on a CUDA device the whole module (and, optionally, the world transform of lanegcn.py:145-150) is ONE kernel."""

    def __init__(self, config):
        super().__init__()
        n, self.num_mods = config["n_actor"], config["num_mods"]
        self.num_preds = config["num_preds"]
        self.pred = nn.ModuleList(
            [nn.Sequential(LinearRes(n), nn.Linear(n, 2 * config["num_preds"])) for _ in range(self.num_mods)]
        )
        self.att_dest = AttDest(n)
        self.cls = nn.Sequential(LinearRes(n), nn.Linear(n, 1))
        self._pp = _ParamPack()

    def forward(self, actors, actor_idcs, actor_ctrs):
        ctrs = actor_ctrs.cat if hasattr(actor_ctrs, "cat") else torch.cat(list(actor_ctrs), 0)
        cls, reg = self.core(actors, ctrs)
        sizes = [len(i) for i in actor_idcs]
        return {"cls": list(torch.split(cls, sizes)), "reg": list(torch.split(reg, sizes))}

    def wpack(self) -> torch.Tensor:
        lib = _C.lib()
        params = []
        for head in list(self.pred):
            lr = head[0]
            params += [lr.linear1.weight, lr.linear2.weight, lr.norm1.weight, lr.norm1.bias, lr.norm2.weight, lr.norm2.bias,
                       head[1].weight, head[1].bias]
        ad = self.att_dest
        params += [ad.dist[0].weight, ad.dist[0].bias, ad.dist[2].linear.weight, ad.dist[2].norm.weight, ad.dist[2].norm.bias,
                   ad.agt.linear.weight, ad.agt.norm.weight, ad.agt.norm.bias]
        lr = self.cls[0]
        params += [lr.linear1.weight, lr.linear2.weight, lr.norm1.weight, lr.norm1.bias, lr.norm2.weight, lr.norm2.bias,
                   self.cls[1].weight, self.cls[1].bias]

        def write(ps, buf):
            arr = (ctypes.c_void_p * len(ps))(*[t.data_ptr() for t in ps])
            _C.check(lib.lgcn_pred_net_pack(arr, buf.data_ptr(), _C.stream_ptr()), "pred_net_pack")
            self._keep = ps
        return self._pp.get(params, lib.lgcn_pred_net_wpack_floats(), write)

    def core(self, actors, ctrs, actor_off=None, rot=None, orig=None, cls=None, reg=None, n_dev=None):
        """Per-actor part (no host-side sizes): actors [A,n], ctrs [A,2] -> cls [A,K], reg [A,K,T,2].  With
        ``actor_off`` (int32 [B+1] scene offsets), ``rot`` [B,2,2] and ``orig`` [B,2] the trajectories come out in world
        coordinates (lanegcn.py:145-150)."""
        if not actors.is_cuda or _use_torch_blocks():
            return self.core_torch(actors, ctrs)
        if actors.shape[1] != 128 or self.num_mods != 6 or self.num_preds != 30:
            raise RuntimeError("lanegcn_b200: the PredNet kernel is built for n_actor = 128, num_mods = 6, num_preds = 30")
        f32c = lambda t: t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()  # noqa: E731
        actors, ctrs = f32c(actors), f32c(ctrs)
        A, dev = actors.shape[0], actors.device
        if cls is None:
            cls = torch.empty(A, self.num_mods, dtype=torch.float32, device=dev)
        if reg is None:
            reg = torch.empty(A, self.num_mods, self.num_preds, 2, dtype=torch.float32, device=dev)
        n_scenes = 1 if actor_off is None else actor_off.numel() - 1
        with torch.no_grad(), torch.cuda.device(dev):
            _C.check(_C.lib().lgcn_pred_net(actors.data_ptr(), ctrs.data_ptr(), _C.ptr(actor_off), n_scenes, _C.ptr(rot),
                                            _C.ptr(orig), self.wpack().data_ptr(), cls.data_ptr(), reg.data_ptr(), A,
                                            _C.ptr(n_dev), _C.stream_ptr()), "pred_net")
        return cls, reg

    def core_torch(self, actors, ctrs):
        reg = torch.stack([head(actors) for head in self.pred], 1)
        reg = reg.view(reg.size(0), reg.size(1), -1, 2) + ctrs.view(-1, 1, 1, 2)
        feats = self.att_dest(actors, ctrs, reg[:, :, -1].detach())
        cls = self.cls(feats).view(-1, self.num_mods)
        cls, order = cls.sort(1, descending=True)
        reg = torch.gather(reg, 1, order.view(-1, self.num_mods, 1, 1).expand(-1, -1, reg.size(2), 2))
        return cls, reg
