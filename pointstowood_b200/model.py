"""PointsToWood's network on the B200-native ops.

Same constructor / forward signatures, module tree and state-dict key names as
/root/reference/pointstowood/src/model.py, so a reference checkpoint
(`{'model_state_dict': ...}`, optional `module.` prefix) loads unchanged.  What differs is
the execution of the hot path (src/model.py:103-127, src/pointnet.py:116-132):

* voxel sub-sampling, radius / kNN search and the whole PointNetConv (gather -> per-edge MLP
  -> BatchNorm -> max) run as libp2w kernels on fixed-width int32 neighbour tables; the
  [E, C+4] / [E, H] / [E, C'] edge tensors are never materialised;
* the ReflectanceYesNo gate is the constant 1.0 (gumbel_softmax over a single logit, SURVEY.md
  Appendix C.1): its parameters are kept for checkpoint compatibility, its compute, RNG draw
  and two host syncs are skipped;
* the dense per-point blocks (stem, InvertedResidualBlock, FP MLPs, head) stay torch/cuBLAS
  modules evaluated in [N, C] layout (k=1 Conv1d == Linear), optionally under bf16 autocast.

Training mode (train.py, BASELINE.json configs[4]) keeps the reference's semantics -- random 50 %
sampling (src/model.py:97-101,114), batch-statistics BatchNorm over all E edges inside local_nn -- with
the graph (radius / kNN tables) from libp2w and the differentiable part in torch autograd on the
fixed-width neighbour table; the fused forward kernels are eval-mode only.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import ops

__all__ = ["Net", "SAModule", "GlobalSAModule", "FPModule", "InvertedResidualBlock", "PointNetConv", "MLP",
           "initialize_weights", "load_model", "randomise_bn_", "make_data"]


def initialize_weights(model: nn.Module) -> None:
    """Xavier-uniform Linear, Kaiming-uniform Conv1d, zero bias (src/model.py:9-16)."""
    for m in model.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
        elif isinstance(m, nn.Conv1d):
            nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
        else:
            continue
        if m.bias is not None:
            nn.init.zeros_(m.bias)


def MLP(channels) -> nn.Sequential:
    """Linear -> ReLU -> BatchNorm1d per layer, no BatchNorm after the first (src/model.py:198-202)."""
    layers = []
    for i, (cin, cout) in enumerate(zip(channels[:-1], channels[1:])):
        block = [nn.Linear(cin, cout), nn.ReLU()]
        if i > 0:
            block.append(nn.BatchNorm1d(cout))
        layers.append(nn.Sequential(*block))
    return nn.Sequential(*layers)


def _bn_affine(bn: nn.BatchNorm1d):
    """Eval-mode BatchNorm as y = x * scale + shift."""
    scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
    return scale, bn.bias - bn.running_mean * scale


def _bn(bn: nn.BatchNorm1d, x: Tensor) -> Tensor:
    """BatchNorm1d over the rows of [N, C]: running statistics in eval mode, batch statistics (and the
    running-statistics update) in training mode, as the module itself would do."""
    if bn.training:
        return bn(x)
    return F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps)


def _pointwise(conv: nn.Conv1d, x: Tensor) -> Tensor:
    """k=1 Conv1d applied to rows of [N, C]."""
    return F.linear(x, conv.weight.squeeze(-1), conv.bias)


class DepthwiseSeparableConv1d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0):
        super().__init__()
        self.depthwise_conv = nn.Conv1d(in_channels, in_channels, kernel_size, stride, padding, groups=in_channels)
        self.depthwise_bn = nn.BatchNorm1d(in_channels)
        self.pointwise_conv = nn.Conv1d(in_channels, out_channels, 1)
        self.pointwise_bn = nn.BatchNorm1d(in_channels)

    def forward(self, x: Tensor) -> Tensor:          # x: [N, C]
        x = x * self.depthwise_conv.weight.view(1, -1) + self.depthwise_conv.bias
        x = F.relu(_bn(self.depthwise_bn, x))
        return F.relu(_bn(self.pointwise_bn, _pointwise(self.pointwise_conv, x)))


class InvertedResidualBlock(nn.Module):
    """src/model.py:46-85 evaluated in [N, C] layout."""

    def __init__(self, in_channels, out_channels, expansion_factor=4):
        super().__init__()
        e = in_channels * expansion_factor
        self.expansion_factor = expansion_factor
        self.expand = nn.Sequential(nn.Conv1d(in_channels, e, 1), nn.BatchNorm1d(e), nn.ReLU())
        self.conv = nn.Sequential(DepthwiseSeparableConv1d(e, e), nn.BatchNorm1d(e), nn.ReLU(),
                                  DepthwiseSeparableConv1d(e, e), nn.BatchNorm1d(e))
        self.project = nn.Sequential(nn.Conv1d(e, out_channels, 1), nn.BatchNorm1d(out_channels))
        self.shortcut = nn.Sequential()
        if in_channels != out_channels:
            self.shortcut = nn.Sequential(nn.Conv1d(in_channels, out_channels, 1), nn.BatchNorm1d(out_channels))

    def forward(self, x: Tensor) -> Tensor:
        out = F.relu(_bn(self.expand[1], _pointwise(self.expand[0], x)))
        out = F.relu(_bn(self.conv[1], self.conv[0](out)))
        out = _bn(self.conv[4], self.conv[3](out))
        out = _bn(self.project[1], _pointwise(self.project[0], out))
        res = x if len(self.shortcut) == 0 else _bn(self.shortcut[1], _pointwise(self.shortcut[0], x))
        return F.relu(out + res)


class PointNetConv(nn.Module):
    """Drop-in for src/pointnet.py's PointNetConv(aggr='max') with a 2-layer local_nn.

    forward(x, (pos_src, pos_tgt), nbr): `nbr` is either a [N_tgt, K] int32 neighbour table
    (-1 padded; the fast path) or the reference's LongTensor edge_index [2, E] (row 0 = source j,
    row 1 = target i, targets ascending as knn/radius emit them)."""

    def __init__(self, local_nn=None, global_nn=None, add_self_loops=False, conv_mode=ops.CONV_FP32, **kwargs):
        super().__init__()
        if global_nn is not None or add_self_loops:
            raise NotImplementedError("the reference builds PointNetConv(global_nn=None, add_self_loops=False)")
        self.radius = kwargs.pop("radius", None)
        self.local_nn = local_nn
        self.global_nn = None
        self.add_self_loops = False
        self.conv_mode = conv_mode

    @staticmethod
    def edge_index_to_table(edge_index: Tensor, n_tgt: int, k: int = 32) -> Tensor:
        j, i = edge_index[0], edge_index[1]
        start = torch.searchsorted(i.contiguous(), torch.arange(n_tgt, device=i.device))
        slot = torch.arange(i.numel(), device=i.device) - start[i]
        table = torch.full((n_tgt, k), -1, device=i.device, dtype=torch.int32)
        table[i, slot] = j.to(torch.int32)
        return table

    def _forward_autograd(self, x: Tensor, pos_src: Tensor, pos_tgt: Tensor, nbr: Tensor) -> Tensor:
        """src/pointnet.py:116-132 + aggr='max' in plain torch on the VALID edges of the table (BatchNorm in
        local_nn sees exactly the E edges the reference sees); differentiable w.r.t. x and the weights."""
        valid = nbr >= 0
        i = torch.arange(nbr.size(0), device=nbr.device).view(-1, 1).expand_as(nbr)[valid]
        j = nbr[valid].long()
        rel = pos_src[j, :3] - pos_tgt[i, :3]
        dist = torch.norm(rel, dim=1)
        dtab = dist.new_full(nbr.shape, float("-inf"))
        dtab[valid] = dist
        maxd = dtab.max(dim=1).values                      # scatter_max(|rel|, i), src/pointnet.py:122
        msg = torch.cat([x[j], rel / (maxd[i].unsqueeze(1) + 1e-8), pos_src[j, 3:4]], dim=1)
        h = self.local_nn(msg)
        tab = h.new_full((nbr.size(0), nbr.size(1), h.size(1)), float("-inf"))
        tab[valid] = h
        out = tab.max(dim=1).values
        return torch.where(valid.any(dim=1, keepdim=True), out, torch.zeros_like(out))   # no edge -> 0 (PyG)

    def forward(self, x, pos, nbr: Tensor) -> Tensor:
        if isinstance(x, tuple):
            x = x[0]
        pos_src, pos_tgt = pos if isinstance(pos, tuple) else (pos, pos)
        if nbr.dtype == torch.int64 and nbr.dim() == 2 and nbr.size(0) == 2:
            nbr = self.edge_index_to_table(nbr, pos_tgt.size(0))
        if self.training or torch.is_grad_enabled() and x.requires_grad:
            return self._forward_autograd(x.float(), pos_src, pos_tgt, nbr)
        lin1, lin2, bn = self.local_nn[0][0], self.local_nn[1][0], self.local_nn[1][2]
        scale, shift = _bn_affine(bn)
        return ops.pointnet_conv_max(x.float(), pos_src, pos_tgt, nbr, lin1.weight, lin1.bias, lin2.weight, lin2.bias,
                                     scale, shift, self.conv_mode)


class ReflectanceYesNo(nn.Module):
    """Parameters kept for checkpoint compatibility; the reference's output is identically 1.0
    (hard gumbel-softmax over ONE logit, src/model.py:160,173-175), so forward returns ones."""

    def __init__(self, input_dim, hidden_dim, temperature=1.0):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, hidden_dim)
        self.fc3 = nn.Linear(hidden_dim, 1)
        self.temperature = temperature

    def forward(self, x: Tensor, batch: Tensor) -> Tensor:
        return torch.ones(x.size(0), device=x.device, dtype=torch.float32)


class SAModule(nn.Module):
    def __init__(self, resolution, radius, k, NN, RNN, conv_mode=ops.CONV_FP32):
        super().__init__()
        self.resolution, self.radius, self.k = resolution, radius, k
        self.conv = PointNetConv(local_nn=MLP(NN), global_nn=None, add_self_loops=False, radius=radius,
                                 conv_mode=conv_mode)
        self.residual_block = InvertedResidualBlock(RNN, RNN)
        self.reflectanceyesno = ReflectanceYesNo(input_dim=1, hidden_dim=32)

    def voxelsample(self, pos: Tensor, batch: Tensor, resolution: float) -> Tensor:
        return ops.voxel_sample(pos, resolution, batch)

    def random_sample(self, num_points: int, device) -> Tensor:
        """src/model.py:97-101: a sorted random half of the rows.  `self.sample_idx` (a tensor, consumed once)
        pins the draw for parity tests; `self.generator` seeds it otherwise."""
        pinned = getattr(self, "sample_idx", None)
        if pinned is not None:
            self.sample_idx = None
            return pinned.to(device)
        gen = getattr(self, "generator", None)
        perm = torch.randperm(num_points, generator=gen, device=gen.device if gen is not None else "cpu")
        return torch.sort(perm[: int(num_points * 0.5)]).values.to(device)

    def forward(self, x, pos, batch, reflectance, sf):
        B = sf.numel()
        pos = pos[:, :3].contiguous()
        ptr = ops.batch_to_ptr(batch, B)
        if self.training:
            idx = self.random_sample(pos.size(0), pos.device)
        else:
            idx = self.voxelsample(pos, batch, self.resolution)
        batch_t = batch[idx]
        ptr_t = ops.batch_to_ptr(batch_t, B)
        pos_t = pos[idx]
        if self.resolution == 0.04:
            nbr, _ = ops.radius_table(pos, pos_t, self.resolution * 2, ptr, ptr_t, self.k)
        else:
            nbr = ops.knn_table(pos, pos_t, self.k, ptr, ptr_t)
        pos4, pos_back = ops.sa_prepare(pos, reflectance, ptr, sf)
        x = self.conv(x, (pos4, pos4[idx]), nbr)
        x = self.residual_block(x)
        return x, pos_back[idx], batch_t, reflectance[idx], sf


class GlobalSAModule(nn.Module):
    def __init__(self, NN):
        super().__init__()
        self.NN = MLP(NN)

    def forward(self, x, pos, batch, reflectance, sf):
        B = sf.numel()
        x = self.NN(torch.cat([x, pos], dim=1))
        if x.requires_grad:          # global_max_pool with autograd (src/model.py:136)
            x = torch.zeros((B, x.size(1)), device=x.device, dtype=x.dtype).scatter_reduce(
                0, batch.view(-1, 1).expand(-1, x.size(1)), x, "amax", include_self=False)
        else:
            x = ops.global_max_pool(x.float(), batch, ptr=ops.batch_to_ptr(batch, B))
        pos = pos.new_zeros((B, 3))
        batch = torch.arange(B, device=batch.device)
        return x, pos, batch, reflectance.new_zeros(B), sf


class FPModule(nn.Module):
    def __init__(self, k, NN):
        super().__init__()
        self.k = k
        self.NN = MLP(NN)

    def forward(self, x, pos, batch, x_skip, pos_skip, batch_skip, num_tiles: Optional[int] = None):
        if num_tiles is None:
            num_tiles = int(batch_skip[-1].item()) + 1
        ptr_x, ptr_y = ops.batch_to_ptr(batch, num_tiles), ops.batch_to_ptr(batch_skip, num_tiles)
        if x.requires_grad:          # knn_interpolate with autograd (src/model.py:149): the graph from libp2w
            nbr, d2 = ops.knn_table(pos.contiguous(), pos_skip.contiguous(), self.k, ptr_x, ptr_y, return_d2=True)
            valid = nbr >= 0
            # d2 as upstream computes it for the weights: plain sum of squares of the coordinate differences
            diff = pos[nbr.clamp(min=0).long()] - pos_skip.unsqueeze(1)
            w = torch.where(valid, 1.0 / torch.clamp((diff * diff).sum(-1), min=1e-16), torch.zeros_like(d2))
            y = (x[nbr.clamp(min=0).long()] * w.unsqueeze(-1)).sum(1) / w.sum(1, keepdim=True)
            y = y if x_skip is None else torch.cat([y, x_skip], dim=1)
            return self.NN(y), pos_skip, batch_skip
        c = x.size(1)
        cs = 0 if x_skip is None else x_skip.size(1)
        buf = torch.empty((pos_skip.size(0), c + cs), device=x.device, dtype=torch.float32)
        ops.knn_interpolate(x.float(), pos, pos_skip, k=self.k, ptr_x=ptr_x, ptr_y=ptr_y, out=buf)
        if x_skip is not None:
            buf[:, c:] = x_skip
        return self.NN(buf), pos_skip, batch_skip


class Net(nn.Module):
    def __init__(self, num_classes, C=32, conv_mode=ops.CONV_FP32):
        super().__init__()
        self.stem_mlp = MLP([3, C])
        self.sa1_module = SAModule(0.04, 0.04, 32, [C + 4, C * 2, C * 4], C * 4, conv_mode)
        self.sa2_module = SAModule(0.08, 0.08, 32, [C * 4 + 4, C * 6, C * 8], C * 8, conv_mode)
        self.sa3_module = SAModule(0.16, 0.16, 32, [C * 8 + 4, C * 12, C * 16], C * 16, conv_mode)
        self.sa4_module = GlobalSAModule([C * 16 + 3, C * 16, C * 16])
        self.fp4_module = FPModule(2, [C * 32, C * 24, C * 16])
        self.fp3_module = FPModule(2, [C * 24, C * 20, C * 16])
        self.fp2_module = FPModule(2, [C * 20, C * 16, C * 16])
        self.fp1_module = FPModule(2, [C * 17, C * 16, C * 16])
        self.conv1 = nn.Conv1d(C * 16, C * 16, 1)
        self.conv2 = nn.Conv1d(C * 16, num_classes, 1)
        self.norm = nn.BatchNorm1d(C * 16)
        initialize_weights(self)
        self.conv_mode = conv_mode
        self.inference_dtype = torch.float32
        self.use_engine = True          # eval mode runs through engine.InferenceEngine (folded BN, super-batches)
        self._engine = None

    def set_conv_mode(self, mode: int) -> "Net":
        for m in self.modules():
            if isinstance(m, PointNetConv):
                m.conv_mode = mode
        self.conv_mode = mode
        return self

    def set_precision(self, precision: str) -> "Net":
        """'fp32': FP32 everywhere (parity mode, |dp| <= 1e-3).  'bf16': bf16 activations / weights with FP32
        accumulation, PointNetConv on tcgen05 (|dp| <= 1e-2).  'bf16-conv': only the PointNetConv in bf16."""
        if precision not in ("fp32", "bf16", "bf16-conv"):
            raise ValueError(precision)
        self.set_conv_mode(ops.CONV_FP32 if precision == "fp32" else ops.CONV_BF16_TC)
        self.inference_dtype = torch.bfloat16 if precision == "bf16" else torch.float32
        return self

    def engine(self):
        """The eval-mode executor for the current weights / precision (re-folded when either changes)."""
        from .engine import InferenceEngine
        tensors = getattr(self, "_engine_tensors", None)
        if tensors is None:          # the module tree is fixed: walk it once (the walk alone cost ~0.5 ms per forward)
            tensors = self._engine_tensors = list(self.parameters()) + list(self.buffers())
        key = (self.inference_dtype, self.conv_mode, tensors[0].device, sum(t._version for t in tensors),
               tensors[0].data_ptr(), tensors[-1].data_ptr())
        if self._engine is None or self._engine[0] != key:
            self._engine = (key, InferenceEngine(self, self.inference_dtype, self.conv_mode))
        return self._engine[1]

    def forward(self, data):
        if not self.training and self.use_engine:
            return self.engine()(data.pos, data.reflectance, data.batch, data.sf, getattr(data, "ptr", None),
                                 getattr(data, "group_ptr", None))
        B = data.sf.numel()
        pos = data.pos[:, :3].contiguous()
        data.x = self.stem_mlp(pos)
        sa0 = (data.x, pos, data.batch, data.reflectance, data.sf)
        sa1 = self.sa1_module(*sa0)
        sa2 = self.sa2_module(*sa1)
        sa3 = self.sa3_module(*sa2)
        sa4 = self.sa4_module(*sa3)
        fp4 = self.fp4_module(*sa4[:3], *sa3[:3], num_tiles=B)
        fp3 = self.fp3_module(*fp4, *sa2[:3], num_tiles=B)
        fp2 = self.fp2_module(*fp3, *sa1[:3], num_tiles=B)
        x, _, _ = self.fp1_module(*fp2, *sa0[:3], num_tiles=B)
        x = F.relu(_bn(self.norm, _pointwise(self.conv1, x)))
        x = _pointwise(self.conv2, x)
        return torch.squeeze(x).to(torch.float)


def load_model(path: str, model: nn.Module, device) -> nn.Module:
    """src/predicter.py:97-105: strips a DataParallel `module.` prefix, strict=False."""
    checkpoint = torch.load(path, map_location=device)
    state = {(k[7:] if k.startswith("module.") else k): v for k, v in checkpoint["model_state_dict"].items()}
    model.load_state_dict(state, strict=False)
    return model


def randomise_bn_(model: nn.Module, seed: int = 5) -> nn.Module:
    """Seeded, non-trivial BatchNorm statistics for random-weight benchmarks (the shipped
    checkpoint is absent, so eval-mode BN would otherwise be the identity)."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm1d):
            c = m.num_features
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(c, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(c, generator=g) * 0.5 + 0.75)
                m.weight.copy_(torch.rand(c, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(c, generator=g) * 0.1)
    return model


def make_data(pos: Tensor, reflectance: Tensor, batch: Tensor, sf: Tensor, **extra) -> SimpleNamespace:
    """Minimal stand-in for the PyG Batch the reference hands to Net.forward."""
    return SimpleNamespace(pos=pos, reflectance=reflectance, batch=batch, sf=sf, **extra)
