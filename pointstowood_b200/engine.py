"""Eval-mode executor for pointstowood_b200.model.Net: the same arithmetic as Net.forward
(/root/reference/pointstowood/src/model.py:226-245) re-scheduled for B200.

* SUPER-BATCHES.  The reference runs `batch_size` (8) tiles per forward (src/predicter.py:177-199):
  ~13 k points per launch set on a TLS plot, which leaves a B200 launch-bound.  Every op on the
  path is per tile or per point except the voxel sub-sampling, whose grid origin is the min over one
  batch.  `group_ptr` keeps that origin per reference batch, so any number
  of batches travels through one set of launches with unchanged results.
* FOLDED BATCHNORM.  In eval mode a BatchNorm1d after a k=1 convolution / Linear folds into its
  weights; `Linear -> ReLU -> BN` (MLP, src/model.py:198-202) leaves a per-channel affine AFTER the
  ReLU, which is folded FORWARD into the next Linear -- through knn_interpolate too, whose weights
  sum to one (src/model.py:149).  What cannot fold (depthwise -> BN -> ReLU chains inside
  InvertedResidualBlock, src/model.py:18-85) runs as one fused streaming kernel (p2w_affine_relu).
  Folding is done in FP64 once per (weights, dtype).
* DTYPE.  fp32 (parity mode, 1e-3) or bf16 activations / weights with FP32 accumulation (1e-2).
  Dense GEMMs are library calls (cuBLASLt bias+ReLU epilogues through torch), as SURVEY.md §2 row 8
  scopes them; the hot path -- sampling, radius / kNN, fused PointNetConv, interpolation, pooling
  -- is libp2w.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor, nn

from . import ops

__all__ = ["InferenceEngine"]


def _affine(bn: nn.BatchNorm1d):
    s = bn.weight.double() * torch.rsqrt(bn.running_var.double() + bn.eps)
    return s, bn.bias.double() - bn.running_mean.double() * s


class _Lin:
    """y = act(x @ Wt + b) with everything foldable already inside Wt / b."""

    def __init__(self, w: Tensor, b: Tensor, dtype, relu: bool, s_in=None, t_in=None, s_out=None, t_out=None,
                 pad_in: int = 0):
        w, b = w.double(), b.double()
        if w.dim() == 3:
            w = w.squeeze(-1)
        if s_in is not None:                     # input was (x*s_in + t_in) on its leading len(s_in) columns
            k = s_in.numel()
            b = b + w[:, :k] @ t_in
            w = torch.cat([w[:, :k] * s_in[None, :], w[:, k:]], dim=1)
        if s_out is not None:
            w = w * s_out[:, None]
            b = b * s_out + t_out
        if pad_in:
            w = torch.cat([w, w.new_zeros(w.size(0), pad_in)], dim=1)
        self.wt = w.t().contiguous().to(dtype)
        self.b = b.to(dtype).contiguous()
        self.b32 = b.float().contiguous()
        self.relu = relu

    def __call__(self, x: Tensor) -> Tensor:
        if self.relu:
            return torch._addmm_activation(self.b, x, self.wt, use_gelu=False)
        return torch.addmm(self.b, x, self.wt)


class _Residual:
    """InvertedResidualBlock (src/model.py:46-85), shortcut = identity (in == out channels)."""

    def __init__(self, blk, dtype):
        assert len(blk.shortcut) == 0, "the reference builds InvertedResidualBlock(C, C)"
        f32 = lambda t: t.float().contiguous()
        self.expand = _Lin(blk.expand[0].weight, blk.expand[0].bias, dtype, True, None, None, *_affine(blk.expand[1]))
        ds0, ds3 = blk.conv[0], blk.conv[3]

        def dw(ds):        # depthwise(k=1) -> BN : one per-channel affine
            s, t = _affine(ds.depthwise_bn)
            return ds.depthwise_conv.weight.double().view(-1) * s, ds.depthwise_conv.bias.double() * s + t

        a0, c0 = dw(ds0)
        self.a0, self.c0 = f32(a0), f32(c0)
        # bf16: the expand GEMM, its ReLU and this affine + ReLU in one tcgen05 kernel (csrc/dense_tc.cu); C = 512 stays on the
        # library GEMM + affine pass (0.33 ms fused against 0.35 ms: its weights stream through a 3-slice ring)
        c_in = self.expand.wt.size(0)
        self.fused_expand = dtype == torch.bfloat16 and ops.DENSE_TC and c_in % 64 == 0 and 64 <= c_in <= 256
        if self.fused_expand:
            self.w_expand = self.expand.wt.t().float().contiguous()
            self.ws_expand, self.ws_packed = None, False
        self.pw0 = _Lin(ds0.pointwise_conv.weight, ds0.pointwise_conv.bias, dtype, True, None, None,
                        *_affine(ds0.pointwise_bn))
        s1, t1 = _affine(blk.conv[1])
        a3, c3 = dw(ds3)
        self.s1, self.t1, self.a3, self.c3 = f32(s1), f32(t1), f32(a3), f32(c3)
        self.pw3 = _Lin(ds3.pointwise_conv.weight, ds3.pointwise_conv.bias, dtype, True, None, None,
                        *_affine(ds3.pointwise_bn))
        s4, t4 = _affine(blk.conv[4])
        self.project = _Lin(blk.project[0].weight, blk.project[0].bias, dtype, False, s4, t4, *_affine(blk.project[1]))

    def __call__(self, x: Tensor) -> Tensor:
        if self.fused_expand and x.is_cuda:
            if self.ws_expand is None or self.ws_expand.device != x.device:
                self.ws_expand, self.ws_packed = ops.dense_expand_ws(x.size(1), self.w_expand.size(0), x.device), False
            h = ops.dense_expand(x, self.w_expand, self.expand.b32, self.a0, self.c0, self.ws_expand, self.ws_packed)
            self.ws_packed = True
        else:
            h = ops.affine_relu_(self.expand(x), self.a0, self.c0)
        h = self.pw0(h)
        h = self.pw3(ops.affine_relu_(h, self.s1, self.t1, self.a3, self.c3))
        h = self.project(h)
        if h.numel() % 8 == 0 and h.is_contiguous() and x.is_contiguous() and h.dtype == x.dtype:
            return ops.add_relu_(h, x)
        return torch.relu_(h.add_(x))


class InferenceEngine:
    def __init__(self, net: nn.Module, dtype=torch.float32, conv_mode: Optional[int] = None):
        assert dtype in (torch.float32, torch.bfloat16)
        self.dtype = dtype
        self.conv_mode = conv_mode if conv_mode is not None else (
            ops.CONV_BF16_TC if dtype == torch.bfloat16 else ops.CONV_FP32)
        if dtype == torch.bfloat16 and self.conv_mode != ops.CONV_BF16_TC:
            raise ValueError("bf16 activations need the tensor-core PointNetConv")
        self.fold(net)

    # ------------------------------------------------------------------ folding (once per weights)
    @torch.no_grad()
    def fold(self, net: nn.Module) -> None:
        dt = self.dtype
        dev = next(net.parameters()).device
        f32 = lambda t: t.float().contiguous()
        stem = net.stem_mlp[0][0]
        self.stem = _Lin(stem.weight, stem.bias, torch.float32, True)
        self.sa = []
        for mod in (net.sa1_module, net.sa2_module, net.sa3_module):
            lin1, lin2, bn = mod.conv.local_nn[0][0], mod.conv.local_nn[1][0], mod.conv.local_nn[1][2]
            s, t = _affine(bn)
            w = [f32(lin1.weight), f32(lin1.bias), f32(lin2.weight), f32(lin2.bias), f32(s), f32(t)]
            H, K1 = lin1.weight.shape
            self.sa.append(dict(res=mod.resolution, k=mod.k, w=w, packed=False,
                                ws=ops.pointnet_conv_ws(K1 - 4, H, lin2.weight.size(0), self.conv_mode, dev),
                                residual=_Residual(mod.residual_block, dt)))
        g = net.sa4_module.NN
        k_in = g[0][0].weight.size(1)                                  # 515 = 512 features + xyz
        self.g_pad = (-k_in) % 8
        self.g1 = _Lin(g[0][0].weight, g[0][0].bias, dt, True, pad_in=self.g_pad)
        self.g2 = _Lin(g[1][0].weight, g[1][0].bias, dt, True)
        gs, gt = _affine(g[1][2])
        self.g_s, self.g_t = f32(gs), f32(gt)
        # FP modules (src/model.py:142-153).  The trailing BN affine of each MLP is folded into the consumer of its
        # output, and the FIRST Linear is split over its two inputs: relu([interp(x), x_skip] W^T + b) =
        # relu(interp(x Wc^T) + (x_skip Ws^T + b)) -- the interpolation is linear and its weights sum to one, so the
        # coarse part of the GEMM runs over the coarse rows (2.3-4x fewer) and the [n, C + C_skip] concatenation is
        # never built (p2w_knn_interpolate_add).
        self.fp = []
        pend = None                                                    # (s, t) owed by the previous MLP's output
        c_coarse = g[1][0].weight.size(0)                              # channels of the interpolated features
        for mod in (net.fp4_module, net.fp3_module, net.fp2_module, net.fp1_module):
            nn_ = mod.NN
            w1, b1 = nn_[0][0].weight, nn_[0][0].bias
            lc = _Lin(w1[:, :c_coarse], torch.zeros_like(b1), torch.float64, False, *(pend if pend else (None, None)))
            ls = _Lin(w1[:, c_coarse:], b1.double() + lc.b, dt, False)          # lc.b = Wc t_in: a constant survives the interpolation
            l2 = _Lin(nn_[1][0].weight, nn_[1][0].bias, dt, True)
            self.fp.append(dict(k=mod.k, wc=lc.wt.to(dt).contiguous(), ws=ls.wt, bs=ls.b, l2=l2))
            pend = _affine(nn_[1][2])
            c_coarse = nn_[1][0].weight.size(0)
        self.head1 = _Lin(net.conv1.weight, net.conv1.bias, dt, True, pend[0], pend[1], *_affine(net.norm))
        self.head2_w = net.conv2.weight.detach().reshape(-1).float().contiguous()       # one output channel: a row dot
        self.head2_b = float(net.conv2.bias.detach().reshape(-1)[0]) if net.conv2.weight.size(0) == 1 else None
        self.head2 = _Lin(net.conv2.weight, net.conv2.bias, dt, False)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def __call__(self, pos: Tensor, reflectance: Tensor, batch: Tensor, sf: Tensor, ptr: Optional[Tensor] = None,
                 group_ptr: Optional[Tensor] = None) -> Tensor:
        """pos [N,3] (tile-mean-shifted), reflectance [N], batch [N] int64 sorted, sf [T]; ptr [T+1] CSR of
        the tiles; group_ptr [G+1] CSR of the reference batches over tiles (None: one batch).
        Returns logits [N] fp32."""
        dt = self.dtype
        T = sf.numel()
        pos = pos[:, :3].contiguous()
        if ptr is None:
            ptr = ops.batch_to_ptr(batch, T)
        if group_ptr is None:
            group_ptr = torch.tensor([0, T], device=pos.device, dtype=torch.int64)
        x = self.stem(pos)                                             # [N0,32] fp32
        if dt != torch.float32:
            x = x.to(dt)                                               # one rounding serves SA1's gather and FP1's skip GEMM
        skips = [(x, pos, ptr)]
        refl = reflectance
        for lvl in self.sa:
            idx = ops.voxel_sample(pos, lvl["res"], batch, ptr=ptr, group_ptr=group_ptr)
            batch_t = batch[idx]
            ptr_t = ops.batch_to_ptr(batch_t, T)
            if lvl["res"] == 0.04:                                     # src/model.py:117-118
                nbr, _ = ops.radius_table(pos, pos[idx], lvl["res"] * 2, ptr, ptr_t, lvl["k"])
            else:
                # the fused conv takes a max over a row: the neighbour SET is all it needs (P2W_KNN_UNORDERED)
                nbr = ops.knn_table(pos, pos[idx], lvl["k"], ptr, ptr_t, unordered=True)
            pos4, back = ops.sa_prepare(pos, refl, ptr, sf)
            if self.conv_mode == ops.CONV_BF16_TC:          # targets addressed through idx: no pos4[idx] gather
                # the kernel rounds the rows to bf16 as it gathers them (each ~14 times): round them once instead
                xb = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
                h = ops.pointnet_conv_max(xb, pos4, pos4, nbr, *lvl["w"], mode=self.conv_mode, ws=lvl["ws"],
                                          packed=lvl["packed"], out_dtype=dt, tgt_index=idx)
            else:
                h = ops.pointnet_conv_max(x, pos4, pos4[idx], nbr, *lvl["w"], mode=self.conv_mode, ws=lvl["ws"],
                                          packed=lvl["packed"], out_dtype=dt)
            lvl["packed"] = lvl["packed"] or idx.numel() > 0          # an empty call returns before packing
            x = lvl["residual"](h)
            pos, batch, refl, ptr = back[idx], batch_t, refl[idx], ptr_t
            skips.append((x, pos, ptr))
        # ---- GlobalSAModule (src/model.py:134-140)
        n3 = x.size(0)
        buf = torch.zeros((n3, x.size(1) + 3 + self.g_pad), device=x.device, dtype=dt)
        buf[:, : x.size(1)] = x
        buf[:, x.size(1): x.size(1) + 3] = pos
        h = self.g2(self.g1(buf))
        x = ops.global_max_pool(h, batch, ptr=ptr, scale=self.g_s, shift=self.g_t)   # BN + pooling in one pass: [T,512] fp32
        pos_c = pos.new_zeros((T, 3))
        ptr_c = torch.arange(T + 1, device=pos.device, dtype=torch.int64)
        # ---- FPModules (src/model.py:148-153), coarse -> fine
        for fp, (x_skip, pos_skip, ptr_skip) in zip(self.fp, reversed(skips)):
            y = torch.mm(x if x.dtype == dt else x.to(dt), fp["wc"])                          # coarse rows
            z = torch.addmm(fp["bs"], x_skip if x_skip.dtype == dt else x_skip.to(dt), fp["ws"])    # fine rows
            h = ops.knn_interpolate_add_(y, pos_c, pos_skip, z, fp["k"], ptr_c, ptr_skip, relu=True)
            x = fp["l2"](h)
            pos_c, ptr_c = pos_skip, ptr_skip
        x = self.head1(x)
        if self.head2_b is not None and x.size(1) % 8 == 0 and x.size(1) <= 1024:
            return ops.rowdot(x, self.head2_w, self.head2_b)
        return self.head2(x).reshape(-1).float()
