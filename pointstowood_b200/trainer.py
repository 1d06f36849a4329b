"""Training step: the B200 counterpart of src/trainer.py (SemanticTraining) for BASELINE.json configs[4].

What the reference does per step (src/trainer.py:165-186): forward in train mode under autocast,
Poly-1 focal loss (src/loss.py, reduction mean, gamma 2, alpha None, label smoothing 0.1 -- :113),
backward, gradient clipping at 1.0, AdamW(lr 1e-4, weight_decay 1e-2) (:120).  It is single-GPU.
Here the same step runs data-parallel, one process per GPU: every rank forwards / backwards its own
batch of tiles and the gradients are averaged with ONE bucketed NCCL all-reduce per step over NVLink
(18.2 M parameters, 72.6 MB in fp32), BatchNorm statistics stay local -- the only collective of the
whole code base besides the inference path's final gather.

The graph ops of the forward (random half sampling, radius / kNN tables, position scaling) come from
libp2w; the differentiable part runs in torch autograd on the fixed-width neighbour tables
(model.PointNetConv._forward_autograd).  Parity: tests/golden/train.npz, produced by the reference's own
model and loss code (oracle/make_golden_train.py).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import model as M

__all__ = ["Poly1FocalLoss", "freeze_constant_gate", "GradientAllReduce", "train_step", "make_training_batch",
           "SemanticTraining"]


class Poly1FocalLoss(nn.Module):
    """src/loss.py:6-80: focal BCE on clamped logits / probabilities plus the Poly-1 term
    epsilon * (1 - pt)^(gamma + 1); returns (loss, gamma) like the reference."""

    def __init__(self, epsilon: float = 0.1, gamma: float = 2.0, alpha: Optional[float] = 0.25, reduction: str = "none",
                 weight: Optional[Tensor] = None, label_smoothing: Optional[float] = None, eps: float = 1e-6):
        super().__init__()
        self.epsilon, self.gamma, self.alpha, self.reduction = epsilon, gamma, alpha, reduction
        self.weight, self.label_smoothing, self.eps = weight, label_smoothing, eps

    def forward(self, logits: Tensor, labels: Tensor, label_weights: Optional[Tensor] = None):
        z = torch.clamp(logits, min=-10, max=10)
        y = labels
        if self.label_smoothing is not None:
            y = y * (1 - self.label_smoothing) + 0.5 * self.label_smoothing
        p = torch.clamp(torch.sigmoid(z), min=self.eps, max=1 - self.eps)
        ce = torch.clamp(F.binary_cross_entropy_with_logits(z, y, reduction="none", weight=self.weight), max=100.0)
        pt = torch.clamp(y * p + (1 - y) * (1 - p), min=self.eps, max=1 - self.eps)
        loss = torch.clamp(torch.pow(1 - pt, self.gamma), max=2.0) * ce
        if self.alpha is not None:
            loss = (self.alpha * y + (1 - self.alpha) * (1 - y)) * loss
        loss = loss + torch.clamp(self.epsilon * torch.pow(1 - pt, self.gamma + 1), max=100.0)
        loss = torch.clamp(loss, min=0.0, max=100.0)
        loss = torch.where(torch.isnan(loss), torch.zeros_like(loss), loss)
        if self.reduction == "mean":
            loss = loss.mean()
        elif self.reduction == "sum":
            loss = loss.sum()
        return loss, self.gamma


def freeze_constant_gate(net: nn.Module) -> nn.Module:
    """ReflectanceYesNo is the constant 1.0: its parameters get exactly zero
    gradient in the reference; here they simply do not take part in the step."""
    for name, p in net.named_parameters():
        if ".reflectanceyesno." in name:
            p.requires_grad_(False)
    return net


class GradientAllReduce:
    """Data-parallel gradient averaging: the gradients are packed into a few flat fp32 buckets, each
    averaged with one asynchronous NCCL all-reduce issued as soon as it is packed, and unpacked in place.
    (The exchange is 72.6 MB per step: over NVLink 5 / NVSwitch a fraction of a millisecond; what matters
    is few launches, not link count.)"""

    def __init__(self, params: List[Tensor], bucket_bytes: int = 32 << 20):
        import torch.distributed as dist
        self.dist = dist
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets: List[List[Tensor]] = [[]]
        size = 0
        for p in self.params:
            if size + p.numel() * 4 > bucket_bytes and self.buckets[-1]:
                self.buckets.append([])
                size = 0
            self.buckets[-1].append(p)
            size += p.numel() * 4
        dev = self.params[0].device
        self.flat = [torch.empty(sum(p.numel() for p in b), device=dev, dtype=torch.float32) for b in self.buckets]

    def __call__(self) -> None:
        if self.world == 1:
            return
        handles = []
        for flat, bucket in zip(self.flat, self.buckets):
            o = 0
            for p in bucket:
                n = p.numel()
                flat[o:o + n].copy_(p.grad.reshape(-1) if p.grad is not None else torch.zeros(n, device=flat.device))
                o += n
            handles.append(self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM, async_op=True))
        for h, flat, bucket in zip(handles, self.flat, self.buckets):
            h.wait()
            flat.div_(self.world)
            o = 0
            for p in bucket:
                n = p.numel()
                if p.grad is None:
                    p.grad = torch.empty_like(p)
                p.grad.copy_(flat[o:o + n].view_as(p))
                o += n


def train_step(net: nn.Module, optimizer: torch.optim.Optimizer, criterion: nn.Module, data, allreduce=None,
               autocast_bf16: bool = False, max_norm: float = 1.0) -> Dict[str, float]:
    """One optimisation step (src/trainer.py:167-186): zero_grad, train-mode forward, loss, backward,
    [gradient all-reduce], clip_grad_norm_(1.0), optimizer.step.  Returns loss and gradient norm."""
    net.train()
    optimizer.zero_grad(set_to_none=True)
    if autocast_bf16:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = net(data)
    else:
        logits = net(data)
    loss, _ = criterion(logits.float(), data.y)
    loss.backward()
    if allreduce is not None:
        allreduce()
    params = [p for p in net.parameters() if p.requires_grad]
    gnorm = torch.nn.utils.clip_grad_norm_(params, max_norm=max_norm)
    optimizer.step()
    return dict(loss=loss.detach(), grad_norm=gnorm.detach(), logits=logits.detach())


def make_training_batch(cloud: Tensor, labels: Tensor, tiles, tile_ids, device=None) -> SimpleNamespace:
    """TrainingDataset.__getitem__ + collate for labelled tiles taken from a TileStore: mean shift, sf,
    batch vector (as predicter.classify_tiles packs them) plus the per-point labels `y`."""
    from . import ops
    import numpy as np
    ptr = tiles.ptr
    pieces = [tiles.members[int(ptr[t]): int(ptr[t + 1])] for t in tile_ids]
    members = torch.cat(pieces)
    sizes = [int(ptr[t + 1] - ptr[t]) for t in tile_ids]
    bptr = torch.as_tensor(np.concatenate([[0], np.cumsum(sizes)]), device=members.device, dtype=torch.int64)
    pos, refl, batch, shift, sf = ops.pack_tiles(tiles.feat, members, bptr)
    return M.make_data(pos, refl, batch, sf, ptr=bptr, local_shift=shift.reshape(-1), y=labels[members].float())


def broadcast_module_(net: nn.Module, src: int = 0) -> None:
    """Every parameter AND buffer (BatchNorm running statistics) of `net` from rank `src` to all ranks, in a few
    flat buckets per dtype.  Replicas of a data-parallel run must start from the same weights: averaging gradients
    keeps equal what starts equal, nothing more."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    tensors = [p.data for p in net.parameters()] + [b.data for b in net.buffers()]
    by_dtype: Dict[torch.dtype, List[Tensor]] = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for group in by_dtype.values():
        flat = torch.cat([t.reshape(-1) for t in group])
        dist.broadcast(flat, src=src)
        o = 0
        for t in group:
            t.copy_(flat[o:o + t.numel()].view_as(t))
            o += t.numel()


def SemanticTraining(args):
    """src/trainer.py:96-320 reduced to the optimisation loop: args.net (optional), args.batches (an iterable
    of collated batches with .y), args.num_epochs; AdamW(lr 1e-4, wd 1e-2) + OneCycleLR(max_lr 1e-4, total_steps =
    num_epochs, pct_start 0.05, cos, div_factor 100) stepped once per epoch (:122-123,219) + Poly1FocalLoss(mean,
    gamma 2, alpha None, label_smoothing 0.1) as the reference.  Data-parallel when torch.distributed is
    initialised: rank 0's parameters and buffers are broadcast before the first step (a fresh Net is built under
    the reference's seed 141190, src/trainer.py:25-26), gradients are averaged every step, BatchNorm statistics stay
    local to a rank during training and RANK 0's are the ones that are saved.  With args.wdir / args.model the
    final weights are written by rank 0 as {'model_state_dict': ...} (:304-306), the format load_model reads."""
    device = torch.device("cuda")
    net = getattr(args, "net", None)
    if net is None:
        torch.manual_seed(141190)
        net = M.Net(num_classes=1).to(device)
    freeze_constant_gate(net)
    broadcast_module_(net, src=0)
    criterion = Poly1FocalLoss(reduction="mean", gamma=2.0, alpha=None, label_smoothing=0.1)
    optimizer = torch.optim.AdamW([p for p in net.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-2)
    epochs = int(getattr(args, "num_epochs", 1))
    scheduler = torch.optim.lr_scheduler.OneCycleLR(optimizer, max_lr=1e-4, total_steps=max(epochs, 2), pct_start=0.05,
                                                    anneal_strategy="cos", div_factor=100)
    allreduce = GradientAllReduce(list(net.parameters()))
    history, lrs = [], []
    for epoch in range(epochs):
        lrs.append(optimizer.param_groups[0]["lr"])
        for data in args.batches:
            out = train_step(net, optimizer, criterion, data, allreduce,
                             autocast_bf16=getattr(args, "autocast_bf16", False))
            history.append(float(out["loss"]))
        if epoch + 1 < max(epochs, 2):
            scheduler.step()
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if rank == 0 and getattr(args, "wdir", None) and getattr(args, "model", None):
        import os
        os.makedirs(os.path.join(args.wdir, "model"), exist_ok=True)
        torch.save({"model_state_dict": net.state_dict()}, os.path.join(args.wdir, "model", args.model))
    args.net, args.history, args.lr_history = net, history, lrs
    return args
