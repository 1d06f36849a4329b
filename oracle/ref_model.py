"""CPU restatement of PointsToWood's eval-mode forward pass (torch CPU fp32, functional).

TEST INFRASTRUCTURE: the checker for the CUDA path and the timed CPU arm of bench.py.
It follows /root/reference/pointstowood/src/model.py:226-245 (Net.forward), :108-127
(SAModule.forward), :134-140 (GlobalSAModule), :148-153 (FPModule), :75-85
(InvertedResidualBlock) and src/pointnet.py:116-132 (PointNetConv.message, aggr='max'),
driven directly by a reference-format state dict.  The
neighbourhood primitives come from oracle.py (Appendix A; parity unpinned -- see there).

Pinned here, in this container, against the reference's OWN model code imported
unmodified through oracle/shim (oracle/make_golden.py writes tests/golden/net_*.npz and
tests/test_oracle_golden.py replays them): the two agree bit for bit on the index outputs
and to float rounding on the logits.

Choices where the reference is nondeterministic or constant:
the ReflectanceYesNo gate is exactly 1.0 (gumbel_softmax over one element), so it is
skipped; voxel representatives are the highest member index (A.5).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as O

BN_EPS = 1e-5


_BN_TRAIN = False        # net_forward_train flips it: batch statistics, as nn.BatchNorm1d in train mode


def _bn(sd, p, x):
    """BatchNorm1d over the channel (last) dim of [N, C]: running statistics in eval mode, batch
    statistics (biased variance, running buffers left alone) in train mode."""
    if _BN_TRAIN:
        return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], True, 0.0, BN_EPS)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def _mlp(sd, p, x, n):
    """model.py:198-202: per layer Linear -> ReLU -> BN, no BN on the first layer."""
    for i in range(n):
        x = F.relu(F.linear(x, sd[f"{p}.{i}.0.weight"], sd[f"{p}.{i}.0.bias"]))
        if i != 0:
            x = _bn(sd, f"{p}.{i}.2", x)
    return x


def _dwsep(sd, p, x):
    """model.py:37-44 on [N, C]: k=1 depthwise conv is a per-channel affine."""
    x = x * sd[p + ".depthwise_conv.weight"].view(1, -1) + sd[p + ".depthwise_conv.bias"]
    x = F.relu(_bn(sd, p + ".depthwise_bn", x))
    x = F.linear(x, sd[p + ".pointwise_conv.weight"].squeeze(-1), sd[p + ".pointwise_conv.bias"])
    return F.relu(_bn(sd, p + ".pointwise_bn", x))


def _residual(sd, p, x):
    """model.py:75-85 (shortcut is the identity: in_channels == out_channels)."""
    out = F.linear(x, sd[p + ".expand.0.weight"].squeeze(-1), sd[p + ".expand.0.bias"])
    out = F.relu(_bn(sd, p + ".expand.1", out))
    out = _dwsep(sd, p + ".conv.0", out)
    out = F.relu(_bn(sd, p + ".conv.1", out))
    out = _dwsep(sd, p + ".conv.3", out)
    out = _bn(sd, p + ".conv.4", out)
    out = F.linear(out, sd[p + ".project.0.weight"].squeeze(-1), sd[p + ".project.0.bias"])
    out = _bn(sd, p + ".project.1", out)
    return F.relu(out + x)


def voxelsample(pos, batch, res):
    """model.py:103-106."""
    ids = O.voxel_grid(pos.numpy(), res, batch.numpy())
    _, perm = O.consecutive_cluster(ids)
    return torch.from_numpy(perm)


def pointnet_conv(sd, p, x, pos4_src, pos4_tgt, nbr):
    """pointnet.py:116-132 + aggr='max' on a [Nt, K] -1-padded neighbour table."""
    nbr = torch.as_tensor(nbr)
    valid = nbr >= 0
    i = torch.arange(nbr.size(0)).view(-1, 1).expand_as(nbr)[valid]
    j = nbr[valid]
    rel = pos4_src[j, :3] - pos4_tgt[i, :3]
    dist = torch.norm(rel, dim=1, keepdim=True)
    maxd = torch.zeros(nbr.size(0), 1).scatter_reduce(0, i.view(-1, 1), dist, "amax", include_self=False)
    msg = torch.cat([x[j], rel / (maxd[i] + 1e-8), pos4_src[j, 3:4]], 1)
    msg = _mlp(sd, (p + "." if p else "") + "local_nn", msg, 2)
    out = msg.new_zeros(nbr.size(0), msg.size(1))
    return out.scatter_reduce(0, i.view(-1, 1).expand_as(msg), msg, "amax", include_self=False)


def sa_module(sd, name, x, pos, batch, refl, sf, res, k, trace=None, idx=None):
    """model.py:108-127; eval mode samples by voxel, train mode takes the given random half `idx`
    (model.py:97-101,114: sorted randperm, pinned by the caller)."""
    pos4 = torch.cat([pos[:, :3], refl.unsqueeze(-1)], -1)
    if idx is None:
        idx = voxelsample(pos4[:, :3], batch, res)
    B = sf.numel()
    ptr_x = O.batch_to_ptr(batch.numpy(), B)
    ptr_y = O.batch_to_ptr(batch[idx].numpy(), B)
    if res == 0.04:
        nbr, _ = O.radius(pos4[:, :3].numpy(), pos4[idx, :3].numpy(), res * 2, ptr_x, ptr_y, k)
    else:
        nbr = O.knn(pos4[:, :3].numpy(), pos4[idx, :3].numpy(), k, ptr_x, ptr_y)
    s = sf[batch].unsqueeze(-1)
    pos4[:, :3] = pos4[:, :3] / s
    x = pointnet_conv(sd, name + ".conv", x, pos4, pos4[idx], nbr)
    pos4[:, :3] = pos4[:, :3] * s
    if trace is not None:
        trace[name] = dict(idx=idx.numpy().copy(), nbr=np.asarray(nbr).copy(), conv=x.detach().numpy().copy())
    x = _residual(sd, name + ".residual_block", x)
    return x, pos4[idx, :3], batch[idx], refl[idx]


def knn_interpolate(x, pos_x, pos_y, batch_x, batch_y, k, B):
    """Appendix A.8 (model.py:149)."""
    nbr = torch.from_numpy(O.knn(pos_x.numpy(), pos_y.numpy(), k, O.batch_to_ptr(batch_x.numpy(), B),
                                 O.batch_to_ptr(batch_y.numpy(), B)))
    valid = nbr >= 0
    yi = torch.arange(nbr.size(0)).view(-1, 1).expand_as(nbr)[valid]
    xi = nbr[valid]
    diff = pos_x[xi] - pos_y[yi]
    w = 1.0 / torch.clamp((diff * diff).sum(-1, keepdim=True), min=1e-16)
    num = x.new_zeros(pos_y.size(0), x.size(1)).index_add_(0, yi, x[xi] * w)
    den = w.new_zeros(pos_y.size(0), 1).index_add_(0, yi, w)
    return num / den


@torch.no_grad()
def net_forward(sd, pos, reflectance, batch, sf, trace=None):
    """model.py:226-245 -> logits [N0] (eval mode)."""
    return _net_forward(sd, pos, reflectance, batch, sf, trace, (None, None, None))


def net_forward_train(sd, pos, reflectance, batch, sf, idx_list, trace=None):
    """The same forward in TRAIN mode (batch-statistics BatchNorm, the three random halves given by the
    caller), with autograd on: tensors of `sd` that require grad receive gradients from the result."""
    global _BN_TRAIN
    _BN_TRAIN = True
    try:
        return _net_forward(sd, pos, reflectance, batch, sf, trace, idx_list)
    finally:
        _BN_TRAIN = False


def poly1_focal_loss(logits, labels, epsilon=0.1, gamma=2.0, alpha=None, label_smoothing=0.1, eps=1e-6):
    """src/loss.py:27-80 with the trainer's settings (src/trainer.py:113: reduction mean, gamma 2, alpha None,
    label smoothing 0.1): clamped logits, smoothed targets, BCE x clamped focal weight + epsilon (1-pt)^(gamma+1)."""
    z = torch.clamp(logits, min=-10, max=10)
    y = labels * (1 - label_smoothing) + 0.5 * label_smoothing if label_smoothing is not None else labels
    p = torch.clamp(torch.sigmoid(z), min=eps, max=1 - eps)
    ce = torch.clamp(F.binary_cross_entropy_with_logits(z, y, reduction="none"), max=100.0)
    pt = torch.clamp(y * p + (1 - y) * (1 - p), min=eps, max=1 - eps)
    loss = torch.clamp(torch.pow(1 - pt, gamma), max=2.0) * ce
    if alpha is not None:
        loss = (alpha * y + (1 - alpha) * (1 - y)) * loss
    loss = loss + torch.clamp(epsilon * torch.pow(1 - pt, gamma + 1), max=100.0)
    loss = torch.clamp(loss, min=0.0, max=100.0)
    return torch.where(torch.isnan(loss), torch.zeros_like(loss), loss).mean()


def _net_forward(sd, pos, reflectance, batch, sf, trace, idx_list):
    B = sf.numel()
    x0 = _mlp(sd, "stem_mlp", pos[:, :3], 1)
    l0 = (x0, pos, batch, reflectance)
    l1 = sa_module(sd, "sa1_module", *l0, sf, 0.04, 32, trace, idx_list[0])
    l2 = sa_module(sd, "sa2_module", *l1, sf, 0.08, 32, trace, idx_list[1])
    l3 = sa_module(sd, "sa3_module", *l2, sf, 0.16, 32, trace, idx_list[2])
    # GlobalSAModule: model.py:134-140
    x4 = _mlp(sd, "sa4_module.NN", torch.cat([l3[0], l3[1]], 1), 2)
    g = x4.new_zeros(B, x4.size(1)).scatter_reduce(0, l3[2].view(-1, 1).expand_as(x4), x4, "amax",
                                                   include_self=False)
    pos4, batch4 = pos.new_zeros(B, 3), torch.arange(B)
    # FP modules: model.py:148-153
    x = knn_interpolate(g, pos4, l3[1], batch4, l3[2], 2, B)
    x = _mlp(sd, "fp4_module.NN", torch.cat([x, l3[0]], 1), 2)
    x = knn_interpolate(x, l3[1], l2[1], l3[2], l2[2], 2, B)
    x = _mlp(sd, "fp3_module.NN", torch.cat([x, l2[0]], 1), 2)
    x = knn_interpolate(x, l2[1], l1[1], l2[2], l1[2], 2, B)
    x = _mlp(sd, "fp2_module.NN", torch.cat([x, l1[0]], 1), 2)
    x = knn_interpolate(x, l1[1], l0[1], l1[2], l0[2], 2, B)
    x = _mlp(sd, "fp1_module.NN", torch.cat([x, l0[0]], 1), 2)
    # head: model.py:241-243
    x = F.linear(x, sd["conv1.weight"].squeeze(-1), sd["conv1.bias"])
    x = F.relu(_bn(sd, "norm", x))
    x = F.linear(x, sd["conv2.weight"].squeeze(-1), sd["conv2.bias"])
    return x.squeeze(-1).float()


def seeded_state_dict(seed: int = 141190, bn_seed: int = 5, randomise: bool = True):
    """Reference-format state dict with seeded weights (the checkpoint is absent:
    /root/reference/.MISSING_LARGE_BLOBS).  Shapes follow SURVEY.md Appendix D; init
    follows model.py:9-16 (Xavier-uniform Linear, Kaiming-uniform Conv1d, zero bias).
    BN affine and running statistics are randomised so eval-mode BN is non-trivial."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, o, i):
        a = (6.0 / (i + o)) ** 0.5
        sd[name + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * a
        sd[name + ".bias"] = torch.zeros(o)

    def conv(name, o, i):
        a = (6.0 / i) ** 0.5                   # kaiming_uniform, fan_in, relu gain sqrt(2)
        sd[name + ".weight"] = ((torch.rand(o, i, generator=g) * 2 - 1) * a).unsqueeze(-1)
        sd[name + ".bias"] = torch.zeros(o)

    def bn(name, c):
        sd[name + ".weight"] = torch.ones(c)
        sd[name + ".bias"] = torch.zeros(c)
        sd[name + ".running_mean"] = torch.zeros(c)
        sd[name + ".running_var"] = torch.ones(c)
        sd[name + ".num_batches_tracked"] = torch.tensor(0)

    def mlp(name, ch):
        for i in range(1, len(ch)):
            lin(f"{name}.{i - 1}.0", ch[i], ch[i - 1])
            if i != 1:
                bn(f"{name}.{i - 1}.2", ch[i])

    C = 32
    mlp("stem_mlp", [3, C])
    for n, (cin, h, cout) in enumerate([(C + 4, 2 * C, 4 * C), (4 * C + 4, 6 * C, 8 * C),
                                        (8 * C + 4, 12 * C, 16 * C)], 1):
        p = f"sa{n}_module"
        mlp(p + ".conv.local_nn", [cin, h, cout])
        e = 4 * cout
        r = p + ".residual_block"
        conv(r + ".expand.0", e, cout); bn(r + ".expand.1", e)
        for d in (0, 3):
            sd[f"{r}.conv.{d}.depthwise_conv.weight"] = ((torch.rand(e, 1, generator=g) * 2 - 1) * 6.0 ** 0.5).unsqueeze(-1)
            sd[f"{r}.conv.{d}.depthwise_conv.bias"] = torch.zeros(e)
            bn(f"{r}.conv.{d}.depthwise_bn", e)
            conv(f"{r}.conv.{d}.pointwise_conv", e, e)
            bn(f"{r}.conv.{d}.pointwise_bn", e)
            bn(f"{r}.conv.{d + 1}", e)
        conv(r + ".project.0", cout, e); bn(r + ".project.1", cout)
        lin(p + ".reflectanceyesno.fc1", 32, 1)
        lin(p + ".reflectanceyesno.fc2", 32, 32)
        lin(p + ".reflectanceyesno.fc3", 1, 32)
    mlp("sa4_module.NN", [16 * C + 3, 16 * C, 16 * C])
    mlp("fp4_module.NN", [32 * C, 24 * C, 16 * C])
    mlp("fp3_module.NN", [24 * C, 20 * C, 16 * C])
    mlp("fp2_module.NN", [20 * C, 16 * C, 16 * C])
    mlp("fp1_module.NN", [17 * C, 16 * C, 16 * C])
    conv("conv1", 16 * C, 16 * C)
    conv("conv2", 1, 16 * C)
    bn("norm", 16 * C)
    if randomise:        # False: BatchNorm as freshly initialised (weight 1, bias 0), the state training starts from
        randomise_bn(sd, bn_seed)
    return sd


def randomise_bn(sd, seed: int = 5):
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd):
        if k.endswith(".running_mean"):
            p = k[: -len(".running_mean")]
            c = sd[k].numel()
            sd[p + ".running_mean"] = torch.randn(c, generator=g) * 0.1
            sd[p + ".running_var"] = torch.rand(c, generator=g) * 0.5 + 0.75
            sign = torch.where(torch.rand(c, generator=g) < 0.15, -1.0, 1.0)   # BN scale may be negative
            sd[p + ".weight"] = (torch.rand(c, generator=g) * 0.5 + 0.75) * sign
            sd[p + ".bias"] = torch.randn(c, generator=g) * 0.1
    return sd
