"""Generate tests/golden/train.npz: one TRAIN-mode forward + Poly-1 focal loss + backward of the REFERENCE's
own model and loss code (src/model.py, src/pointnet.py, src/loss.py imported unmodified through
oracle/shim), on seeded inputs / weights, with the three random halves of SAModule.random_sample
(src/model.py:97-101) pinned by the harness (nondeterminism is neutralised, not
reproduced).  Run in the build container only:  python oracle/make_golden_train.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference/pointstowood")

from oracle import ref_model  # noqa: E402
from oracle.make_golden import GOLD  # noqa: E402
from pointstowood_b200.synthetic import tls_plot  # noqa: E402
import src.model as refmodel  # noqa: E402  (the reference, unmodified)
from src.loss import Poly1FocalLoss  # noqa: E402

FULL_GRADS = ["stem_mlp.0.0.weight", "sa1_module.conv.local_nn.0.0.weight", "sa1_module.conv.local_nn.1.2.weight",
              "sa2_module.residual_block.conv.0.depthwise_conv.weight", "fp1_module.NN.1.2.bias", "conv2.weight"]


def make_batch(n_points, side, seed, n_tiles):
    p, lab = tls_plot(n_points, seed, side=side)
    edges = np.linspace(0, side, n_tiles + 1)
    pos, refl, batch, sf, y = [], [], [], [], []
    for t in range(n_tiles):
        m = (p[:, 0] >= edges[t]) & (p[:, 0] < edges[t + 1])
        xyz = torch.from_numpy(p[m, :3].copy())
        xyz = xyz - torch.mean(xyz, axis=0)
        sf.append(torch.sqrt((xyz ** 2).sum(dim=1)).max())
        pos.append(xyz)
        r = p[m, 3]
        refl.append(torch.from_numpy(((r - r.min()) / (r.max() - r.min()) * 2 - 1).astype(np.float32)))
        batch.append(torch.full((int(m.sum()),), t, dtype=torch.long))
        y.append(torch.from_numpy(lab[m].astype(np.float32)))
    return torch.cat(pos), torch.cat(refl), torch.cat(batch), torch.stack(sf), torch.cat(y)


def main():
    sd = ref_model.seeded_state_dict(randomise=False)      # fresh BatchNorm affine: the state training starts from
    pos, refl, batch, sf, y = make_batch(5000, 2.4, 21, 2)
    g = torch.Generator().manual_seed(4)
    halves, n = [], pos.size(0)
    for _ in range(3):
        idx = torch.sort(torch.randperm(n, generator=g)[: int(n * 0.5)]).values
        halves.append(idx)
        n = idx.numel()
    net = refmodel.Net(num_classes=1)
    net.load_state_dict(sd, strict=True)
    net.train()
    queue = list(halves)
    orig = refmodel.SAModule.random_sample
    refmodel.SAModule.random_sample = lambda self, num_points: queue.pop(0)
    try:
        torch.manual_seed(0)                      # ReflectanceYesNo's gumbel noise (its output is 1.0 regardless)
        data = types.SimpleNamespace(pos=pos.clone(), batch=batch.clone(), reflectance=refl.clone(), sf=sf.clone())
        logits = net(data)
        loss, _ = Poly1FocalLoss(reduction="mean", gamma=2.0, alpha=None, label_smoothing=0.1)(logits, y)
        loss.backward()
    finally:
        refmodel.SAModule.random_sample = orig
    out = dict(pos=pos.numpy(), reflectance=refl.numpy(), batch=batch.numpy().astype(np.int32), sf=sf.numpy(),
               y=y.numpy(), idx1=halves[0].numpy().astype(np.int32), idx2=halves[1].numpy().astype(np.int32),
               idx3=halves[2].numpy().astype(np.int32), logits=logits.detach().numpy(), loss=np.float64(loss.item()))
    names, norms = [], []
    for k, p in net.named_parameters():
        if p.grad is not None:
            names.append(k)
            norms.append(float(p.grad.norm()))
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms, dtype=np.float64)
    params = dict(net.named_parameters())
    for k in FULL_GRADS:
        out["grad." + k] = params[k].grad.numpy()
    np.savez_compressed(os.path.join(GOLD, "train.npz"), **out)
    print("loss", loss.item(), "params with grad", len(names), "logits", logits[:4].detach().numpy())


if __name__ == "__main__":
    main()
