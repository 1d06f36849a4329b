"""GPU tests of the training step (pointstowood_b200/trainer.py, BASELINE.json configs[4]) against the
fixture produced by the reference's own model / loss code in train mode (tests/golden/train.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_model

pytestmark = pytest.mark.gpu

import re

# as tests/test_oracle_train.py: only the Linear / k=1 conv weights have well-conditioned gradient norms
WELL_CONDITIONED = re.compile(r"(stem_mlp\.0\.0|local_nn\.\d\.0|NN\.\d\.0|conv1|conv2)\.weight$")


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pointstowood_b200 import model, trainer
    return model, trainer


def _fixture_batch(model_mod, golden_dir):
    g = np.load(os.path.join(golden_dir, "train.npz"))
    t = lambda k, dt=None: torch.from_numpy(g[k].astype(dt) if dt else g[k]).cuda()
    data = model_mod.make_data(t("pos"), t("reflectance"), t("batch", np.int64), t("sf"), y=t("y"))
    halves = [t(f"idx{i}", np.int64) for i in (1, 2, 3)]
    return g, data, halves


def _net(model_mod, trainer_mod):
    net = model_mod.Net(num_classes=1)
    net.load_state_dict(ref_model.seeded_state_dict(randomise=False), strict=True)
    return trainer_mod.freeze_constant_gate(net.cuda())


def test_train_forward_backward_matches_reference_fixture(mods, golden_dir):
    model_mod, trainer_mod = mods
    g, data, halves = _fixture_batch(model_mod, golden_dir)
    net = _net(model_mod, trainer_mod).train()
    for sa, idx in zip((net.sa1_module, net.sa2_module, net.sa3_module), halves):
        sa.sample_idx = idx                                   # pin the random halves
    logits = net(data)
    loss, _ = trainer_mod.Poly1FocalLoss(reduction="mean", gamma=2.0, alpha=None, label_smoothing=0.1)(logits, data.y)
    loss.backward()
    # tolerances: see tests/test_oracle_train.py (train-mode BatchNorm amplifies FP32 rounding on near-constant channels)
    # (the GPU GEMMs sum in yet another order than the two CPU formulations: a little looser again)
    d = np.abs(logits.detach().cpu().numpy() - g["logits"])
    print("train parity: logits mean/max diff", d.mean(), d.max(), "loss diff", abs(loss.item() - float(g["loss"])))
    assert d.mean() <= 5e-3 and d.max() <= 1e-1
    assert abs(loss.item() - float(g["loss"])) <= 5e-4
    params = dict(net.named_parameters())
    worst = 0.0
    for name, want in zip(g["grad_names"].tolist(), g["grad_norms"].tolist()):
        if WELL_CONDITIONED.search(name):      # see tests/test_oracle_train.py
            rel = abs(float(params[name].grad.norm()) - want) / want
            worst = max(worst, rel)
            assert rel <= 0.10, name
    print("train parity: worst grad-norm deviation", worst)
    for k in g.files:
        if k.startswith("grad.") and k != "grad.fp1_module.NN.1.2.bias":
            a, b = params[k[5:]].grad.cpu().numpy().ravel(), g[k].ravel()
            cos = float(a @ b / np.linalg.norm(a) / np.linalg.norm(b))
            print("train parity: cos", k, cos)
            assert cos >= 0.99, k
    assert all(p.grad is None for n, p in params.items() if "reflectanceyesno" in n)


def test_train_steps_reduce_the_loss(mods, golden_dir):
    model_mod, trainer_mod = mods
    _, data, _ = _fixture_batch(model_mod, golden_dir)
    net = _net(model_mod, trainer_mod)
    gen = torch.Generator(device="cuda").manual_seed(7)
    for sa in (net.sa1_module, net.sa2_module, net.sa3_module):
        sa.generator = gen
    crit = trainer_mod.Poly1FocalLoss(reduction="mean", gamma=2.0, alpha=None, label_smoothing=0.1)
    opt = torch.optim.AdamW([p for p in net.parameters() if p.requires_grad], lr=1e-3, weight_decay=1e-2)
    losses = [float(trainer_mod.train_step(net, opt, crit, data)["loss"]) for _ in range(12)]
    assert np.isfinite(losses).all() and min(losses[-3:]) < 0.8 * losses[0]
    # and the trained weights still serve the eval-mode engine
    net.eval()
    with torch.no_grad():
        out = net(data)
    assert torch.isfinite(out).all() and out.numel() == data.pos.size(0)


def test_training_batch_from_tile_store(mods):
    model_mod, trainer_mod = mods
    from pointstowood_b200.preprocessing import Voxelise
    from pointstowood_b200.synthetic import tls_plot
    cloud, label = tls_plot(60_000, 51, side=5.0)
    dev = torch.from_numpy(cloud).cuda()
    store = Voxelise(dev, minpoints=512, maxpoints=4096, gridsize=(2.0, 4.0)).write_voxels()
    data = trainer_mod.make_training_batch(dev, torch.from_numpy(label).cuda(), store, [0, 1, 2])
    assert data.pos.size(0) == data.y.numel() == int(store.ptr[3]) and data.sf.numel() == 3
    assert set(torch.unique(data.y).tolist()) <= {0.0, 1.0}


def test_semantic_training_schedule_and_checkpoint(mods, golden_dir, tmp_path):
    """src/trainer.py:120-123,219,304-306: OneCycleLR from 1e-6 stepped per epoch, final weights saved as
    {'model_state_dict': ...} -- the file load_model reads back."""
    from types import SimpleNamespace
    model_mod, trainer_mod = mods
    _, data, _ = _fixture_batch(model_mod, golden_dir)
    args = SimpleNamespace(net=_net(model_mod, trainer_mod), batches=[data], num_epochs=40, wdir=str(tmp_path), model="m.pth")
    out = trainer_mod.SemanticTraining(args)
    assert len(out.history) == 40 and np.isfinite(out.history).all()
    # torch's OneCycleLR(max_lr 1e-4, total_steps 40, pct_start 0.05, cos, div_factor 100): 1e-6, 1e-4, then cosine decay
    assert abs(out.lr_history[0] - 1e-6) < 1e-11 and abs(out.lr_history[1] - 1e-4) < 1e-11
    assert max(out.lr_history) <= 1e-4 + 1e-12 and out.lr_history[-1] < out.lr_history[2] < out.lr_history[1]
    ckpt = torch.load(os.path.join(str(tmp_path), "model", "m.pth"), map_location="cpu")
    assert set(ckpt) == {"model_state_dict"}
    fresh = model_mod.Net(num_classes=1)
    model_mod.load_model(os.path.join(str(tmp_path), "model", "m.pth"), fresh, torch.device("cpu"))
    for k, v in out.net.state_dict().items():
        assert torch.equal(v.cpu(), fresh.state_dict()[k]), k
    # a fresh Net is built under the reference's seed: two calls without args.net start from identical weights
    a = trainer_mod.SemanticTraining(SimpleNamespace(net=None, batches=[], num_epochs=1)).net
    b = trainer_mod.SemanticTraining(SimpleNamespace(net=None, batches=[], num_epochs=1)).net
    assert all(torch.equal(p, q) for p, q in zip(a.parameters(), b.parameters()))
