"""Automated-pairing path (reference models/dafnet.py:224-334,352-361; costs.py:24-26,88-108,138-143;
model_components/balancer.py): kernels of csrc/pairing.cu and the whole generator graph against the oracle.

Tolerances: fp32 kernels against the fp64 oracle, 1e-4 relative L2 (north-star fp32 bound; the loss kernels
accumulate in double, so they are far below it); the whole-graph gradient carries the fp32 conditioning of the deep
BatchNorm UNet, exactly as in tests/test_models_gpu.py (1e-2)."""
import numpy as np
import pytest
import torch

from oracle import ref_models as RM
from oracle import ref_ops as R
from tests.util import cpu, gpu, rel_l2, t

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from multimodal_segmentation_b200 import ops as o
    return o


def _softmax(a):
    e = np.exp(a - a.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


@pytest.mark.parametrize("B,H,W", [(3, 8, 6), (5, 33, 17), (2, 64, 64)])
def test_weighted_per_sample_segmentation_loss(ops, B, H, W):
    """loss = weight * mean_b sum_j w[b,j] * (dice_b + .01 * wBCE_b)(pred_j): value, d/dpred_j (including the path
    through the batch-wide class weights) and d/dw"""
    rs = np.random.RandomState(B)
    P, C, nch, weight = 3, 5, 4, 10.0
    preds = [_softmax(rs.normal(size=(B, H, W, C))).astype(np.float32) for _ in range(P)]
    tgt = np.eye(C, dtype=np.float32)[rs.randint(0, C, size=(B, H, W))]
    w = _softmax(rs.normal(size=(B, P))).astype(np.float32)
    pt = [t(p, torch.float64, grad=True) for p in preds]
    wt = t(w, torch.float64, grad=True)
    ref = weight * sum(wt[:, j:j + 1] * R.combined_dice_bce_perbatch(t(tgt, torch.float64), pt[j], nch)[:, None]
                       for j in range(P)).mean()
    ref.backward()

    L = torch.empty(P, B, device="cuda")
    dp = [gpu(p) for p in preds]
    dt = gpu(tgt)
    wss = [ops.segloss_pb_fwd(dp[j], dt, nch, L[j]) for j in range(P)]
    loss = torch.zeros(1, device="cuda")
    dw, coef = ops.pair_combine(gpu(w), L, weight, loss)
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert rel_l2(cpu(dw), wt.grad.numpy()) < 1e-5
    for j in range(P):
        g = ops.segloss_pb_bwd(dp[j], dt, nch, wss[j], coef[j])
        assert rel_l2(cpu(g), pt[j].grad.numpy()) < 1e-4, j


def test_weighted_per_sample_mae_and_overlap(ops):
    rs = np.random.RandomState(0)
    B, H, W, P, weight = 4, 16, 12, 3, 10.0
    x = rs.uniform(-1, 1, size=(B, H, W, 1)).astype(np.float32)
    ys = [rs.uniform(-1, 1, size=(B, H, W, 1)).astype(np.float32) for _ in range(P)]
    w = _softmax(rs.normal(size=(B, P))).astype(np.float32)
    yt = [t(y, torch.float64, grad=True) for y in ys]
    wt = t(w, torch.float64, grad=True)
    ref = weight * sum(wt[:, j:j + 1] * R.mae_single_input(t(x, torch.float64), yt[j]) for j in range(P)).mean()
    ref.backward()
    L = torch.empty(P, B, device="cuda")
    dx, dy = gpu(x), [gpu(y) for y in ys]
    for j in range(P):
        ops.mae_pb_fwd(dy[j], dx, L[j])
    loss = torch.zeros(1, device="cuda")
    dw, coef = ops.pair_combine(gpu(w), L, weight, loss)
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert rel_l2(cpu(dw), wt.grad.numpy()) < 1e-5
    for j in range(P):
        assert rel_l2(cpu(ops.mae_pb_bwd(dy[j], dx, coef[j])), yt[j].grad.numpy()) < 1e-6
    # Balancer overlap and its backward
    a = (rs.uniform(size=(B, H, W, 8)) > 0.5).astype(np.float32)
    b = rs.uniform(size=(B, H, W, 8)).astype(np.float32)
    at, bt = t(a, torch.float64, grad=True), t(b, torch.float64, grad=True)
    g = rs.normal(size=(B, 1)).astype(np.float32)
    d = R.pair_dice(at, bt)
    (d * t(g, torch.float64)).sum().backward()
    out, ws = ops.pair_dice(gpu(a), gpu(b), want_ws=True)
    assert rel_l2(cpu(out), d.detach().numpy()) < 1e-6
    da, db = ops.pair_dice_bwd(gpu(a), gpu(b), ws, gpu(g))
    assert rel_l2(cpu(da), at.grad.numpy()) < 1e-5 and rel_l2(cpu(db), bt.grad.numpy()) < 1e-5


def _build(seed=3):
    from multimodal_segmentation_b200 import engine as E
    from multimodal_segmentation_b200.configuration import dafnet_config_chaos
    from multimodal_segmentation_b200.keras_like import EasyDict
    from multimodal_segmentation_b200.models.dafnet import DAFNet
    E.USE_TC = False
    conf = EasyDict(dafnet_config_chaos.get((64, 64, 1)))
    conf.anatomy_encoder.filters = 16
    conf.anatomy_encoder.rounding = False
    conf.automatedpairing = True
    conf.n_pairs = 3
    conf.seed = seed
    conf.folder = "/tmp/dafk_test_no_such_folder"
    net = DAFNet(conf)
    net.build()
    rs = np.random.RandomState(0)
    loc = net.Anatomy_Fuser.locnet.layers[-1]     # move theta off the kink of the bilinear sampler
    loc.kernel.data.copy_(torch.from_numpy((rs.normal(size=loc.kernel.shape) * 2e-3).astype(np.float32)))
    return net, conf


@pytest.mark.parametrize("supervised", [True, False])
def test_automated_pairing_generator_step(supervised):
    from multimodal_segmentation_b200.loaders.synthetic_chaos import make_pairs
    from tests.test_models_gpu import all_weights, compare_grads
    net, conf = _build()
    assert any(p.name.startswith("beta") for p in net.generator_params())      # the Balancer is trained
    B, P, H = 2, 3, 64
    cand = [make_pairs(B, (H, H, 1), 4, seed=40 + j) for j in range(P)]
    x1_lst, x2_lst = [c[0] for c in cand], [c[1] for c in cand]
    res = lambda m: np.concatenate([m, 1 - np.clip(m.sum(-1, keepdims=True), 0, 1)], -1).astype(np.float32)
    m1, m2 = res(cand[0][2]), res(cand[0][3])
    rs = np.random.RandomState(5)
    z1, z2, e1, e2 = (rs.normal(size=(B, conf.num_z)).astype(np.float32) for _ in range(4))

    # ---- oracle (fp64)
    W = all_weights(net, torch.float64)
    for k, v in net.Balancer.named_weights().items():
        W[k] = torch.from_numpy(v).double()
    train_names = {p.name for p in net.generator_params()}
    for k in W:
        if k in train_names:
            W[k].requires_grad_(True)
    c = dict(num_masks=conf.num_masks, decoder_type=conf.decoder_type, w_sup_M=conf.w_sup_M, w_adv_M=conf.w_adv_M,
             w_rec_X=conf.w_rec_X, w_adv_X=conf.w_adv_X, w_kl=conf.w_kl, w_rec_Z=conf.w_rec_Z)
    T = lambda a: torch.from_numpy(a).double()
    orig = RM.anatomy_encoder
    RM.anatomy_encoder = lambda *a, **k: orig(*a, rounding=False, **k)
    try:
        total, L, inter, st = RM.dafnet_generator_loss_automated(
            W, c, [T(x) for x in x1_lst], [T(x) for x in x2_lst], T(z1), T(z2), T(e1), T(e2), T(m1),
            T(m2) if supervised else None, supervised)
    finally:
        RM.anatomy_encoder = orig
    total.backward()

    # ---- product
    tr = net.supervised_trainer if supervised else net.unsupervised_trainer
    dev = [torch.from_numpy(a).cuda() for a in x1_lst + x2_lst + [z1, z2, e1, e2, m1] + ([m2] if supervised else [])]
    tr.forward_backward(*dev)
    torch.cuda.synchronize()
    vals = tr.book.buf.cpu().numpy()
    ref = np.array([v.item() for v in L.values()])
    assert len(vals) == len(ref)
    assert np.abs(vals - ref).max() < 1e-4 * max(1.0, np.abs(ref).max()), (vals, ref)
    report = []
    worst, glob = compare_grads(net, W, 2e-3, report)
    assert glob < 1e-2, (glob, report[:5])
    # the Balancer's own weights: shallow path, tight
    for p in net.Balancer.params():
        r = W[p.name].grad.numpy()
        if np.linalg.norm(r) > 1e-9:
            assert rel_l2(p.grad.cpu().numpy(), r) < 1e-3, p.name



def test_experiment_automated_pairing_one_epoch(tmp_path, monkeypatch):
    """experiment.py --automatedpairing 1: train (supervised + unsupervised paired trainers, discriminators) ->
    validate -> save (incl. the Balancer) -> test"""
    import os
    from multimodal_segmentation_b200.experiment import Experiment
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("DAFK_TRAIN_PAIRS", "16")     # two volumes: l_mix = 0.5 labels one of them, the other is unlabelled
    Experiment().run(["--config", "dafnet_config_chaos", "--split", "0", "--l_mix", "0.5", "--input_size", "64",
                      "--epochs", "1", "--batch_size", "4", "--automatedpairing", "1"])
    folder = [f for f in os.listdir(".") if f.startswith("dafnet_chaos_automatedpairing_l05")][0]
    rows = open(os.path.join(folder, "training.csv")).read().strip().split("\n")
    assert len(rows) == 2
    vals = dict(zip(rows[0].split(","), rows[1].split(",")))
    assert np.isfinite(float(vals["rec_X"])) and float(vals["supervised_Mask"]) > 0
    # the reference's columns (dafnet_executor.py:200-205), incl. both deformation directions and the Balancer's weights
    for col in ("val_loss_mod2_mod1def", "val_loss_mod1_mod2def", "val_loss_mod1_fused", "val_weight_0", "val_weight_2"):
        assert col in vals and np.isfinite(float(vals[col])), col
    assert "loss" not in vals
    w = [float(vals["val_weight_%d" % j]) for j in range(3)]          # --automatedpairing 1 sets n_pairs = 3
    assert all(0.0 <= v <= 1.0 for v in w) and abs(sum(w) - 1.0) < 1e-3          # softmax over the n_pairs candidates
    assert os.path.exists(os.path.join(folder, "models", "Balancer.npz")) or \
        any(f.startswith("Balancer") for f in os.listdir(os.path.join(folder, "models")))
