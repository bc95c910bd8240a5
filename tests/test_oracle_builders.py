"""The oracle's restatement of every component graph, and this repository's component weight order, pinned by the
reference's own BUILDER code (CPU).

tests/golden/make_golden.py imports the reference's model_components/*.py, models/unet.py, models/discriminator.py and
layers/stn_spline.build_locnet unmodified and runs them on a define-then-run numpy Keras (tests/golden/keras_graph.py;
inference phase, random weights for every layer).  tests/golden/golden_builders.npz holds, per component, the inputs,
the outputs and the weights in the component's weight order.  Here the weights go into THIS repository's components
through ``Model.set_weights`` (the Keras-ordered list the executors and the SWA callback use), come back by name through
``named_weights`` and feed the oracle (oracle/ref_models.py): the outputs must agree.  A layer wired differently, a
shared layer that is not shared, or a weight list in another order fails here; the CUDA kernels are then checked against
the same oracle functions in tests/test_models_gpu.py.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ref_models as RM
from oracle import ref_ops as R

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_builders.npz"))
S = 48
TOL = 2e-6          # float32 storage of the golden outputs


def golden_weights(tag):
    k, shapes, so = G[tag + "_wk"], G[tag + "_wshape"], G[tag + "_wso"]
    out, pos = [], 0
    for shp, (scale, offset) in zip(shapes, so):
        shp = tuple(int(v) for v in shp if v > 0)
        n = int(np.prod(shp))
        out.append((offset + k[pos:pos + n].astype(np.float64) * scale).astype(np.float32).reshape(shp))
        pos += n
    assert pos == k.size
    return out


def t(a):
    return torch.from_numpy(np.asarray(a, np.float64))


def inputs(tag):
    return [t(G["%s_in%d" % (tag, i)]) for i in range(8) if "%s_in%d" % (tag, i) in G.files]


def close(got, key, tol=TOL):
    ref = G[key].astype(np.float64)
    got = got.detach().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    if got.ndim == 4 and got.shape[1] == S and ref.shape[1] == S // 2:
        got = got[:, ::2, ::2]                      # the fixture keeps every other pixel of the maps
    assert got.shape == ref.shape, (key, got.shape, ref.shape)
    err = np.abs(got - ref).max() / max(1.0, np.abs(ref).max())
    assert err < tol, (key, err)


def weights_of(model, tag):
    ws = golden_weights(tag)
    assert [tuple(w.shape) for w in ws] == [tuple(p.shape) for p in model.weight_list()], model.name
    model.set_weights(ws)
    return {k: t(v) for k, v in model.named_weights().items()}


@pytest.fixture(scope="module")
def net():
    from multimodal_segmentation_b200.configuration import dafnet_config_chaos
    from multimodal_segmentation_b200.keras_like import EasyDict
    from multimodal_segmentation_b200.models.dafnet import DAFNet
    conf = EasyDict(dafnet_config_chaos.get((S, S, 1)))
    conf.anatomy_encoder.filters = 2
    conf.anatomy_encoder.rounding = False
    conf.d_mask_params.filters = 4
    conf.d_image_params.filters = 4
    conf.automatedpairing = True
    conf.n_pairs = 3
    conf.folder = "/tmp/dafk_test_no_such_folder"
    np.random.seed(0)
    n = DAFNet(conf)
    n.build()
    return n


def test_segmentor(net):
    W = weights_of(net.Segmentor, "segmentor")
    close(RM.segmentor(W, *inputs("segmentor"), RM.BNState(W, False)), "segmentor_out0")


def test_modality_encoder(net):
    W = weights_of(net.Enc_Modality, "modality_encoder")
    mu, lv = RM.modality_encoder(W, *inputs("modality_encoder"))
    close(mu, "modality_encoder_out0")
    close(lv, "modality_encoder_out1")
    close(R.kl(mu, lv), "modality_encoder_out2")


def test_film_decoder(net):
    W = weights_of(net.Decoder, "decoder_film")
    close(RM.decoder_film(W, *inputs("decoder_film")), "decoder_film_out0")


def test_anatomy_fuser(net):
    W = weights_of(net.Anatomy_Fuser, "anatomy_fuser")
    deformed, fused, theta = RM.anatomy_fuser(W, *inputs("anatomy_fuser"))
    close(theta, "anatomy_fuser_out2")
    # the reference layer carries the sampling grid through float32 (layers/stn_spline.py:38-53): ~1e-6 pixel, times the
    # unit slope of a bilinearly sampled one-hot map (measured 1.7e-5)
    close(deformed, "anatomy_fuser_out0", 1e-4)
    close(fused, "anatomy_fuser_out1", 1e-4)


def test_shared_anatomy_encoders(net):
    e1, e2 = net.Encoders_Anatomy
    W = weights_of(e1, "anatomy_encoders_1")
    W.update(weights_of(e2, "anatomy_encoders_2"))
    # the up path and the 1x1 head are ONE set of layers in both models (model_components/anatomy_encoder.py:57-70)
    w1, w2 = golden_weights("anatomy_encoders_1"), golden_weights("anatomy_encoders_2")
    n_down = len(w1) - sum(1 for a, b in zip(w1[::-1], w2[::-1]) if np.array_equal(a, b))
    assert 0 < n_down < len(w1) and all(np.array_equal(a, b) for a, b in zip(w1[n_down:], w2[n_down:]))
    st = RM.BNState(W, False)
    close(RM.anatomy_encoder(W, *inputs("anatomy_encoders_1"), st, "enc1_", "shared_", rounding=False), "anatomy_encoders_1_out0")
    close(RM.anatomy_encoder(W, *inputs("anatomy_encoders_2"), st, "enc2_", "shared_", rounding=False), "anatomy_encoders_2_out0")


def test_single_anatomy_encoder():
    from multimodal_segmentation_b200.keras_like import BuildScope, EasyDict
    from multimodal_segmentation_b200.model_components import anatomy_encoder
    ae = EasyDict(dict(input_shape=(S, S, 1), output_shape=(S, S, 8), out_channels=8, filters=2, downsample=4,
                       normalise="batch", rounding=False))
    with BuildScope(rng=np.random.RandomState(0)):
        m = anatomy_encoder.build(ae)
    W = weights_of(m, "anatomy_encoder")
    close(RM.anatomy_encoder(W, *inputs("anatomy_encoder"), RM.BNState(W, False), "", "", rounding=False), "anatomy_encoder_out0")


def test_discriminator(net):
    W = weights_of(net.D_Mask, "discriminator")
    close(RM.discriminator(W, "D_Mask", *inputs("discriminator")), "discriminator_out0")


def test_balancer(net):
    W = weights_of(net.Balancer, "balancer")
    xs = inputs("balancer")
    close(RM.balancer(W, xs[0], xs[1:]), "balancer_out0")


def test_generator_step_graph_matches_the_reference_trainer():
    """models/dafnet.py:140-222,336-350: the reference DAFNet class wires its supervised trainer out of its own
    components; its 20 outputs (inference phase, fixed reparametrisation noise) against the oracle's restatement of the
    generator-step graph, with the weights loaded component by component into this repository's DAFNet"""
    from multimodal_segmentation_b200.configuration import dafnet_config_chaos
    from multimodal_segmentation_b200.keras_like import EasyDict
    from multimodal_segmentation_b200.models.dafnet import DAFNet
    conf = EasyDict(dafnet_config_chaos.get((S, S, 1)))
    conf.anatomy_encoder.filters = 2
    conf.d_mask_params.filters = 4
    conf.d_image_params.filters = 4
    conf.n_pairs = 1
    conf.folder = "/tmp/dafk_test_no_such_folder"
    assert conf.anatomy_encoder.rounding and not conf.automatedpairing
    np.random.seed(0)
    n = DAFNet(conf)
    n.build()
    W = {}
    for tag, m in (("enc1", n.Encoders_Anatomy[0]), ("enc2", n.Encoders_Anatomy[1]), ("encm", n.Enc_Modality),
                   ("fuser", n.Anatomy_Fuser), ("seg", n.Segmentor), ("dec", n.Decoder), ("dmask", n.D_Mask),
                   ("dimg1", n.D_Image1), ("dimg2", n.D_Image2)):
        W.update(weights_of(m, "trainer_" + tag))
    x1, x2, z1, z2, eps = (t(G["trainer_in%d" % i]) for i in range(5))
    m_dummy = torch.zeros(x1.shape[0], S, S, 5, dtype=torch.float64)
    c = dict(num_masks=4, decoder_type="film", w_sup_M=10, w_adv_M=1, w_rec_X=1, w_adv_X=1, w_rec_Z=1, w_kl=0.1)
    _, _, inter, _ = RM.dafnet_generator_loss(W, c, x1, x2, z1, z2, eps, eps, m_dummy, m_dummy, supervised=True,
                                              training=False)
    outs = inter["outputs"]
    assert len(outs) == 20
    # the compiled losses and loss weights (models/dafnet.py:145-149) on the executor's targets
    # (model_executors/dafnet_executor.py:404-410), evaluated by the stand-in as Keras sums them
    m1, m2 = t(G["trainer_m1"]), t(G["trainer_m2"])
    _, L, _, _ = RM.dafnet_generator_loss(W, c, x1, x2, z1, z2, eps, eps, m1, m2, supervised=True, training=False)
    assert np.allclose([v.item() for v in L.values()], G["trainer_loss"], rtol=1e-4, atol=1e-6), (list(L), G["trainer_loss"])
    _, Lu, _, _ = RM.dafnet_generator_loss(W, c, x1, x2, z1, z2, eps, eps, m1, None, supervised=False, training=False)
    assert np.allclose([v.item() for v in Lu.values()], G["trainer_unsup_loss"], rtol=1e-4, atol=1e-6)
    for i, o in enumerate(outs):
        # outputs downstream of the TPS warp inherit its float32 sampling grid (see test_anatomy_fuser); one binarised
        # anatomy pixel on the 0.5 boundary would show up as an O(1) difference
        close(o, "trainer_out%02d" % i, 1e-4)
    # the unsupervised trainer (models/dafnet.py:151-155): 18 outputs, each equal to one of the supervised ones
    _, _, inter_u, _ = RM.dafnet_generator_loss(W, c, x1, x2, z1, z2, eps, eps, m_dummy, None, supervised=False,
                                                training=False)
    idx = G["trainer_unsup_index"]
    assert len(inter_u["outputs"]) == 18 == len(idx)
    for o, j in zip(inter_u["outputs"], idx):
        close(o, "trainer_out%02d" % int(j), 1e-4)


def _expert_weights(n):
    W = {}
    for tag, m in (("enc1", n.Encoders_Anatomy[0]), ("enc2", n.Encoders_Anatomy[1]), ("encm", n.Enc_Modality),
                   ("fuser", n.Anatomy_Fuser), ("seg", n.Segmentor), ("dec", n.Decoder), ("dmask", n.D_Mask),
                   ("dimg1", n.D_Image1), ("dimg2", n.D_Image2)):
        W.update(weights_of(m, "trainer_" + tag))
    return W


@pytest.mark.parametrize("mi,ty", [(1, "simple"), (1, "def"), (1, "max"), (1, "maxnostn"), (0, "simple"), (0, "def")])
def test_predict_mask_matches_the_reference_method(net, mi, ty):
    """models/mmsdnet.py:210-232 `predict_mask` run by the reference class on its own components"""
    W = _expert_weights(net)
    x = [t(G["trainer_in0"]), t(G["trainer_in1"])]
    close(RM.predict_mask(W, mi, ty, x), "predict_mask_%d_%s" % (mi, ty), 1e-4)


def test_mask_discriminator_trainer_loss_matches_the_reference(net):
    """models/mmsdnet.py:62-78 + models/discriminator.py:36-41 + layers/spectralnorm.py:199-239: mse(D(real), 1),
    mse(D(fake), 0) and the three Spectral regularisation terms Keras adds for the regularised kernels"""
    W = _expert_weights(net)
    real = t(G["trainer_m1"])[..., :4]
    fake = t(np.repeat(G["trainer_in0"].astype(np.float64), 4, -1) * 0.5 + 0.5)
    u0s = [t(G["dtrain_u0_%d" % j]) for j in range(3)]
    total, (lr, lf, reg) = RM.discriminator_trainer_loss(W, "D_Mask", real, fake, u0s)
    g = G["dtrain_loss"]
    assert len(g) == 5
    assert np.allclose([lr.item(), lf.item()], g[:2], rtol=1e-6)
    assert np.isclose(reg.item(), g[2:].sum(), rtol=1e-6) and np.isclose(total.item(), g.sum(), rtol=1e-6)


def test_discriminator_step_fakes_match_the_reference_executor(net):
    """model_executors/dafnet_executor.py:511-583 (`train_batch_mask_discriminator`, `train_batch_image_discriminator`)
    run UNMODIFIED on the reference network of the golden run, its trainers' `fit` recording what they are fed: the real
    and fake batches of the two D_Mask and the two D_Image updates against the oracle's restatement of the fake
    generation (oracle/ref_step.py, the CPU-baseline step) followed by this repository's `data_utils.sample`"""
    from oracle import ref_step as RS
    from multimodal_segmentation_b200.utils import data_utils
    W = _expert_weights(net)
    x1, x2 = t(G["dstep_x1"]), t(G["dstep_x2"])
    # reals: the executor draws two mask batches in a row and keeps the organ channels
    assert np.array_equal(G["dstep_mask_real1"], G["dstep_m1"][..., :4]) and np.array_equal(G["dstep_mask_real2"], G["dstep_m2"][..., :4])
    c1, c2 = RS.mask_d_candidates(W, x1, x2, 4)
    np.random.seed(31)
    f1 = data_utils.sample(c1.numpy(), 2)
    f2 = data_utils.sample(c2.numpy(), 2)
    close(f1, "dstep_mask_fake1", 1e-4)
    close(f2, "dstep_mask_fake2", 1e-4)
    eps = t(G["trainer_in4"]).expand(2, -1)
    y1, y2 = RS.image_d_candidates(W, x1, x2, eps, eps)
    np.random.seed(32)
    close(data_utils.sample(y1.numpy(), 2), "dstep_img_fake1", 1e-4)
    close(data_utils.sample(y2.numpy(), 2), "dstep_img_fake2", 1e-4)


def test_automated_pairing_graph_matches_the_reference_trainer(net):
    """models/dafnet.py:224-334,352-361: three candidates per modality, Balancer weights, per-sample dice + swapped
    per-batch cross entropy and `mae_single_input` combined INSIDE the graph -- the reference trainer's 20 outputs against
    the oracle's `dafnet_generator_loss_automated` (inference phase).  Component weights are those of the expert-pairing
    golden (same construction sequence and seed in the generator), plus the Balancer's."""
    W = {}
    for tag, m in (("enc1", net.Encoders_Anatomy[0]), ("enc2", net.Encoders_Anatomy[1]), ("encm", net.Enc_Modality),
                   ("fuser", net.Anatomy_Fuser), ("seg", net.Segmentor), ("dec", net.Decoder), ("dmask", net.D_Mask),
                   ("dimg1", net.D_Image1), ("dimg2", net.D_Image2)):
        W.update(weights_of(m, "trainer_" + tag))
    W.update(weights_of(net.Balancer, "auto_balancer"))
    ins = [t(G["auto_in%d" % i]) for i in range(10)]
    x1_lst, x2_lst, m1, m2, z1, z2 = ins[0:3], ins[3:6], ins[6], ins[7], ins[8], ins[9]
    eps = t(G["trainer_in4"])
    c = dict(num_masks=4, decoder_type="film", w_sup_M=10, w_adv_M=1, w_rec_X=1, w_adv_X=1, w_rec_Z=1, w_kl=0.1)
    # this fixture's network was built with rounding off; the golden run rounds (anatomy_encoder.rounding = True)
    _, _, inter, _ = RM.dafnet_generator_loss_automated(W, c, x1_lst, x2_lst, z1, z2, eps, eps, m1, m2, supervised=True,
                                                        training=False)
    outs = inter["outputs"]
    assert len(outs) == 20
    for i, o in enumerate(outs):
        close(o, "auto_out%02d" % i, 1e-4)
    _, L, _, _ = RM.dafnet_generator_loss_automated(W, c, x1_lst, x2_lst, z1, z2, eps, eps, m1, m2, supervised=True,
                                                    training=False)
    assert np.allclose([v.item() for v in L.values()], G["auto_loss"], rtol=1e-4, atol=1e-6), (list(L), G["auto_loss"])
    # unsupervised variant: no masks of modality 2, 18 outputs
    _, Lu, _, _ = RM.dafnet_generator_loss_automated(W, c, x1_lst, x2_lst, z1, z2, eps, eps, m1, None, supervised=False,
                                                     training=False)
    assert np.allclose([v.item() for v in Lu.values()], G["auto_unsup_loss"], rtol=1e-4, atol=1e-6)


def _mmsdnet_golden_weights():
    from multimodal_segmentation_b200.configuration import mmsdnet_config_chaos
    from multimodal_segmentation_b200.keras_like import EasyDict
    from multimodal_segmentation_b200.models.mmsdnet import MMSDNet
    conf = EasyDict(mmsdnet_config_chaos.get((S, S, 1)))
    conf.anatomy_encoder.filters = 2
    conf.d_mask_params.filters = 4
    conf.folder = "/tmp/dafk_test_no_such_folder"
    np.random.seed(0)
    n = MMSDNet(conf)
    n.build()
    W = {}
    for tag, m in (("enc1", n.Encoders_Anatomy[0]), ("enc2", n.Encoders_Anatomy[1]), ("encm", n.Enc_Modality),
                   ("fuser", n.Anatomy_Fuser), ("seg", n.Segmentor), ("dec", n.Decoder), ("dmask", n.D_Mask)):
        W.update(weights_of(m, "mmsd_" + tag))
    return W


def test_mmsdnet_step_matches_the_reference_executor():
    """model_executors/mmsdnet_executor.py:238-331 run UNMODIFIED on the reference MMSDNet of the golden run (recording
    trainers): the 4*B fake masks the single D_Mask update samples from, the six inference-phase anatomies the Z regressor
    is fitted on, the trainer order of `train_batch` for l_mix in {1, 0.5, 0} -- ONE mask-discriminator update per
    train_batch whatever l_mix -- and (asserted inside the generator) the 24 targets of the supervised trainer."""
    from oracle import ref_step as RS
    from multimodal_segmentation_b200.utils import data_utils
    assert int(G["mmsd_executor_targets_checked"]) == 1
    W = _mmsdnet_golden_weights()
    x1, x2 = t(G["mstep_x1"]), t(G["mstep_x2"])
    cand = RS.mmsdnet_mask_d_candidates(W, x1, x2, 4)
    assert tuple(cand.shape) == (8, S, S, 4)
    np.random.seed(41)
    close(data_utils.sample(cand.numpy(), 2), "mstep_mask_fake", 1e-4)
    for i, a in enumerate(RS.mmsdnet_zreg_anatomies(W, x1, x2)):
        close(a[:, ::2, ::2], "mstep_zreg_s%d" % i, 1e-4)
    assert list(G["mmsd_schedule_l_mix_1"]) == ["supervised_trainer", "Z_Regressor", "D_Mask_trainer"]
    assert list(G["mmsd_schedule_l_mix_0"]) == ["unsupervised_trainer", "Z_Regressor", "D_Mask_trainer"]
    assert list(G["mmsd_schedule_l_mix_0.5"]) == ["supervised_trainer", "Z_Regressor", "unsupervised_trainer", "Z_Regressor",
                                                  "D_Mask_trainer"]


def test_mmsdnet_graph_matches_the_reference_trainer():
    """models/mmsdnet.py:62-192: two independent anatomy encoders (each with its own 1x1 head), the deformed and the
    fused anatomies segmented, re-encoded and decoded, one D_Mask: the reference supervised trainer's 24 outputs against
    the oracle's `mmsdnet_generator_loss` (inference phase, one reparametrisation noise array for the six samplings)"""
    W = _mmsdnet_golden_weights()
    x1, x2, eps = t(G["trainer_in0"]), t(G["trainer_in1"]), t(G["trainer_in4"])
    dummy_m = [torch.zeros(x1.shape[0], S, S, 5, dtype=torch.float64)] * 6
    dummy_x = [x1] * 6
    c = dict(num_masks=4, decoder_type="film", w_sup_M=10, w_adv_M=1, w_rec_X=1, w_kl=0.1, w_rec_Z=1)
    _, _, outs = RM.mmsdnet_generator_loss(W, c, x1, x2, [eps] * 6, dummy_m, dummy_x, supervised=True, rounding=True,
                                           training=False, return_outputs=True)
    assert len(outs) == 24
    for i, o in enumerate(outs):
        close(o, "mmsd_out%02d" % i, 1e-4)
    # the six-input Z regressor (models/mmsdnet.py:194-208)
    za, zz = G["mmsd_zreg_s"].astype(np.float64), t(G["mmsd_zreg_z"])
    s_list = [t(np.roll(za, i, axis=-1)) for i in range(6)]
    z_list = [zz[i:i + 1] for i in range(6)]
    zrec = RM.z_regressor(W, s_list, z_list)
    close(torch.cat(zrec, 0), "mmsd_zreg_out", 1e-5)
    assert np.allclose([c["w_rec_Z"] * R.mae(z, r).item() for z, r in zip(z_list, zrec)], G["mmsd_zreg_loss"], rtol=1e-5)
    # loss list and weights of the supervised trainer (models/mmsdnet.py:181-190) on the executor's targets
    # (model_executors/mmsdnet_executor.py:257-260): masks without the residual channel
    m1, m2 = t(G["trainer_m1"])[..., :4], t(G["trainer_m2"])[..., :4]
    _, L = RM.mmsdnet_generator_loss(W, c, x1, x2, [eps] * 6, [m1, m2, m2, m2, m1, m1], [x1, x2, x2, x2, x1, x1],
                                     supervised=True, rounding=True, training=False)
    assert np.allclose([v.item() for v in L.values()], G["mmsd_loss"], rtol=1e-4, atol=1e-6), (list(L), G["mmsd_loss"])


def test_spade_decoder_matches_the_reference_builder():
    """model_components/decoder.py:67-81 + layers/spade.py:7-55 (config 3): Dense -> 2x2x128, six SPADE blocks with
    nearest-neighbour up-sampling, per-sample instance normalisation, anatomy resized to every resolution, learnt 1x1
    shortcuts without bias.  The 4.3 M weights are not stored: both sides draw the same list from (position, shape)."""
    from tests.golden.make_golden import seeded_weights
    from multimodal_segmentation_b200.configuration import dafnet_config_chaos
    from multimodal_segmentation_b200.keras_like import BuildScope, EasyDict
    from multimodal_segmentation_b200.model_components import decoder
    conf = EasyDict(dafnet_config_chaos.get((64, 64, 1), decoder_type="spade"))
    with BuildScope(rng=np.random.RandomState(0)):
        m = decoder.build(conf)
    shapes = [tuple(p.shape) for p in m.weight_list()]
    assert len(shapes) == int(G["decoder_spade_nw"])
    m.set_weights(seeded_weights(shapes, 4010))
    W = {k: t(v) for k, v in m.named_weights().items()}
    close(RM.decoder_spade(W, t(G["decoder_spade_in0"]), t(G["decoder_spade_in1"]))[:, ::2, ::2], "decoder_spade_out0", 1e-5)


def test_step_schedule_of_the_reference_executor():
    """model_executors/dafnet_executor.py:369-387 `train_batch`, run unmodified with recording trainers: one generator
    update, two mask-discriminator updates, the two image-discriminator updates -- once per labelled / unlabelled branch.
    This is the order DAFNetExecutor.train_batch_on replays (multimodal_segmentation_b200/model_executors/dafnet_executor.py)
    and oracle/ref_step.py restates; the generator-step inputs and targets were asserted inside the generator
    (tests/golden/make_golden.py, "executor_targets_checked")."""
    assert int(G["executor_targets_checked"]) == 1
    d_steps = ["D_Mask_trainer", "D_Mask_trainer", "D_Image1_trainer", "D_Image2_trainer"]
    assert list(G["schedule_l_mix_1"]) == ["supervised_trainer"] + d_steps
    assert list(G["schedule_l_mix_0"]) == ["unsupervised_trainer"] + d_steps
    assert list(G["schedule_l_mix_0.5"]) == ["supervised_trainer"] + d_steps + ["unsupervised_trainer"] + d_steps
