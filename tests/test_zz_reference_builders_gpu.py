"""The CUDA components against the outputs of the reference's OWN builder code (tests/golden/golden_builders.npz, made
by tests/golden/make_golden.py from the reference's model_components / models sources; see tests/test_oracle_builders.py
for the same fixtures against the oracle on the CPU).  Weights go in through ``Model.set_weights`` in Keras order, the
inputs through ``Model.predict`` (inference phase, strict fp32 kernels): fp32 bound 1e-4 relative L2.

The golden UNets have 2 filters (the numpy Keras stand-in that produced them is slow): their 2-channel levels take the
any-channel-count forward kernels (bn_apply / maxpool / nearest resize).  Named to run last.
"""
import os

import numpy as np
import pytest

from tests.util import rel_l2

pytestmark = [pytest.mark.gpu]

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_builders.npz"))
S = 48


def golden_weights(tag):
    k, shapes, so = G[tag + "_wk"], G[tag + "_wshape"], G[tag + "_wso"]
    out, pos = [], 0
    for shp, (scale, offset) in zip(shapes, so):
        shp = tuple(int(v) for v in shp if v > 0)
        n = int(np.prod(shp))
        out.append((offset + k[pos:pos + n].astype(np.float64) * scale).astype(np.float32).reshape(shp))
        pos += n
    return out


def inputs(tag):
    return [G["%s_in%d" % (tag, i)].astype(np.float32) for i in range(8) if "%s_in%d" % (tag, i) in G.files]


def check(got, key, tol=1e-4):
    ref = G[key].astype(np.float64)
    got = np.asarray(got, np.float64)
    if got.ndim == 4 and got.shape[1] == S and ref.shape[1] == S // 2:
        got = got[:, ::2, ::2]
    assert got.shape == ref.shape, (key, got.shape, ref.shape)
    err = rel_l2(got, ref)
    assert err < tol, (key, err)


@pytest.fixture(scope="module")
def net():
    from multimodal_segmentation_b200 import engine as E
    from multimodal_segmentation_b200.configuration import dafnet_config_chaos
    from multimodal_segmentation_b200.keras_like import EasyDict
    from multimodal_segmentation_b200.models.dafnet import DAFNet
    old = E.USE_TC
    E.USE_TC = False                      # strict fp32 kernels
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
    yield n
    E.USE_TC = old


def test_segmentor(net):
    net.Segmentor.set_weights(golden_weights("segmentor"))
    check(net.Segmentor.predict(inputs("segmentor")[0]), "segmentor_out0")


def test_film_decoder(net):
    net.Decoder.set_weights(golden_weights("decoder_film"))
    check(net.Decoder.predict(inputs("decoder_film")), "decoder_film_out0")


def test_anatomy_fuser(net):
    net.Anatomy_Fuser.set_weights(golden_weights("anatomy_fuser"))
    deformed, fused = net.Anatomy_Fuser.predict(inputs("anatomy_fuser"))
    # bilinear samples of one-hot maps: a 1e-6 pixel difference in the sampling grid is a 1e-6 difference in the value
    check(deformed, "anatomy_fuser_out0", 1e-3)
    check(fused, "anatomy_fuser_out1", 1e-3)


def test_shared_anatomy_encoders(net):
    e1, e2 = net.Encoders_Anatomy
    e1.set_weights(golden_weights("anatomy_encoders_1"))
    e2.set_weights(golden_weights("anatomy_encoders_2"))
    check(e1.predict(inputs("anatomy_encoders_1")[0]), "anatomy_encoders_1_out0")
    check(e2.predict(inputs("anatomy_encoders_2")[0]), "anatomy_encoders_2_out0")


def test_discriminator(net):
    net.D_Mask.set_weights(golden_weights("discriminator"))
    check(net.D_Mask.predict(inputs("discriminator")[0]), "discriminator_out0")


def test_balancer(net):
    net.Balancer.set_weights(golden_weights("balancer"))
    check(net.Balancer.predict(inputs("balancer")), "balancer_out0")


@pytest.mark.parametrize("mi,ty", [(1, "simple"), (1, "def"), (1, "max"), (1, "maxnostn"), (0, "simple")])
def test_predict_mask(net, mi, ty):
    """models/mmsdnet.py:210-232 on the weights of the reference's DAFNet golden run (binarised anatomies: a pixel on
    the 0.5 boundary could flip under fp32, hence the looser bound on the soft masks)"""
    from multimodal_segmentation_b200 import engine as E
    from multimodal_segmentation_b200.configuration import dafnet_config_chaos
    from multimodal_segmentation_b200.keras_like import EasyDict
    from multimodal_segmentation_b200.models.dafnet import DAFNet
    assert not E.USE_TC
    conf = EasyDict(dafnet_config_chaos.get((S, S, 1)))
    conf.anatomy_encoder.filters = 2
    conf.d_mask_params.filters = 4
    conf.d_image_params.filters = 4
    conf.n_pairs = 1
    conf.folder = "/tmp/dafk_test_no_such_folder"
    np.random.seed(0)
    n = DAFNet(conf)
    n.build()
    for tag, m in (("enc1", n.Encoders_Anatomy[0]), ("enc2", n.Encoders_Anatomy[1]), ("fuser", n.Anatomy_Fuser),
                   ("seg", n.Segmentor)):
        m.set_weights(golden_weights("trainer_" + tag))
    x = [G["trainer_in0"].astype(np.float32), G["trainer_in1"].astype(np.float32)]
    check(n.predict_mask(mi, ty, x), "predict_mask_%d_%s" % (mi, ty), 2e-2)
