"""Stochastic weight averaging (reference callbacks/swa.py:16-47): the running average over the epochs after swa_epoch,
the clone built from the component's own builder, and the write-back at the end of training.  The arithmetic is host
side (once per epoch), so most of it runs on the CPU box; the clone / write-back need device memory."""
import numpy as np
import pytest

from multimodal_segmentation_b200.callbacks.swa import SWA


class _FakeModel(object):
    def __init__(self, ws):
        self.ws = [np.array(w, np.float64) for w in ws]

    def get_weights(self):
        return [w.copy() for w in self.ws]

    def set_weights(self, ws):
        self.ws = [np.array(w) for w in ws]


def test_running_average_matches_the_reference_recurrence():
    m = _FakeModel([[1.0, 2.0], [[3.0]]])
    swa = SWA(2, lambda: None, None)
    swa.model = m
    history = []
    for epoch in range(7):
        m.ws = [w + epoch for w in m.ws]                     # "training" changes the weights
        history.append(m.get_weights())
        swa.on_epoch_end(epoch)
        if epoch <= 2:
            assert all(np.array_equal(a, b) for a, b in zip(swa.swa_weights, history[-1]))     # tracking, not averaging
    # epochs 2 (the last tracked one) .. 6 averaged with equal weights: the recurrence (w*k + cur)/(k+1)
    expect = [np.mean([h[i] for h in history[2:]], axis=0) for i in range(2)]
    assert all(np.allclose(a, b) for a, b in zip(swa.swa_weights, expect))
    swa.on_train_end()
    assert all(np.allclose(a, b) for a, b in zip(m.get_weights(), expect))


def test_running_average_matches_the_reference_callback_bit_for_bit():
    """golden vectors from the reference's own callbacks/swa.py driven over six epochs with swa_epoch=2
    (tests/golden/make_golden.py): same formula, same evaluation order, identical fp32 result"""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_ref.npz"))

    class Live(object):
        e = 0

        def get_weights(self):
            return [G["swa_hist%d_a" % self.e].copy(), G["swa_hist%d_b" % self.e].copy()]

    swa = SWA(2, lambda: None, None)
    swa.model = Live()
    for e in range(6):
        swa.model.e = e
        swa.on_epoch_end(e)
    assert swa.swa_weights[0].dtype == np.float32
    assert np.array_equal(swa.swa_weights[0], G["swa_avg_a"]) and np.array_equal(swa.swa_weights[1], G["swa_avg_b"])


def test_on_train_begin_reads_keras_params(capsys):
    swa = SWA(40, lambda: None, None)
    swa.params = {"epochs": 100}
    swa.on_train_begin()
    assert "last 60 epochs" in capsys.readouterr().out


@pytest.mark.gpu
def test_clone_and_executor_hooks(tmp_path, monkeypatch):
    """the executor's SWA objects follow the live components, save the averaged weights in the per-component files
    (dafnet_executor.py:286-301) and the clone reproduces the component's predictions"""
    import os
    from multimodal_segmentation_b200 import engine as E
    from multimodal_segmentation_b200.configuration import dafnet_config_chaos
    from multimodal_segmentation_b200.keras_like import EasyDict
    from multimodal_segmentation_b200.models.dafnet import DAFNet
    from multimodal_segmentation_b200.model_executors.dafnet_executor import DAFNetExecutor
    monkeypatch.setenv("DAFK_INPUT_SHAPE", "64x64x1")
    E.USE_TC = True
    conf = EasyDict(dafnet_config_chaos.get((64, 64, 1)))
    conf.anatomy_encoder.filters = 16
    conf.n_pairs, conf.l_mix, conf.batch_size, conf.seed = 1, 1.0, 4, 3
    conf.folder = str(tmp_path / "run")
    net = DAFNet(conf)
    net.build()
    ex = DAFNetExecutor(conf, net)
    ex.SWA_EPOCH = 0
    for swa_m in ex.get_swa_models():
        swa_m.swa_epoch = 0
    assert len(ex.get_swa_models()) == 10 and ex.swa_Segmentor.model is net.Segmentor
    x = np.random.RandomState(0).uniform(-1, 1, size=(2, 64, 64, 1)).astype(np.float32)
    w0 = net.Segmentor.get_weights()
    for swa_m in ex.get_swa_models():
        swa_m.on_epoch_end(0)
    net.Segmentor.set_weights([w + 1.0 for w in w0])           # the next epoch moved the weights
    for swa_m in ex.get_swa_models():
        swa_m.on_epoch_end(1)
    avg = [w + 0.5 for w in w0]
    assert all(np.allclose(a, b, atol=1e-6) for a, b in zip(ex.swa_Segmentor.swa_weights, avg))
    clone = ex.swa_Segmentor.get_clone_model()
    assert clone is not net.Segmentor and all(np.allclose(a, b, atol=1e-6) for a, b in zip(clone.get_weights(), avg))
    # an encoder clone built by anatomy_encoder.build reproduces the live (shared-decoder) encoder's output
    enc_clone = ex.swa_Enc_Anatomy1.get_clone_model()
    assert np.array_equal(enc_clone.predict(x), net.Encoders_Anatomy[0].predict(x))
    ex.save_models()
    z = np.load(os.path.join(conf.folder, "models", "Segmentor.npz"))
    assert all(np.allclose(z[n], a, atol=1e-6) for n, a in zip(z["__order__"], avg))
    ex.swa_Segmentor.on_train_end()
    assert all(np.allclose(a, b, atol=1e-6) for a, b in zip(net.Segmentor.get_weights(), avg))
    net.load_models()                                           # the files round-trip through DAFNet.load_models
