"""The executors' data set-up on the host (no GPU): which volumes are labelled / unlabelled for a given l_mix, what the
mask-discriminator pool holds, how many batches an epoch has, and the training.csv columns
(model_executors/dafnet_executor.py:67-184,200-205; mmsdnet_executor.py:67-157)."""
import numpy as np
import pytest


def _executor(monkeypatch, l_mix, pairs=32, kind="dafnet", seed=3):
    monkeypatch.setenv("DAFK_TRAIN_PAIRS", str(pairs))
    monkeypatch.setenv("DAFK_INPUT_SHAPE", "32x32x1")
    from multimodal_segmentation_b200.keras_like import EasyDict
    if kind == "dafnet":
        from multimodal_segmentation_b200.configuration import dafnet_config_chaos as cfg
        from multimodal_segmentation_b200.model_executors.dafnet_executor import DAFNetExecutor as Ex
        from multimodal_segmentation_b200.models.dafnet import DAFNet as Net
    else:
        from multimodal_segmentation_b200.configuration import mmsdnet_config_chaos as cfg
        from multimodal_segmentation_b200.model_executors.mmsdnet_executor import MMSDNetExecutor as Ex
        from multimodal_segmentation_b200.models.mmsdnet import MMSDNet as Net
    conf = EasyDict(cfg.get((32, 32, 1)))
    conf.l_mix, conf.batch_size, conf.seed, conf.n_pairs, conf.folder = l_mix, 4, seed, 1, "/tmp/dafk_no_such_folder"
    conf.anatomy_encoder.filters = 8
    net = Net(conf)
    net.build()
    ex = Ex(conf, net)
    ex.init_train_data()
    return ex, conf


@pytest.mark.parametrize("kind", ["dafnet", "mmsdnet"])
def test_l_mix_half_labels_the_sampled_volumes_and_leaves_the_rest_unlabelled(monkeypatch, kind):
    ex, conf = _executor(monkeypatch, 0.5, pairs=48, kind=kind)          # six volumes of eight slices
    np.random.seed(conf.seed)
    want = np.random.choice(list(range(6)), size=3, replace=False).tolist()      # Data.sample's draw (loaders/data.py:123-136)
    lab = list(dict.fromkeys(ex.data.index.tolist()))
    unl = sorted(set(ex.ul_data.index.tolist()))
    assert lab == want and unl == sorted(set(range(6)) - set(want))
    assert ex.data.size() == 24 and ex.ul_data.size() == 24 and ex.data.num_volumes == ex.ul_data.num_volumes == 3
    assert ex.batches == 6                                                        # ceil(24 / 4)
    assert ex.gen_labelled is not None and ex.gen_unlabelled is not None
    m = next(ex.discriminator_masks)
    assert tuple(m.shape) == (4, 32, 32, 4)
    # the labelled slices of both modalities plus modality 1 of the unlabelled ones (dafnet_executor.py:155-176)
    assert ex.discriminator_masks.flows[0].array.shape[0] == 2 * 24 + 24


def test_l_mix_one_and_zero(monkeypatch):
    ex, _ = _executor(monkeypatch, 1.0)
    assert ex.ul_data is None and ex.gen_unlabelled is None and ex.data.size() == 32 and ex.batches == 8
    ex, _ = _executor(monkeypatch, 0.0)
    assert ex.data is None and ex.gen_labelled is None and ex.ul_data.size() == 32 and ex.ul_data.num_volumes == 4
    assert ex.batches == 8


def test_unlabelled_set_larger_than_the_labelled_one_sets_the_epoch_length(monkeypatch):
    ex, _ = _executor(monkeypatch, 0.25, pairs=64)          # 8 volumes: 2 labelled, 6 unlabelled
    assert ex.data.size() == 16 and ex.ul_data.size() == 48
    assert ex.batches == 12                                  # dafnet_executor.py:110-111: the larger of the two sets


def test_training_csv_columns_are_the_reference_ones(monkeypatch):
    ex, _ = _executor(monkeypatch, 1.0)
    assert ex.get_loss_names() == ["adv_M", "adv_X1", "adv_X2", "rec_X", "dis_M", "dis_X1", "dis_X2",
                                   "val_loss", "val_loss_mod1", "val_loss_mod2",
                                   "val_loss_mod2_mod1def", "val_loss_mod1_mod2def", "val_loss_mod2_fused", "val_loss_mod1_fused",
                                   "val_weight_0", "val_weight_1", "val_weight_2",
                                   "supervised_Mask", "KL", "rec_Z"]                         # dafnet_executor.py:200-205
    ex, _ = _executor(monkeypatch, 1.0, kind="mmsdnet")
    assert ex.get_loss_names() == ["adv_M", "rec_X", "dis_M", "val_loss", "val_loss_mod1", "val_loss_mod2",
                                   "val_loss_mod2_s1def", "val_loss_mod2_fused", "supervised_Mask", "loss", "KL", "rec_Z"]   # mmsdnet_executor.py:155-157
