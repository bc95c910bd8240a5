"""Host-side data path (no GPU): aligned batch flows with shared rotation angles (reference
model_executors/base_executor.py:37-78,103-110) and the candidate-pair expansion of the automated-pairing path
(loaders/MultimodalPairedData.py:91-141)."""
import numpy as np


def test_flows_with_one_seed_rotate_images_and_masks_together():
    from multimodal_segmentation_b200.model_executors.base_executor import BatchFlow, FlowGroup
    a = np.arange(40, dtype=np.float32).reshape(10, 2, 2, 1)
    g = FlowGroup([BatchFlow(a, 4, 10, 20.0), BatchFlow(a * 2, 4, 10, 20.0)])
    for _ in range(4):
        xa, xb = next(g)
        assert np.array_equal(xa.numpy() * 2, xb.numpy())                       # same order
        assert np.array_equal(g.flows[0].last_theta, g.flows[1].last_theta)     # same angles
        assert g.last_theta.shape == (xa.shape[0],) and np.abs(g.last_theta).max() <= np.deg2rad(20.0)
        g.mark_copied()
    assert BatchFlow(a, 4, 10).last_theta is None


def test_expand_pairs_keeps_the_expert_pair_first():
    from multimodal_segmentation_b200.loaders.synthetic_chaos import PairedData
    n = 19
    imgs = [np.arange(n, dtype=np.float32).reshape(n, 1, 1, 1) + 100 * m for m in range(2)]
    d = PairedData([imgs[0].copy(), imgs[1].copy()], [np.zeros((n, 1, 1, 4), np.float32)] * 2)
    np.random.seed(0)
    d.expand_pairs(2, 0, neighborhood=3)
    d.expand_pairs(2, 1, neighborhood=3)
    for m in range(2):
        x = d.get_images_modi(m)
        assert x.shape == (n, 1, 1, 3)
        assert np.array_equal(x[..., 0], imgs[m][..., 0])                  # channel 0 = the expert pair
        for i in range(n):
            a = (i // PairedData.SLICES_PER_VOLUME) * PairedData.SLICES_PER_VOLUME
            vol = range(a, min(n, a + PairedData.SLICES_PER_VOLUME))
            cands = x[i, 0, 0, 1:] - 100 * m
            assert all(int(c) in vol for c in cands)                       # neighbours come from the same volume
            if len(vol) >= 5:
                assert all(abs(int(c) - i) <= 4 for c in cands) and len(set(cands)) == 2 and i not in cands


def test_pair_expansion_and_repairing_match_the_reference_class():
    """golden vectors produced by the reference's own MultimodalPairedData.expand_pairs / randomise_pairs
    (tests/golden/make_golden.py): same candidates, same order, same consumption of numpy's global RNG"""
    import os
    from multimodal_segmentation_b200.loaders.synthetic_chaos import PairedData
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_ref.npz"))
    imgs, msks = G["pairs_images"], G["pairs_masks"]
    assert PairedData.SLICES_PER_VOLUME == 8      # the fixture's volume index is i // 8

    def fresh():
        return PairedData([imgs[..., 0:1].copy(), imgs[..., 1:2].copy()], [msks[..., :4].copy(), msks[..., 4:].copy()])

    d = fresh()
    np.random.seed(7)
    d.expand_pairs(2, 0, neighborhood=3)
    d.expand_pairs(2, 1, neighborhood=3)
    assert np.array_equal(d.get_images_modi(0), G["pairs_expand_mod0"])
    assert np.array_equal(d.get_images_modi(1), G["pairs_expand_mod1"])
    d = fresh()
    d.randomise_pairs(length=3, seed=11)
    assert np.array_equal(d.get_images_modi(0), G["pairs_randomise_images"])
    assert np.array_equal(d.get_masks_modi(0), G["pairs_randomise_masks"])
    assert np.array_equal(d.get_images_modi(1), imgs[..., 1:2])          # modality 1 stays in place


def test_l_mix_volume_sampling_partitions_the_training_volumes():
    """loaders/data.py:123-151 + dafnet_executor.py:87-88,136-142: Data.sample keeps round(l_mix * num_volumes) volumes drawn
    with the configuration's seed; the unlabelled set is what the SAME draw leaves out (filter_volumes keeps the slices of a
    volume together and in the order the volumes are listed)."""
    from multimodal_segmentation_b200.loaders.synthetic_chaos import PairedData
    n = 5 * PairedData.SLICES_PER_VOLUME + 3                       # five whole volumes and a short one
    tag = np.arange(n, dtype=np.float32).reshape(n, 1, 1, 1)

    def mk():
        return PairedData([tag.copy(), tag.copy() + 1000], [np.zeros((n, 1, 1, 4), np.float32)] * 2)

    lab = mk()
    assert lab.num_volumes == 6 and lab.volume_ids() == [0, 1, 2, 3, 4, 5] and lab.size() == n
    lab.sample(lab.num_volumes, seed=3)                            # everything labelled: untouched, no draw
    assert lab.size() == n
    num = int(np.round(0.5 * lab.num_volumes))
    lab.sample(num, seed=3)
    np.random.seed(3)
    want = np.random.choice([0, 1, 2, 3, 4, 5], size=num, replace=False)           # the reference's draw
    assert lab.num_volumes == num and list(dict.fromkeys(lab.index.tolist())) == want.tolist()
    assert np.array_equal(lab.images[1], lab.images[0] + 1000)                    # modalities stay paired
    for a, b in lab.volumes():                                                     # whole volumes, slices in order
        ids = lab.images[0][a:b, 0, 0, 0]
        assert len(set((ids // PairedData.SLICES_PER_VOLUME).tolist())) == 1 and np.all(np.diff(ids) == 1)
    ul = mk()
    labelled = set(ul.get_sample_volumes(num, seed=3).tolist())
    ul.filter_volumes([v for v in ul.volume_ids() if v not in labelled])
    assert labelled == set(want.tolist())
    got = sorted(lab.images[0].ravel().tolist() + ul.images[0].ravel().tolist())
    assert got == tag.ravel().tolist()                                             # a partition of the training slices
    ul.filter_volumes([])
    assert ul.size() == 0 and ul.num_volumes == 0 and ul.volumes() == []


def test_crop_same_matches_the_reference_function():
    """utils/data_utils.py:37-123 (validation data are cropped / padded to conf.input_shape, dafnet_executor.py:313): golden
    vectors from the reference's own crop_same (tests/golden/make_golden_crop.py), incl. its odd-difference behaviour
    ('equal' removes ceil(diff / 2) pixels from each side and the pad step puts one back) and min-valued constant padding."""
    import os
    from multimodal_segmentation_b200.utils.data_utils import crop_same
    from tests.golden.make_golden_crop import CASES, inputs
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_crop.npz"))
    for i, ((h, w), size) in enumerate(CASES):
        for mode in ("equal", "left", "right"):
            for pad_mode in ("edge", "constant"):
                im, m = inputs(i, h, w)
                [a], [b] = crop_same([im], [m], size=size, mode=mode, pad_mode=pad_mode)
                assert np.array_equal(a, G["%d_%s_%s_image" % (i, mode, pad_mode)])
                assert np.array_equal(b, G["%d_%s_%s_mask" % (i, mode, pad_mode)])
    from multimodal_segmentation_b200.loaders.synthetic_chaos import PairedData
    d = PairedData([np.ones((4, 10, 12, 1), np.float32), np.ones((4, 10, 12, 1), np.float32)], [np.zeros((4, 10, 12, 4), np.float32)] * 2)
    d.crop((8, 8))
    assert d.images[0].shape == (4, 8, 8, 1) and d.masks[1].shape == (4, 8, 8, 4)


def test_volume_sampling_matches_the_reference_classes():
    """golden vectors from the reference's own MultimodalPairedData.sample / get_sample_volumes / filter_volumes
    (tests/golden/make_golden_sampling.py): the labelled slices for l_mix in {0.25, 0.5, 0.75, 1} and the unlabelled complement
    that dafnet_executor.py:136-142 builds from the same draw, slice by slice and in the reference's order."""
    import os
    from multimodal_segmentation_b200.loaders.synthetic_chaos import PairedData
    from tests.golden.make_golden_sampling import CASES, PER_VOLUME, inputs
    assert PairedData.SLICES_PER_VOLUME == PER_VOLUME
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_sampling.npz"))

    def mk():
        images, masks, index = inputs()
        return PairedData([images[..., 0:1].copy(), images[..., 1:2].copy()], [masks[..., :4].copy(), masks[..., 4:].copy()], index.copy())

    for l_mix, seed in CASES:
        key = "l%s_s%d" % (str(l_mix).replace(".", ""), seed)
        lab = mk()
        num = int(np.round(l_mix * lab.num_volumes))
        lab.sample(num, seed=seed)
        assert np.array_equal(lab.get_images_modi(0), G[key + "_lab_images0"])
        assert np.array_equal(lab.get_masks_modi(1), G[key + "_lab_masks1"])
        assert np.array_equal(lab.index, G[key + "_lab_index"])
        ul = mk()
        labelled = set(np.asarray(ul.get_sample_volumes(num, seed=seed)).tolist())
        ul.filter_volumes([v for v in ul.volume_ids() if v not in labelled])
        assert np.array_equal(ul.index, G[key + "_ul_index"])
        if ul.size() > 0:
            assert np.array_equal(ul.get_images_modi(1), G[key + "_ul_images1"])
