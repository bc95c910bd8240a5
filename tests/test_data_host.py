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
