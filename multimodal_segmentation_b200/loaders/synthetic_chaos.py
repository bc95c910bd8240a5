"""Synthetic CHAOS-shaped paired T1/T2 data (shapes and value ranges of loaders/chaos.py:25-26,242-246:
(H,W,1) slices rescaled per slice to exactly [-1,1], 4 binary organ masks).

Images: a smooth random field plus a few constant-intensity ellipses; T2 = a small smooth warp of the
same ellipses with a different intensity map, so that registration has signal.  Masks: 4 disjoint
binary ellipse channels.  Deterministic given the seed.
"""
import os

import numpy as np

from ..utils.data_utils import rescale

DEFAULT_SHAPE = (192, 192, 1)     # loaders/chaos.py:26


def _smooth(rng, h, w, k=9):
    f = rng.normal(size=(h + 2 * k, w + 2 * k))
    ker = np.ones(k) / k
    for _ in range(2):
        f = np.apply_along_axis(lambda m: np.convolve(m, ker, mode="same"), 0, f)
        f = np.apply_along_axis(lambda m: np.convolve(m, ker, mode="same"), 1, f)
    return f[k:-k, k:-k]


def make_pairs(n, shape, num_masks=4, seed=10):
    """-> x1[n,H,W,1], x2[n,H,W,1] in [-1,1] (float32), m1, m2 [n,H,W,num_masks] in {0,1}"""
    rng = np.random.RandomState(seed)
    H, W = shape[0], shape[1]
    yy, xx = np.mgrid[:H, :W].astype(np.float32)
    x1 = np.zeros((n, H, W, 1), np.float32)
    x2 = np.zeros((n, H, W, 1), np.float32)
    m1 = np.zeros((n, H, W, num_masks), np.float32)
    m2 = np.zeros((n, H, W, num_masks), np.float32)
    for i in range(n):
        img1 = 0.6 * _smooth(rng, H, W)
        img2 = 0.6 * _smooth(rng, H, W)
        occupied1 = np.zeros((H, W), bool)
        occupied2 = np.zeros((H, W), bool)
        shift = rng.uniform(-0.03, 0.03, size=2) * np.array([H, W])
        for c in range(num_masks):
            cy, cx = rng.uniform(0.25, 0.75) * H, rng.uniform(0.25, 0.75) * W
            ry, rx = rng.uniform(0.06, 0.16) * H, rng.uniform(0.06, 0.16) * W
            e1 = (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2) <= 1.0
            e2 = (((yy - cy - shift[0]) / (ry * 1.05)) ** 2 + ((xx - cx - shift[1]) / (rx * 0.95)) ** 2) <= 1.0
            e1 &= ~occupied1
            e2 &= ~occupied2
            occupied1 |= e1
            occupied2 |= e2
            m1[i, ..., c] = e1
            m2[i, ..., c] = e2
            img1[e1] = rng.uniform(-1, 1)
            img2[e2] = rng.uniform(-1, 1)
        x1[i, ..., 0] = rescale(img1)
        x2[i, ..., 0] = rescale(img2)
    return x1, x2, m1, m2


class PairedData(object):
    """the slice of loaders/MultimodalPairedData.py + loaders/data.py used by the executors: two modalities' images and
    masks, and `index`, the volume every slice belongs to (synthetic "volumes" = runs of SLICES_PER_VOLUME items)"""

    SLICES_PER_VOLUME = 8

    def __init__(self, images, masks, index=None):
        self.images = images          # list of two arrays
        self.masks = masks
        n = images[0].shape[0]
        self.index = np.arange(n) // self.SLICES_PER_VOLUME if index is None else np.asarray(index)
        self.num_volumes = len(self.volume_ids())

    def size(self):
        return max(im.shape[0] for im in self.images)      # loaders/MultimodalPairedData.py:74-75

    def volume_ids(self):
        """loaders/data.py `volumes()`: the sorted set of volume ids"""
        return sorted(set(self.index.tolist()))

    def get_sample_volumes(self, num, seed=-1):
        """loaders/data.py:123-129: `num` volume ids drawn without replacement from numpy's global generator"""
        if seed > -1:
            np.random.seed(seed)
        return np.random.choice(self.volume_ids(), size=num, replace=False)

    def sample(self, nb_samples, seed=-1):
        """loaders/data.py:131-136: keep `nb_samples` randomly chosen volumes (the labelled share l_mix of the data)"""
        if nb_samples == self.num_volumes:
            return
        self.filter_volumes(self.get_sample_volumes(nb_samples, seed))

    def filter_volumes(self, volumes):
        """loaders/MultimodalPairedData.py:46-62: keep the slices of the listed volumes, in the order listed"""
        volumes = list(volumes)
        if len(volumes) == 0:
            self.images = [im[0:0] for im in self.images]
            self.masks = [m[0:0] for m in self.masks]
            self.index = self.index[0:0]
            self.num_volumes = 0
            return
        keep = np.concatenate([np.nonzero(self.index == v)[0] for v in volumes], axis=0)
        self.images = [im[keep] for im in self.images]
        self.masks = [m[keep] for m in self.masks]
        self.index = self.index[keep]
        self.num_volumes = len(volumes)

    def crop(self, shape):
        """loaders/MultimodalPairedData.py:64-72: images and masks of every modality cropped / padded to `shape`"""
        from ..utils.data_utils import crop_same
        for i in range(len(self.images)):
            [self.images[i]], [self.masks[i]] = crop_same([self.images[i]], [self.masks[i]], size=shape, pad_mode="constant")
            assert self.images[i].shape[1:-1] == self.masks[i].shape[1:-1] == tuple(shape), \
                "Invalid shapes: %s %s %s" % (self.images[i].shape[1:-1], self.masks[i].shape[1:-1], shape)

    def get_images_modi(self, i):
        return self.images[i]

    def get_masks_modi(self, i):
        return self.masks[i]

    def volumes(self):
        """[a, b) item ranges of the volumes, in storage order (filter_volumes keeps a volume's slices together)"""
        n = self.index.shape[0]
        if n == 0:
            return []
        cuts = [0] + [i for i in range(1, n) if self.index[i] != self.index[i - 1]] + [n]
        return list(zip(cuts[:-1], cuts[1:]))

    def randomise_pairs(self, length=3, seed=None):
        """loaders/MultimodalPairedData.py:143-166: re-pair modality 0 with a slice up to `length` positions away
        inside the same volume (images and masks of modality 0 move together)"""
        if seed is not None:
            np.random.seed(seed)
        new_images, new_masks = [], []
        for a, b in self.volumes():
            images, masks = self.images[0][a:b], self.masks[0][a:b]
            n = images.shape[0]
            offsets = np.random.randint(-length, length, size=n)
            for off in range(min(length, n)):
                if offsets[off] + off < 0:
                    offsets[off] = np.random.randint(-off, length, size=1)[0]
            for i in range(1, min(length, n)):
                if offsets[-i] + (n - i) >= n:
                    offsets[-i] = np.random.randint(-length, i, size=1)[0]
            idx = np.clip(np.arange(n) + offsets, 0, n - 1)
            new_images.append(images[idx])
            new_masks.append(masks[idx])
        self.images[0] = np.concatenate(new_images, axis=0)
        self.masks[0] = np.concatenate(new_masks, axis=0)

    def expand_pairs(self, offsets, mod_i, neighborhood=2):
        """loaders/MultimodalPairedData.py:91-141: every image of modality `mod_i` becomes `neighborhood` candidate
        images stacked on the channel axis -- channel 0 the expertly paired slice, the others drawn without
        replacement from the 2*offsets neighbouring slices of the same volume.  Consumes numpy's global RNG exactly as
        the reference does (one `choice` per slice, only when the window is larger than the neighbourhood); pinned
        against the reference's own output in tests/golden/golden_ref.npz."""
        assert mod_i in [0, 1], "mod_i can be in [0, 1]. It defines the neighborhood of which modality to enlarge"
        width = 2 * offsets + 1
        stacked = []
        for a, b in self.volumes():
            src = self.images[mod_i][a:b]
            n_src, n_dst = src.shape[0], self.images[1 - mod_i][a:b].shape[0]
            for i in range(n_dst):
                if n_src < width:                      # volume shorter than the window: padded with slice 0
                    window = list(range(n_src)) + [0] * (width - n_src)
                else:                                  # `width` consecutive slices around i, clamped to the volume
                    start = min(max(i - offsets, 0), n_dst - width)
                    window = list(range(start, start + width))
                window.remove(i)                       # the expertly paired slice always leads
                if width > neighborhood:
                    window = list(np.random.choice(window, size=neighborhood - 1, replace=False))
                stacked.append(np.concatenate([src[k] for k in [i] + window], axis=-1))
        all_images = np.stack(stacked, axis=0)
        assert all_images.shape[-1] == neighborhood, "%s vs %s" % (all_images.shape[-1], neighborhood)
        self.images[mod_i] = all_images


class SyntheticChaosLoader(object):
    def __init__(self):
        shp = os.environ.get("DAFK_INPUT_SHAPE")
        self.input_shape = tuple(int(v) for v in shp.split("x")) if shp else DEFAULT_SHAPE
        self.num_masks = 4
        self.modalities = ["t1", "t2"]
        self.num_pairs = {"training": int(os.environ.get("DAFK_TRAIN_PAIRS", "264")), "validation": 32, "test": 32}

    def load_all_modalities_concatenated(self, split, split_type, downsample=1, seed=10):
        n = self.num_pairs.get(split_type, 32)
        off = {"training": 0, "validation": 1, "test": 2}.get(split_type, 3)
        x1, x2, m1, m2 = make_pairs(n, self.input_shape, self.num_masks, seed=seed + 1000 * off + int(split))
        return PairedData([x1, x2], [m1, m2])

    def load_labelled_data(self, split, split_type, modality, downsample=1):
        d = self.load_all_modalities_concatenated(split, split_type, downsample)
        i = self.modalities.index(modality) if modality in self.modalities else 0
        return PairedData([d.images[i], d.images[i]], [d.masks[i], d.masks[i]])
