"""Golden vectors for the l_mix volume sampling, produced by the REFERENCE's own MultimodalPairedData / Data classes (run in the
build container, where /root/reference exists):  python tests/golden/make_golden_sampling.py -> tests/golden/golden_sampling.npz

The labelled set is Data.sample(round(l_mix * num_volumes), seed) (loaders/data.py:123-151 through MultimodalPairedData's
filter_volumes, loaders/MultimodalPairedData.py:46-62); the unlabelled set is what get_sample_volumes' same draw leaves out
(model_executors/dafnet_executor.py:136-142)."""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
N, PER_VOLUME = 43, 8            # five whole synthetic volumes and a short one
CASES = [(0.5, 3), (0.25, 10), (0.75, 4), (1.0, 3)]       # (l_mix, seed)


def inputs():
    tag = np.arange(N, dtype=np.float32).reshape(N, 1, 1, 1)
    images = np.concatenate([tag, tag + 1000], -1)
    masks = np.concatenate([np.broadcast_to(tag, (N, 1, 1, 4)), np.broadcast_to(tag + 2000, (N, 1, 1, 4))], -1).astype(np.float32)
    return images, masks, np.arange(N) // PER_VOLUME


def main():
    for name in ("skimage", "skimage.measure"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["skimage.measure"].block_reduce = None
    for name in ("loaders", "utils"):
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(REF, name)]
        sys.modules[name] = pkg
    sys.modules["utils.image_utils"] = types.ModuleType("utils.image_utils")     # imports albumentations; unused here

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    load("utils.data_utils", "utils/data_utils.py")
    load("loaders.data", "loaders/data.py")
    MPD = load("loaders.MultimodalPairedData", "loaders/MultimodalPairedData.py").MultimodalPairedData
    out = {}
    for l_mix, seed in CASES:
        images, masks, index = inputs()
        lab = MPD(images.copy(), masks.copy(), index.copy())
        num = int(np.round(l_mix * lab.num_volumes))
        lab.sample(num, seed=seed)
        key = "l%s_s%d" % (str(l_mix).replace(".", ""), seed)
        out[key + "_lab_images0"], out[key + "_lab_masks1"] = lab.get_images_modi(0), lab.get_masks_modi(1)
        out[key + "_lab_index"] = np.asarray(lab.index)
        ul = MPD(images.copy(), masks.copy(), index.copy())
        volumes = ul.get_sample_volumes(num, seed=seed)
        rest = [v for v in ul.volumes() if v not in volumes]
        if len(rest) > 0:
            ul.filter_volumes(rest)
            out[key + "_ul_images1"], out[key + "_ul_index"] = ul.get_images_modi(1), np.asarray(ul.index)
        else:
            out[key + "_ul_index"] = np.zeros((0,), np.int64)
    np.savez_compressed(os.path.join(HERE, "golden_sampling.npz"), **out)
    print("%d arrays" % len(out))


if __name__ == "__main__":
    main()
