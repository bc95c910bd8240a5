"""Golden vectors for utils/data_utils.crop_same, produced by the REFERENCE's own function (run in the build container, where
/root/reference exists):  python tests/golden/make_golden_crop.py  ->  tests/golden/golden_crop.npz"""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [((10, 12), (8, 8)), ((7, 9), (10, 12)), ((11, 8), (8, 11)), ((9, 9), (9, 9)), ((13, 6), (8, 9))]


def inputs(i, h, w):
    r = np.random.RandomState(100 + i)
    return r.normal(size=(3, h, w, 1)).astype(np.float32), (r.uniform(size=(3, h, w, 4)) > 0.5).astype(np.float32)


def main():
    spec = importlib.util.spec_from_file_location("ref_data_utils", "/root/reference/utils/data_utils.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    for i, ((h, w), size) in enumerate(CASES):
        for mode in ("equal", "left", "right"):
            for pad_mode in ("edge", "constant"):
                im, m = inputs(i, h, w)
                [a], [b] = ref.crop_same([im], [m], size=size, mode=mode, pad_mode=pad_mode)
                out["%d_%s_%s_image" % (i, mode, pad_mode)] = a
                out["%d_%s_%s_mask" % (i, mode, pad_mode)] = b
    np.savez_compressed(os.path.join(HERE, "golden_crop.npz"), **out)
    print("%d arrays" % len(out))


if __name__ == "__main__":
    main()
