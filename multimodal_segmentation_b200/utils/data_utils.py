"""reference: utils/data_utils.py (the functions used inside the training step)"""
import numpy as np


def rescale(array, min_value=-1, max_value=1):
    """utils/data_utils.py:7-20: per-array rescale to [min_value, max_value]"""
    if array.max() == array.min():
        array = (array * 0) + min_value
        return array
    array = (max_value - min_value) * (array - float(array.min())) / (array.max() - array.min()) + min_value
    assert array.max() == max_value and array.min() == min_value, "%d, %d" % (array.max(), array.min())
    return array


def sample_indices(n, nb_samples, seed=-1):
    """the index draw of utils/data_utils.py:125-129 (np.random.choice without replacement)"""
    if seed > -1:
        np.random.seed(seed)
    return np.random.choice(n, size=nb_samples, replace=False)


def sample(data, nb_samples, seed=-1):
    idx = sample_indices(len(data), nb_samples, seed)
    return np.array([data[i] for i in idx])


def _crop_axis(a, axis, size, mode):
    """utils/data_utils.py:82-101.  'equal' removes ceil(diff / 2) pixels from EACH side, so an odd difference leaves
    size - 1 pixels (the reference's behaviour, kept: crop_same's pad step then puts one pixel back behind)"""
    diff = a.shape[axis] - size
    if mode == "equal":
        lo = int(np.ceil(diff / 2))
        hi = a.shape[axis] - lo
    elif mode == "right":
        lo, hi = 0, size
    elif mode == "left":
        lo, hi = diff, a.shape[axis]
    else:
        raise ValueError("Unexpected mode: %s. Expected to be one of [equal, left, right]." % mode)
    index = [slice(None)] * a.ndim
    index[axis] = slice(lo, hi)
    return a[tuple(index)]


def _pad_axis(a, axis, size, mode):
    """utils/data_utils.py:104-123: floor(diff / 2) pixels in front, the rest behind; 'constant' pads with the array's minimum"""
    diff = size - a.shape[axis]
    lo = int(diff / 2)
    width = [(0, 0)] * a.ndim
    width[axis] = (lo, int(diff - lo))
    if mode == "edge":
        return np.pad(a, width, "edge")
    if mode == "constant":
        return np.pad(a, width, "constant", constant_values=np.min(a))
    raise Exception("Invalid pad mode: " + mode)


def crop_same(image_list, mask_list, size=(None, None), mode="equal", pad_mode="edge"):
    """utils/data_utils.py:37-79: bring every (image, mask) pair of 4-d arrays [slices, h, w, channels] to size[0] x size[1]
    (default: the smallest mask extent) by cropping or padding axis 1, then axis 2."""
    target = [np.min([m.shape[1] for m in mask_list]) if size[0] is None else size[0],
              np.min([m.shape[2] for m in mask_list]) if size[1] is None else size[1]]
    images, masks = [], []
    for im, m in zip(image_list, mask_list):
        for axis, want in ((1, target[0]), (2, target[1])):
            out = []
            for a in (m, im):
                if a.shape[axis] > want:
                    a = _crop_axis(a, axis, want, mode)
                if a.shape[axis] < want:
                    a = _pad_axis(a, axis, want, pad_mode)
                out.append(a)
            m, im = out
        images.append(im)
        masks.append(m)
    return images, masks
