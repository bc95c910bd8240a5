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
