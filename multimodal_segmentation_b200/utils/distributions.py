"""Sampler of the modality codes z ~ N(0, I) that the executors feed to the decoder / Z-regressor
(reference: utils/distributions.py, call sites model_executors/dafnet_executor.py:497-499)."""
import numpy as np


class NormalDistribution(object):
    """Draws from numpy's GLOBAL generator, like the reference: a run seeded with ``np.random.seed(conf.seed)`` then
    consumes the stream in the reference's order (``sample(N)`` accepts an int or a shape tuple)."""

    def __init__(self, mu=0.0, sigma=1.0):
        self.mu, self.sigma = mu, sigma

    def sample(self, N):
        z = np.random.standard_normal(N)
        return z if (self.mu == 0 and self.sigma == 1) else self.mu + self.sigma * z
