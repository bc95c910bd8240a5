"""reference: utils/sdnet_utils.py"""
import numpy as np

from .distributions import NormalDistribution


def vae_sample(args):
    z_mean, z_log_var = args
    batch, dim = z_mean.shape[0], z_mean.shape[1]
    epsilon = NormalDistribution().sample((batch, dim))
    return z_mean + np.exp(0.5 * z_log_var) * epsilon


def get_net(trainer_model, name):
    layers = [l for l in trainer_model.layers if l.name == name]
    assert len(layers) == 1
    return layers[0]


def make_trainable(model, val):
    """utils/sdnet_utils.py:40-53: (un)freeze every weight of a model"""
    model.trainable = val
