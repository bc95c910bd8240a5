"""reference: configuration/dafnet_spade_config_chaos.py (decoder_type 'spade')"""
from . import dafnet_config_chaos as _base


def get(input_shape=None):
    return _base.get(input_shape=input_shape, decoder_type='spade', folder='dafnet_spade_chaos')
