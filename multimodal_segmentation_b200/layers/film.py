"""FiLM conditioning (reference: layers/film.py:6-39).

    h = FiLM()([h, gamma, beta])      h: (B,H,W,C)   gamma, beta: (B,C)

The reference tiles gamma/beta to full resolution with K.tile and then multiplies; here one
bandwidth-bound kernel reads x once and broadcasts the (B,C) factors from L1/L2.
"""
from .. import engine as E


class FiLM(object):
    def __init__(self, **kwargs):
        self.name = kwargs.get("name", "film")

    def __call__(self, ctx, x):
        h, gamma, beta = x
        return E.film(ctx, h, gamma, beta)

    def compute_output_shape(self, input_shape):
        return input_shape
