"""Stochastic Weight Averaging callback (reference: callbacks/swa.py:16-47, used by
model_executors/dafnet_executor.py:41-68,207-210,240-242,268-301,319-325).

One SWA object per component: up to and including ``swa_epoch`` it tracks the component's current weights, afterwards
the running average  w_swa <- (w_swa * (epoch - swa_epoch) + w) / (epoch - swa_epoch + 1).  ``get_clone_model()``
builds a second instance of the component with the averaged weights (the executors validate and save through the
clones), ``on_train_end()`` writes the average back into the live component.  The averaging runs once per epoch on the
host over the Keras-ordered weight list (``Model.get_weights`` / ``set_weights``), BatchNorm moving statistics included,
exactly as the reference does; nothing of it is on the step's hot path.
"""
import logging

import numpy as np

from ..keras_like import BuildScope

log = logging.getLogger("swa")


class SWA(object):
    """Same constructor and Keras-callback surface as the reference's ``SWA`` (``swa_epoch``, ``model_build_fnc``,
    ``build_params``; ``model`` / ``params`` are filled in by the executor, as ``Callback.set_model`` / ``set_params``
    would)."""

    def __init__(self, swa_epoch, model_build_fnc, build_params):
        self.swa_epoch = swa_epoch
        self.model_build_fnc = model_build_fnc
        self.build_params = build_params
        self.model = None
        self.params = {}
        self.clone = None
        self.swa_weights = None      # Keras-ordered list of host arrays

    def set_model(self, model):
        self.model = model

    def set_params(self, params):
        self.params = dict(params)

    def on_train_begin(self, logs=None):
        self.nb_epoch = self.params["epochs"]
        # the reference announces this on stdout as well
        print("SWA: weights are averaged over the last %d epochs (of %d)" % (self.nb_epoch - self.swa_epoch, self.nb_epoch))

    def on_epoch_end(self, epoch, logs=None):
        current = self.model.get_weights()
        averaged = epoch - self.swa_epoch            # epochs that are already part of the average
        if averaged <= 0 or self.swa_weights is None:
            self.swa_weights = current               # before the averaging window: follow the live weights
            return
        # callbacks/swa.py:31-33, evaluated in the same order so that the fp32 result is the reference's
        self.swa_weights = [(mean * averaged + w) / (averaged + 1) for mean, w in zip(self.swa_weights, current)]

    def on_train_end(self, logs=None):
        self.model.set_weights(self.swa_weights)
        log.debug("live component overwritten with its stochastic weight average")

    def get_clone_model(self):
        """a second instance of the component (built once) carrying the averaged weights"""
        if self.clone is None:
            args = () if self.build_params is None else (self.build_params,)
            with BuildScope(rng=np.random.RandomState(0)):
                self.clone = self.model_build_fnc(*args)
        self.clone.set_weights(self.swa_weights if self.swa_weights is not None else self.model.get_weights())
        return self.clone
