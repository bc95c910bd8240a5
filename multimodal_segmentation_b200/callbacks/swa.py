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
    def __init__(self, swa_epoch, model_build_fnc, build_params):
        self.swa_epoch = swa_epoch
        self.model_build_fnc = model_build_fnc
        self.build_params = build_params
        self.clone = None
        self.model = None            # set by the executor (keras sets it through Callback.set_model)
        self.params = {}
        self.swa_weights = None

    def on_train_begin(self, logs=None):
        self.nb_epoch = self.params["epochs"]
        print("Stochastic weight averaging selected for last {} epochs.".format(self.nb_epoch - self.swa_epoch))

    def on_epoch_end(self, epoch, logs=None):
        if epoch <= self.swa_epoch:
            self.swa_weights = self.model.get_weights()
        elif epoch > self.swa_epoch:
            cur = self.model.get_weights()
            k = epoch - self.swa_epoch
            for i in range(len(self.swa_weights)):
                self.swa_weights[i] = (self.swa_weights[i] * k + cur[i]) / (k + 1)

    def on_train_end(self, logs=None):
        self.model.set_weights(self.swa_weights)
        log.debug("Final model parameters set to stochastic weight average.")

    def get_clone_model(self):
        if self.clone is None:
            with BuildScope(rng=np.random.RandomState(0)):
                if self.build_params is not None:
                    self.clone = self.model_build_fnc(self.build_params)
                else:
                    self.clone = self.model_build_fnc()
        self.clone.set_weights(self.swa_weights if self.swa_weights is not None else self.model.get_weights())
        return self.clone
