"""Spectral kernel regulariser (reference: layers/spectralnorm.py:199-246).

Spectral(dim, alpha): three power iterations on W reshaped to [dim, cout], restarted every call from
the initial random u0 ~ U(-1,1) (the reference rebinds self.u to a graph tensor, so it is never
persisted); loss = alpha * mean|stop_gradient(W/sigma) - W|.
"""
import numpy as np
import torch

from .. import ops


class Spectral(object):
    def __init__(self, dim, alpha=10., rng=None):
        self.dim = dim
        self.alpha = float(alpha)
        rng = rng if rng is not None else np.random
        self.u0_host = (rng.random_sample((dim, 1)) * 2 - 1.).astype(np.float32)
        self._u0 = None

    def u0(self):
        if self._u0 is None:
            self._u0 = torch.from_numpy(self.u0_host).cuda()
        return self._u0

    def __call__(self, kernel_param, loss_slot, with_grad=True):
        """adds alpha*mean|W/sigma - W| to loss_slot[0] and its gradient to kernel_param.grad"""
        w2 = kernel_param.data.view(-1, kernel_param.shape[-1])
        assert w2.shape[0] == self.dim, (w2.shape, self.dim)
        dW = kernel_param.grad.view(-1, kernel_param.shape[-1]) if with_grad else None
        ops.spectral_reg(w2, self.u0(), self.alpha, loss_slot, dW)

    def get_config(self):
        return {"alpha": float(self.alpha)}
