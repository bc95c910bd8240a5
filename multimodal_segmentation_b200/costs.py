"""Loss functions (reference: costs.py).  Losses used inside trainers are small descriptor objects
that the trainers lower onto the fused loss kernels; numpy metrics stay numpy (they are host-side
in the reference too)."""
import numpy as np

lambda_bce = 0.01


class SegLoss(object):
    """dice over the first `restrict_chn` channels (+ lambda_bce * weighted cross entropy with the
    reference's swapped arguments when use_bce) -- costs.py:43-85,129-136"""

    def __init__(self, restrict_chn, use_bce):
        self.restrict_chn, self.use_bce = restrict_chn, use_bce


def make_dice_loss_fnc(restrict_chn=1):
    return SegLoss(restrict_chn, False)


def make_combined_dice_bce(num_classes):
    return SegLoss(num_classes, True)


def ypred(y_true, y_pred):
    return y_pred


def dice(y_true, y_pred, binarise=False, smooth=1e-12):
    """costs.py:31-41 (numpy)"""
    y_pred = y_pred[..., 0:y_true.shape[-1]]
    if binarise:
        y_pred = np.round(y_pred)
    y_int = y_true * y_pred
    return np.mean((2 * np.sum(y_int, axis=(1, 2, 3)) + smooth)
                   / (np.sum(y_true, axis=(1, 2, 3)) + np.sum(y_pred, axis=(1, 2, 3)) + smooth))


def kl(args):
    """costs.py:186-189 on numpy arrays"""
    mean, log_var = args
    return (-0.5 * np.sum(1 + log_var - np.square(mean) - np.exp(log_var), axis=-1)).reshape(-1, 1)
