"""Loss functions (reference: costs.py).  Losses used inside trainers are small descriptor objects
that the trainers lower onto the fused loss kernels; numpy metrics stay numpy (they are host-side
in the reference too)."""
import numpy as np

lambda_bce = 0.01


class SegLoss(object):
    """dice over the first `restrict_chn` channels (+ lambda_bce * weighted cross entropy with the
    reference's swapped arguments when use_bce) -- costs.py:43-85,129-136.

    The trainers lower the descriptor onto the fused loss kernels (value + gradient in one pass); called directly with
    the Keras loss signature ``(y_true, y_pred)`` on CUDA tensors it returns the loss as a 1-element device tensor."""

    def __init__(self, restrict_chn, use_bce):
        self.restrict_chn, self.use_bce = restrict_chn, use_bce

    def __call__(self, y_true, y_pred):
        from . import ops
        loss = ops.zeros(1)
        ops.segloss(y_pred.contiguous(), y_true.contiguous(), self.restrict_chn, self.use_bce, 1.0, loss, lambda_bce,
                    want_grad=False)
        return loss


# ---- functional forms of the reference's loss helpers on CUDA tensors [B, H, W, C] (same names, same argument order,
# ---- including the order quirks of costs.py; every one runs on the kernels of csrc/losses.cu / csrc/pairing.cu)
def _rows(pred, target, nch, lam):
    from . import ops
    L = ops.zeros(pred.shape[0])
    ops.segloss_pb_fwd(pred.contiguous(), target.contiguous(), nch, L, lam)
    return L


def dice_coef_perbatch(y_true, y_pred):
    """costs.py:43-48: 1 - dice per sample over (H, W, C), smooth 1e-12 -> [B]"""
    return _rows(y_pred, y_true, y_pred.shape[-1], 0.0)


def dice_coef_loss(y_true, y_pred):
    """costs.py:50-56: batch mean of dice_coef_perbatch"""
    return SegLoss(y_pred.shape[-1], False)(y_true, y_pred)


def weighted_cross_entropy_loss(y_pred, y_true):
    """costs.py:70-85, signature (y_pred, y_true): class counts from `y_true`, -sum_c y_true_c log(y_pred_c + 1e-12) w_c
    averaged over pixels.  The fused kernel computes dice + lambda * this term with `y_true` in the role the combined loss
    gives it (costs.py:134 passes the network output there), so the term alone is the difference of two launches."""
    from . import ops
    n = y_true.shape[-1]
    both, dice_only = ops.zeros(1), ops.zeros(1)
    ops.segloss(y_true.contiguous(), y_pred.contiguous(), n, True, 1.0, both, 1.0, want_grad=False)
    ops.segloss(y_true.contiguous(), y_pred.contiguous(), n, False, 1.0, dice_only, 1.0, want_grad=False)
    return both - dice_only


def weighted_cross_entropy_perbatch(y_pred, y_true):
    """costs.py:88-108, signature (y_pred, y_true): batch-wide class weights from `y_true`, softmax applied to `y_pred`,
    per-sample pixel mean -> [B]"""
    n = y_true.shape[-1]
    return _rows(y_true, y_pred, n, 1.0) - _rows(y_true, y_pred, n, 0.0)


def make_combined_dice_bce_perbatch(num_classes):
    """costs.py:138-143: per-sample dice over the first `num_classes` channels + lambda_bce * the per-batch cross entropy
    with swapped arguments -> [B]"""
    def fnc(y_true, y_pred):
        return _rows(y_pred, y_true, num_classes, lambda_bce)
    return fnc


def mae_single_input(y):
    """costs.py:24-26: mean |y1 - y2| over (H, W) per sample -> [B, 1] (single-channel images, as in models/dafnet.py)"""
    from . import ops
    y1, y2 = y
    assert y1.shape[-1] == 1, "mae_single_input: one image channel (the reference reduces over H and W only)"
    L = ops.zeros(y1.shape[0])
    ops.mae_pb_fwd(y1.contiguous(), y2.contiguous(), L)
    return L.reshape(-1, 1)


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
