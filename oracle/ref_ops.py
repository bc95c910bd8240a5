"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU restatement (torch-CPU / numpy, fp32 or fp64) of every TF/Keras op on the MMSDNet / DAFNet
hot path, following the reference call sites cited on each function and the pinned third-party
semantics listed in SURVEY.md Appendix A (Keras 2.1.6, tensorflow 1.4.0, keras-contrib 2.0.8 --
none of which are vendored in /root/reference or installable here).

PARITY PINNING: the reference has no tests, golden vectors or fixtures, and its TF 1.4 stack
cannot run in this image.  What pins this oracle:
  * the reference's OWN Python source, imported unmodified from /root/reference on numpy stand-ins
    for the TF / Keras entry points it calls (tests/golden/make_golden.py, tf_shim.py,
    keras_graph.py), produced the committed golden vectors:
      - golden_ref.npz: the custom layers, losses and host-side data functions
        (tests/test_oracle_golden.py, tests/test_data_host.py, tests/test_swa.py);
      - golden_builders.npz: every component builder, the DAFNet (expert / unsupervised /
        automated-pairing) and MMSDNet generator graphs with their compiled losses and weights,
        predict_mask, the discriminator trainer (tests/test_oracle_builders.py);
  * for the arithmetic that lives INSIDE TF / Keras ops (Conv2D, BatchNormalization statistics,
    pooling, Adam, the resampler), which no reference artefact can pin here ("parity unpinned"
    for those): known-answer tests, independent cross-oracles (scipy RBFInterpolator, torch
    grid_sample, np.round, a second pure-numpy loop implementation of conv / BN / pooling) and
    fp64 finite-difference checks (tests/test_oracle_kat.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.

All tensors are NHWC.  Functions take and return torch tensors and are differentiable through
torch autograd with the reference's gradient conventions (custom Functions where TF differs from
torch: LeakyReLU'(0)=0, tf.maximum ties, straight-through rounding).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

EPSILON_SPLINE = 0.0000000001  # layers/interpolate_spline.py:26


# ----------------------------------------------------------------------------- basic layers
def conv2d(x, w, b=None, stride=1, padding="valid"):
    """keras Conv2D (A1): NHWC input, HWIO kernel, cross-correlation; 'same' only used with stride 1
    (models/unet.py:95; model_components/segmentor.py:15; modality_encoder.py:36 is 'valid' stride 2)."""
    kh, kw = w.shape[0], w.shape[1]
    if padding == "same":
        assert stride == 1 and kh % 2 == 1 and kw % 2 == 1
        pad = (kh // 2, kw // 2)
    else:
        pad = (0, 0)
    y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1)


def conv2d_loops(x, w, b=None, stride=1, pad=0):
    """second, independent numpy-loop statement of the same convolution (small shapes only)."""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    n, h, ww, ci = x.shape
    kh, kw, _, co = w.shape
    ho = (h + 2 * pad - kh) // stride + 1
    wo = (ww + 2 * pad - kw) // stride + 1
    xp = np.zeros((n, h + 2 * pad, ww + 2 * pad, ci))
    xp[:, pad:pad + h, pad:pad + ww] = x
    y = np.zeros((n, ho, wo, co))
    for i in range(ho):
        for j in range(wo):
            patch = xp[:, i * stride:i * stride + kh, j * stride:j * stride + kw, :]
            y[:, i, j, :] = np.tensordot(patch, w, axes=([1, 2, 3], [0, 1, 2]))
    if b is not None:
        y = y + np.asarray(b, np.float64)
    return y


def dense(x, w, b=None):
    """keras Dense, kernel [in, out] (modality_encoder.py:46-50)."""
    y = x @ w
    return y if b is None else y + b


def batchnorm_train(x, gamma, beta, eps=1e-3):
    """keras BatchNormalization() in training (A2): batch mean and biased variance over N,H,W.
    Returns (y, mean, biased_var)."""
    mean = x.mean(dim=(0, 1, 2))
    var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
    y = (x - mean) * torch.rsqrt(var + eps) * gamma + beta
    return y, mean, var


def batchnorm_infer(x, gamma, beta, moving_mean, moving_var, eps=1e-3):
    return (x - moving_mean) * torch.rsqrt(moving_var + eps) * gamma + beta


def bn_moving_update(moving_mean, moving_var, mean, var, count, momentum=0.99):
    """moving <- moving*m + batch*(1-m); the moving variance receives the Bessel-corrected variance."""
    unb = var * (count / max(count - 1, 1))
    return moving_mean * momentum + mean * (1 - momentum), moving_var * momentum + unb * (1 - momentum)


class _LeakyReLU(torch.autograd.Function):
    """keras 2.1.6 LeakyReLU = relu(x) - alpha*relu(-x): derivative at exactly 0 is 0 (A3)."""

    @staticmethod
    def forward(ctx, x, alpha):
        ctx.save_for_backward(x)
        ctx.alpha = alpha
        return torch.where(x > 0, x, alpha * x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        d = torch.where(x > 0, torch.ones_like(x), torch.where(x < 0, torch.full_like(x, ctx.alpha), torch.zeros_like(x)))
        return g * d, None


def leaky_relu(x, alpha=0.3):
    return _LeakyReLU.apply(x, alpha)


def relu(x):
    return torch.relu(x)  # derivative 0 at 0 in both TF and torch


def softmax(x):
    return torch.softmax(x, dim=-1)


def maxpool2(x):
    """MaxPooling2D(2,2) 'valid' (A4); torch routes the gradient to the first maximum like TF."""
    return F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)


def upsample2(x):
    """UpSampling2D(2): nearest repeat (utils/model_utils.py:16)."""
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)


def resize_nn(x, ho, wo):
    """tf.image.resize_nearest_neighbor, align_corners=False: src = floor(dst*in/out) (layers/spade.py:36-38)."""
    h, w = x.shape[1], x.shape[2]
    iy = torch.clamp((torch.arange(ho) * h) // ho, max=h - 1)
    ix = torch.clamp((torch.arange(wo) * w) // wo, max=w - 1)
    return x[:, iy][:, :, ix]


class _Round(torch.autograd.Function):
    """layers/rounding.py:33-42: np.round (half to even) forward, identity gradient."""

    @staticmethod
    def forward(ctx, x):
        return torch.from_numpy(np.round(x.detach().numpy())).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g * 1


def rounding(x):
    return _Round.apply(x)


def film(x, gamma, beta):
    """layers/film.py:26-36."""
    return x * gamma[:, None, None, :] + beta[:, None, None, :]


def instance_norm_axis_none(x, eps=1e-3):
    """keras_contrib InstanceNormalization(axis=None, scale=False, center=False) (A6):
    statistics over H,W,C jointly per sample, (x-mean)/(std+eps)."""
    mean = x.mean(dim=(1, 2, 3), keepdim=True)
    std = torch.sqrt(((x - mean) ** 2).mean(dim=(1, 2, 3), keepdim=True))
    return (x - mean) / (std + eps)


def spade_cond(x, gamma, beta):
    """layers/spade.py:52-55."""
    return x * (1 + gamma) + beta


class _TFMaximum(torch.autograd.Function):
    """tf.maximum (A5): where a >= b the whole gradient goes to a (ties -> first input)."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return torch.maximum(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        m = (a >= b).to(g.dtype)
        return g * m, g * (1 - m)


def tf_maximum(a, b):
    return _TFMaximum.apply(a, b)


# ----------------------------------------------------------------------------- polyharmonic spline
def _phi(r, order):
    """layers/interpolate_spline.py:182-209."""
    eps = torch.tensor(EPSILON_SPLINE, dtype=r.dtype)
    if order == 1:
        return torch.sqrt(torch.maximum(r, eps))
    if order == 2:
        return 0.5 * r * torch.log(torch.maximum(r, eps))
    if order == 4:
        return 0.5 * r * r * torch.log(torch.maximum(r, eps))
    if order % 2 == 0:
        r = torch.maximum(r, eps)
        return 0.5 * torch.pow(r, 0.5 * order) * torch.log(r)
    r = torch.maximum(r, eps)
    return torch.pow(r, 0.5 * order)


def _cross_squared_distance_matrix(x, y):
    """interpolate_spline.py:30-52."""
    xn = (x * x).sum(2)
    yn = (y * y).sum(2)
    xy = x @ y.transpose(1, 2)
    return xn[:, :, None] - 2 * xy + yn[:, None, :]


def _pairwise_squared_distance_matrix(x):
    """interpolate_spline.py:55-73."""
    xx = x @ x.transpose(1, 2)
    xn = torch.diagonal(xx, dim1=1, dim2=2)
    return xn[:, :, None] - 2 * xx + xn[:, None, :]


def solve_interpolation(train_points, train_values, order, regularization_weight=0.0):
    """interpolate_spline.py:76-147 -> (w [b,n,k], v [b,d+1,k])."""
    b, n, d = train_points.shape
    k = train_values.shape[-1]
    c, f = train_points, train_values
    matrix_a = _phi(_pairwise_squared_distance_matrix(c), order)
    if regularization_weight > 0:
        matrix_a = matrix_a + regularization_weight * torch.eye(n, dtype=c.dtype)[None]
    ones = torch.ones_like(c[..., :1])
    matrix_b = torch.cat([c, ones], 2)
    left_block = torch.cat([matrix_a, matrix_b.transpose(1, 2)], 1)
    lhs_zeros = torch.zeros(b, d + 1, d + 1, dtype=c.dtype)
    right_block = torch.cat([matrix_b, lhs_zeros], 1)
    lhs = torch.cat([left_block, right_block], 2)
    rhs = torch.cat([f, torch.zeros(b, d + 1, k, dtype=c.dtype)], 1)
    w_v = torch.linalg.solve(lhs, rhs)  # tf.matrix_solve: LU with partial pivoting
    return w_v[:, :n, :], w_v[:, n:, :]


def apply_interpolation(query_points, train_points, w, v, order):
    """interpolate_spline.py:150-179."""
    pairwise = _cross_squared_distance_matrix(query_points, train_points)
    rbf_term = _phi(pairwise, order) @ w
    qpad = torch.cat([query_points, torch.ones_like(query_points[..., :1])], 2)
    return rbf_term + qpad @ v


def interpolate_spline(train_points, train_values, query_points, order, regularization_weight=0.0):
    """interpolate_spline.py:212-278."""
    w, v = solve_interpolation(train_points, train_values, order, regularization_weight)
    return apply_interpolation(query_points, train_points, w, v, order)


def nDgrid(dims, dtype=torch.float32):
    """layers/stn_spline.py:70-91 (2-D, normalise=True, center=False): float32 cast of the fp64 grid."""
    g = np.expand_dims(np.mgrid[:dims[0], :dims[1]].reshape((2, -1)).T, 0)
    g = g / (1.0 * (np.array([[dims]]) - 1))
    return torch.from_numpy(g.astype(np.float32)).to(dtype)


def resampler(data, warp):
    """tf.contrib.resampler.resampler (A7): data [B,H,W,C], warp [B,M,2] = (x,y) pixels; bilinear,
    zero outside; differentiable w.r.t. data and warp with the analytic TF gradient."""
    B, H, W, C = data.shape
    x, y = warp[..., 0], warp[..., 1]
    valid = (x > -1) & (y > -1) & (x < W) & (y < H)
    fx, fy = torch.floor(x), torch.floor(y)
    cx, cy = fx + 1, fy + 1
    dx, dy = cx - x, cy - y

    def get(ix, iy):
        inb = (ix >= 0) & (ix <= W - 1) & (iy >= 0) & (iy <= H - 1)
        ixc = ix.clamp(0, W - 1).long()
        iyc = iy.clamp(0, H - 1).long()
        flat = data.reshape(B, H * W, C)
        idx = (iyc * W + ixc)[..., None].expand(-1, -1, C)
        return torch.gather(flat, 1, idx) * inb[..., None].to(data.dtype)

    out = (dx * dy)[..., None] * get(fx, fy) + ((1 - dx) * (1 - dy))[..., None] * get(cx, cy) \
        + (dx * (1 - dy))[..., None] * get(fx, cy) + ((1 - dx) * dy)[..., None] * get(cx, fy)
    return out * valid[..., None].to(data.dtype)


def thin_plate_spline_2d(vol, cp_offsets, cp_dims=(5, 5), order=2, inverse=False):
    """layers/stn_spline.py:38-67 ThinPlateSpline2D.call, literally: one spline fit per sample
    (tf.map_fn), reverse (row,col)->(x,y), scale by [W-1, H-1], bilinear resample."""
    B, H, W, C = vol.shape
    dt = vol.dtype
    cp_grid = nDgrid(cp_dims, dt)          # [1,25,2]
    flt_grid = nDgrid((H, W), dt)          # [1,H*W,2]
    locs = []
    for b in range(B):
        warped = cp_grid + cp_offsets[b:b + 1]
        if inverse:
            locs.append(interpolate_spline(warped, cp_grid, flt_grid, order))
        else:
            locs.append(interpolate_spline(cp_grid, warped, flt_grid, order))
    locs = torch.cat(locs, 0)              # [B,m,2] (row, col) normalised
    locs = torch.flip(locs, dims=[-1])     # -> (x, y)
    locs = locs * torch.tensor([W - 1, H - 1], dtype=dt)
    warped_vol = resampler(vol, locs)
    return warped_vol.reshape(B, H, W, C), locs


# ----------------------------------------------------------------------------- losses (costs.py)
def dice_coef_perbatch(y_true, y_pred):
    """costs.py:43-48."""
    inter = (y_true * y_pred).sum(dim=(1, 2, 3))
    union = y_true.sum(dim=(1, 2, 3)) + y_pred.sum(dim=(1, 2, 3))
    return 1 - (2 * inter + 1e-12) / (union + 1e-12)


def dice_loss(y_true, y_pred, restrict_chn):
    """costs.py:50-67 make_dice_loss_fnc(restrict_chn)."""
    return dice_coef_perbatch(y_true[..., :restrict_chn], y_pred[..., :restrict_chn]).mean(0)


def weighted_cross_entropy_loss(y_pred, y_true):
    """costs.py:70-85, with the parameter names of the reference signature."""
    num_classes = y_true.shape[-1]
    n = [y_true[..., c].sum() for c in range(num_classes)]
    n_tot = sum(n)
    weights = torch.stack([n_tot / (n[c] + 1e-12) for c in range(num_classes)])
    yp = y_pred.reshape(-1, num_classes)
    yt = y_true.reshape(-1, num_classes)
    w_ce = yt * torch.log(yp + 1e-12) * weights
    return (-w_ce.sum(1)).mean()


def combined_dice_bce(y_true, y_pred, num_classes, lambda_bce=0.01):
    """costs.py:129-136: NOTE the call passes (y_true, y_pred) into a (y_pred, y_true) signature."""
    return dice_loss(y_true, y_pred, num_classes) + lambda_bce * weighted_cross_entropy_loss(y_true, y_pred)


def weighted_cross_entropy_perbatch(y_pred, y_true):
    """costs.py:88-108, with the parameter names of the reference signature -> [B] (mean over pixels per sample)."""
    B, H, W, C = y_true.shape
    n = y_true.sum(dim=(0, 1, 2))
    n_tot = n.sum()
    weights = n_tot / (n + 1e-12)
    yp = y_pred.reshape(-1, H * W, C)
    yt = y_true.reshape(-1, H * W, C)
    sm = torch.softmax(yp, dim=-1)
    w_ce = -(yt * torch.log(sm + 1e-12) * weights).sum(2)
    return w_ce.mean(1)


def combined_dice_bce_perbatch(y_true, y_pred, num_classes, lambda_bce=0.01):
    """costs.py:138-143 make_combined_dice_bce_perbatch: per-sample Dice over the first num_classes channels + the
    per-batch cross entropy, again called (y_true, y_pred) into a (y_pred, y_true) signature -> [B]."""
    return (dice_coef_perbatch(y_true[..., :num_classes], y_pred[..., :num_classes])
            + lambda_bce * weighted_cross_entropy_perbatch(y_true, y_pred))


def mae_single_input(y1, y2):
    """costs.py:24-26 -> [B, C]."""
    return (y1 - y2).abs().mean(dim=(1, 2))


def mae(y_true, y_pred):
    return (y_pred - y_true).abs().mean()


def mse(y_true, y_pred):
    return ((y_pred - y_true) ** 2).mean()


def kl(mean, log_var):
    """costs.py:186-189 -> [B,1]."""
    return (-0.5 * (1 + log_var - mean ** 2 - torch.exp(log_var)).sum(-1)).reshape(-1, 1)


def sampling(z_mean, z_log_var, eps):
    """utils/sdnet_utils.py:9-21 with the normal sample injected."""
    return z_mean + torch.exp(0.5 * z_log_var) * eps


def np_dice(y_true, y_pred, binarise=False, smooth=1e-12):
    """costs.py:31-41 (numpy metric)."""
    y_pred = y_pred[..., 0:y_true.shape[-1]]
    if binarise:
        y_pred = np.round(y_pred)
    y_int = y_true * y_pred
    return np.mean((2 * np.sum(y_int, axis=(1, 2, 3)) + smooth)
                   / (np.sum(y_true, axis=(1, 2, 3)) + np.sum(y_pred, axis=(1, 2, 3)) + smooth))


def pair_dice(a, b):
    """model_components/balancer.py:33-38."""
    inter = (a * b).sum(dim=(1, 2, 3))
    union = a.sum(dim=(1, 2, 3)) + b.sum(dim=(1, 2, 3))
    return ((2 * inter + 1e-12) / (union + 1e-12))[:, None]


def spectral_reg(kernel, u0, alpha=10.0):
    """layers/spectralnorm.py:215-239: 3 power iterations restarted from u0 every call (A12);
    alpha*mean|stop_gradient(W/sigma) - W|."""
    x = kernel.reshape(-1, kernel.shape[-1])
    u = u0
    with torch.no_grad():
        xd = x.detach()
        for _ in range(3):
            wtu = xd.t() @ u
            v = wtu / torch.sqrt((wtu ** 2).sum())
            wv = xd @ v
            u = wv / torch.sqrt((wv ** 2).sum())
        sigma = (u.t() @ xd @ v).reshape(())
        target = xd / sigma
    return alpha * (target - x).abs().mean()


def adam_step(p, g, m, v, t, lr=1e-4, b1=0.9, b2=0.999, eps=1e-7):
    """keras 2.1.6 Adam (A9); t is the 1-based step count.  numpy in, numpy out.
    beta_1/beta_2 are float32 backend variables in Keras, so (1 - beta) is formed in float32
    (1 - float32(0.999) = 0.0010000467, not 0.001)."""
    lr_t = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    b1f, b2f = float(np.float32(b1)), float(np.float32(b2))
    omb1 = float(np.float32(1.0) - np.float32(b1))
    omb2 = float(np.float32(1.0) - np.float32(b2))
    m = b1f * m + omb1 * g
    v = b2f * v + omb2 * g * g
    p = p - lr_t * m / (np.sqrt(v) + eps)
    return p, m, v
