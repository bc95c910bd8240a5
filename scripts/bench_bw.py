#!/usr/bin/env python
"""Per-kernel HBM roofline of the bandwidth-bound kernels (SURVEY.md 8d byte counts), at the shapes of one DAFNet
train_batch (B=32 pairs @224^2).  Every kernel is launched `reps` times back to back between two CUDA events on the
launching stream; operands are far larger than the 126 MB L2 or rotated through a pool that is.

    python scripts/bench_bw.py [--reps 20] > profiles/rN_bench_bw.txt
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_segmentation_b200 import ops  # noqa: E402
from multimodal_segmentation_b200._lib import ACT_LRELU, ACT_RELU  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--only", default="")
args = ap.parse_args()
PEAK = 6544.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
dev = "cuda"
B, H, W = 32, 224, 224
rows = []


def timeit(name, nbytes, fn, pool=1):
    """fn(i) launches the kernel on operand set i % pool"""
    if args.only and args.only not in name:
        return
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    # `reps` launches captured into one CUDA graph (the product replays a whole step as a graph as well): the host
    # cost of a ctypes call (~10 us) would otherwise bound the small shapes
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(args.reps):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000.0 / args.reps
    gbs = nbytes / us / 1e3
    rows.append((name, nbytes / 1e6, us, gbs, gbs / PEAK))
    print("%-44s %9.1f MB %9.1f us %8.0f GB/s  %5.1f %% of %.0f" % (name, nbytes / 1e6, us, gbs, 100 * gbs / PEAK, PEAK), flush=True)


def rnd(*shape, dtype=torch.float32):
    return torch.randn(*shape, device=dev, dtype=torch.float32).to(dtype)


# ---- rounding / softmax (anatomy [B,224,224,8] f32)
x8 = [rnd(B, H, W, 8) for _ in range(3)]
n8 = x8[0].numel()
timeit("round_fwd [32,224,224,8] f32", 8.0 * n8, lambda i: ops.round_fwd(x8[i % 3]), 3)
timeit("softmax_fwd+round [..,8] f32", 12.0 * n8, lambda i: ops.softmax_fwd(x8[i % 3], True), 3)
timeit("softmax_bwd [..,8] f32", 12.0 * n8, lambda i: ops.softmax_bwd(x8[i % 3], x8[(i + 1) % 3]), 3)
timeit("max_fwd (anatomy fuser) [..,8] f32", 12.0 * n8, lambda i: ops.max_fwd(x8[i % 3], x8[(i + 1) % 3]), 3)
timeit("max_bwd [..,8] f32", 20.0 * n8, lambda i: ops.max_bwd(x8[i % 3], x8[(i + 1) % 3], x8[(i + 2) % 3]), 3)

# ---- FiLM (decoder: 6 call sites batched -> [192,224,224,8])
xf = [rnd(192, H, W, 8) for _ in range(2)]
gam, bet = rnd(192, 8), rnd(192, 8)
nf = xf[0].numel()
timeit("film_fwd [192,224,224,8] f32", 8.0 * nf, lambda i: ops.film_fwd(xf[i % 2], gam, bet), 2)
timeit("film_bwd [192,224,224,8] f32", 12.0 * nf, lambda i: ops.film_bwd(xf[i % 2], xf[(i + 1) % 2], gam), 2)
del xf

# ---- TPS warp (STN): vol [32,224,224,8]
theta = 0.05 * rnd(B, 25, 2)
timeit("tps_warp_fwd [32,224,224,8] f32", 8.0 * n8, lambda i: ops.tps_warp_fwd(x8[i % 3], theta), 3)
timeit("tps_warp_bwd [32,224,224,8] f32", 12.0 * n8, lambda i: ops.tps_warp_bwd(x8[i % 3], theta, x8[(i + 1) % 3]), 3)

# ---- BatchNorm at the UNet shapes
for (h, c) in ((224, 64), (112, 128), (56, 256), (28, 512), (14, 1024)):
    for dt, e in ((torch.bfloat16, 2), (torch.float32, 4)):
        if dt == torch.float32 and h != 224:
            continue
        xs = [rnd(B, h, h, c, dtype=dt) for _ in range(3)]
        n = xs[0].numel()
        mean, rstd, g, b = rnd(c), rnd(c).abs() + 0.5, rnd(c), rnd(c)
        tag = "[32,%d,%d,%d] %s" % (h, h, c, "bf16" if e == 2 else "f32")
        timeit("bn_stats " + tag, float(e) * n, lambda i: ops.bn_stats_finalize(xs[i % 3], 1e-3, 0.99), 3)
        timeit("bn_apply+relu -> bf16 " + tag, float(e + 2) * n,
               lambda i: ops.bn_apply(xs[i % 3], mean, rstd, g, b, ACT_RELU, torch.bfloat16), 3)
        dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
        dbias = torch.zeros(c, device=dev)
        gs = [rnd(B, h, h, c, dtype=torch.bfloat16) for _ in range(2)]
        timeit("bn_bwd (reduce+apply) " + tag, (2.0 * (e + 2) + e) * n,
               lambda i: ops.bn_bwd(gs[i % 2], xs[i % 3], mean, rstd, g, b, ACT_RELU, dg, db, dx_dtype=dt, dbias_prev=dbias), 3)
        del xs, gs

# ---- SPADE: per-sample instance norm + modulation, [32,224,224,128]? (decoder.py / spade.py: fin = 128 .. 8)
for c in (128, 8):
    xs = [rnd(B, H, W, c) for _ in range(2)]
    g2, b2 = rnd(B, H, W, c), rnd(B, H, W, c)
    n = xs[0].numel()
    acc = ops.in_stats(xs[0])
    timeit("in_stats [32,224,224,%d] f32" % c, 4.0 * n, lambda i: ops.in_stats(xs[i % 2]), 2)
    timeit("spade_fwd [32,224,224,%d] f32" % c, 16.0 * n, lambda i: ops.spade_fwd(xs[i % 2], acc, g2, b2), 2)
    timeit("spade_bwd [32,224,224,%d] f32" % c, 40.0 * n, lambda i: ops.spade_bwd(xs[(i + 1) % 2], xs[i % 2], acc, g2, b2), 2)
    del xs, g2, b2

# ---- pooling / upsampling / activations / casts / adds (bf16 feature maps)
xb = [rnd(B, H, W, 64, dtype=torch.bfloat16) for _ in range(3)]
nb = xb[0].numel()
timeit("maxpool2_fwd [32,224,224,64] bf16", 2.0 * nb * 1.25, lambda i: ops.maxpool2_fwd(xb[i % 3]), 3)
xh = [rnd(B, 112, 112, 128, dtype=torch.bfloat16) for _ in range(3)]
timeit("upsample2_fwd [32,112,112,128] bf16", 2.0 * xh[0].numel() * 5, lambda i: ops.upsample2_fwd(xh[i % 3]), 3)
timeit("add bf16 [32,224,224,64]", 6.0 * nb, lambda i: ops.add(xb[i % 3], xb[(i + 1) % 3]), 3)
timeit("cast bf16->f32 [32,224,224,64]", 6.0 * nb, lambda i: ops.cast(xb[i % 3], torch.float32), 3)
xf32 = [rnd(B, H, W, 64) for _ in range(2)]
timeit("cast f32->bf16 [32,224,224,64]", 6.0 * nb, lambda i: ops.cast(xf32[i % 2], torch.bfloat16), 2)
timeit("act_fwd lrelu f32 [32,224,224,64]", 8.0 * nb, lambda i: ops.act_fwd(xf32[i % 2], ACT_LRELU, 0.2), 2)
timeit("act_bwd lrelu f32 [32,224,224,64]", 12.0 * nb, lambda i: ops.act_bwd(xf32[i % 2], xf32[(i + 1) % 2], ACT_LRELU, 0.2), 2)
del xb, xh, xf32

# ---- pointwise heads on the 64-channel map (anatomy 64 -> 8, segmentor 64 -> 5)
xh64 = [rnd(B, H, W, 64, dtype=torch.bfloat16) for _ in range(3)]
for co in (8, 5):
    w = rnd(1, 1, 64, co)
    bias = rnd(co)
    dy = [rnd(B, H, W, co) for _ in range(2)]
    dw, db = torch.zeros(1, 1, 64, co, device=dev), torch.zeros(co, device=dev)
    npx = B * H * W
    timeit("conv1x1_fwd 64->%d bf16 -> f32" % co, npx * (128.0 + 4 * co), lambda i: ops.conv1x1_fwd(xh64[i % 3], w, bias), 3)
    timeit("conv1x1_dgrad %d->64 f32 -> bf16" % co, npx * (128.0 + 4 * co), lambda i: ops.conv1x1_dgrad(dy[i % 2], w), 2)
    timeit("conv1x1_wgrad 64x%d" % co, npx * (128.0 + 4 * co), lambda i: ops.conv1x1_wgrad(xh64[i % 3], dy[i % 2], dw, db), 3)
del xh64

# ---- Adam over the generator arena (45 M parameters, 28 B/param)
n = 45_000_000
p, g, m, v = rnd(n), rnd(n), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
state = torch.zeros(4, device=dev)
ops.adam_tick(state, 1e-4)
timeit("adam_step 45M params", 28.0 * n, lambda i: ops.adam_step_dev(p, g, m, v, None, state))
