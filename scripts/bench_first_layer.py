"""micro-benchmark: the backward pass of the 8 -> 64 first layers (segmentor / discriminator conv1) with the output gradient stored
in fp32 (raster-strip data gradient) or bf16 (swizzled tcgen05 kernel for the data gradient, raster-strip weight gradient)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_b200 import ops
from scripts.bench_nc import timeit

N, H, W, Cin, Cout, k = 32, 224, 224, 8, 64, 3
x = torch.randn(N, H, W, Cin, device="cuda")
w = torch.randn(k, k, Cin, Cout, device="cuda") * 0.1
dy = torch.randn(N, H, W, Cout, device="cuda")
dyh = dy.to(torch.bfloat16)
dw, db = ops.zeros(k, k, Cin, Cout), ops.zeros(Cout)
wpd_nc = ops.pack_conv_nc(w, 1)
wpd_tc = ops.pack_conv(w, 1)
print("nc dgrad, fp32 dy      %.1f us" % timeit(lambda: ops.conv_nc_fwd(dy, wpd_nc, None, Cin, k, k, 1), 10))
print("nc dgrad, bf16 dy      %.1f us" % timeit(lambda: ops.conv_nc_fwd(dyh, wpd_nc, None, Cin, k, k, 1), 10))
print("tc dgrad, bf16 dy      %.1f us" % timeit(lambda: ops.conv_tc_fwd(dyh, None, wpd_tc, None, Cin, k, k, 1, 1, torch.float32), 10))
a = ops.conv_nc_fwd(dyh, wpd_nc, None, Cin, k, k, 1)
b = ops.conv_tc_fwd(dyh, None, wpd_tc, None, Cin, k, k, 1, 1, torch.float32)
print("tc vs nc data gradient: rel L2 %.2e" % ((a - b).norm() / a.norm()).item())
print("nc wgrad, fp32 dy      %.1f us" % timeit(lambda: ops.conv_nc_wgrad(x, dy, dw, db, 1), 10))
print("nc wgrad, bf16 dy      %.1f us" % timeit(lambda: ops.conv_nc_wgrad(x, dyh, dw, db, 1), 10))
x1 = torch.randn(N, H, W, 1, device="cuda")
dw1 = ops.zeros(k, k, 1, Cout)
for raw in ("0", "1"):
    os.environ["DAFK_NC_RAW"] = raw
    print("RAW=%s 1 -> 64 wgrad fp32 dy %.1f us, bf16 dy %.1f us" % (raw, timeit(lambda: ops.conv_nc_wgrad(x1, dy, dw1, db, 1), 10),
                                                                  timeit(lambda: ops.conv_nc_wgrad(x1, dyh, dw1, db, 1), 10)))
    w1 = torch.randn(k, k, 1, Cout, device="cuda")
    wp1 = ops.pack_conv_nc(w1, 0)
    print("RAW=%s 1 -> 64 fwd %.1f us" % (raw, timeit(lambda: ops.conv_nc_fwd(x1, wp1, db, Cout, k, k, 1, 0, 0.0, torch.bfloat16), 10)))
