"""micro-benchmark: the tcgen05 raster-strip kernels (csrc/conv_nc.cu) on the narrow-channel layers of one DAFNet train_batch (B = 32 pairs, 224^2; the FiLM decoder runs 6 decodes as one batch of 192).
Bytes = algorithmic (input + output once, in their storage dtypes); peak = MEASURED_PEAKS.json hbm_gbs.
usage: python scripts/bench_nc.py [case-substring] [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from multimodal_segmentation_b200 import ops  # noqa: E402
from multimodal_segmentation_b200._lib import ACT_LRELU, ACT_NONE  # noqa: E402

F, H = torch.float32, torch.bfloat16
CASES = {
    # name: (N, H, W, Cin, Cout, k, pad, x dtype, y dtype)
    "film8x8 B192 f32": (192, 224, 224, 8, 8, 3, 1, F, F),
    "film8x8 B192 bf16": (192, 224, 224, 8, 8, 3, 1, H, H),
    "seg8x64 f32->bf16": (32, 224, 224, 8, 64, 3, 1, F, H),
    "unet1x64 f32->bf16": (32, 224, 224, 1, 64, 3, 1, F, H),
    "loc16x20 5x5": (32, 224, 224, 16, 20, 5, 0, F, F),
    "loc20x20 5x5 @110": (32, 110, 110, 20, 20, 5, 0, F, F),
    "d0 s2d 4x64 k2": (32, 112, 112, 4, 64, 2, 0, H, F),
    "d0 s2d 16x64 k2": (32, 112, 112, 16, 64, 2, 0, H, F),
    "encm s2d 36x16 k2": (32, 112, 112, 36, 16, 2, 0, H, F),
    "encm s2d 64x32 k2": (32, 56, 56, 64, 32, 2, 0, H, F),
    "out8x1 1x1 B192": (192, 224, 224, 8, 1, 1, 0, F, F),
}


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def main():
    sel = sys.argv[1] if len(sys.argv) > 1 else ""
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6500.0
    print("# us per launch | GB/s algorithmic | fraction of the measured copy bandwidth (%.0f GB/s)" % peak)
    for name, (N, Hh, W, Cin, Cout, k, pad, xdt, ydt) in CASES.items():
        if sel and sel not in name:
            continue
        x = torch.randn(N, Hh, W, Cin, device="cuda").to(xdt)
        w = torch.randn(k, k, Cin, Cout, device="cuda") * 0.1
        b = torch.zeros(Cout, device="cuda")
        Ho, Wo = Hh + 2 * pad - k + 1, W + 2 * pad - k + 1
        dy = torch.randn(N, Ho, Wo, Cout, device="cuda")           # gradients arrive in fp32
        yact = torch.randn(N, Ho, Wo, Cout, device="cuda").to(ydt)
        dw, db = ops.zeros(k, k, Cin, Cout), ops.zeros(Cout)
        nb_f = x.numel() * x.element_size() + dy.numel() * yact.element_size()
        nb_d = dy.numel() * 4 + x.numel() * 4
        nb_w = x.numel() * x.element_size() + dy.numel() * 4
        out = []

        def rec(label, t, nb):
            out.append("%s %7.1f us %5.0f GB/s %3.0f%%" % (label, t, nb / t / 1e3, 100.0 * nb / t / 1e3 / peak))
        if ops.nc_supported(Cin, Cout, k, k, W, pad, 0):
            wp = ops.pack_conv_nc(w, 0)
            rec("nc fwd", timeit(lambda: ops.conv_nc_fwd(x, wp, b, Cout, k, k, pad, ACT_LRELU, 0.3, ydt), reps), nb_f)
        if ops.nc_supported(Cin, Cout, k, k, W, pad, 1):
            wpd = ops.pack_conv_nc(w, 1)
            rec("nc dgrad", timeit(lambda: ops.conv_nc_fwd(dy, wpd, None, Cin, k, k, k - 1 - pad), reps), nb_d)
        if ops.nc_supported(Cin, Cout, k, k, W, pad, 2):
            rec("nc wgrad", timeit(lambda: ops.conv_nc_wgrad(x, dy, dw, db, pad), reps), nb_w)
        print("%-20s %s" % (name, "\n                     ".join(out)), flush=True)
        del x, dy, yact
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
