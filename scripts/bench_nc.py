"""micro-benchmark of the narrow-channel tcgen05 convolutions at the DAFNet shapes (B=32, 224^2).
usage: python scripts/bench_nc.py [case-substring] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from multimodal_segmentation_b200 import ops  # noqa: E402

B = int(os.environ.get("NC_B", "32"))
CASES = {
    # name: (N, H, W, Cin, Cout, k, pad, x dtype)
    "film8x8": (B, 224, 224, 8, 8, 3, 1, torch.float32),
    "film8x8_bf16": (B, 224, 224, 8, 8, 3, 1, torch.bfloat16),
    "seg8x64": (B, 224, 224, 8, 64, 3, 1, torch.float32),
    "seg8x64_bf16": (B, 224, 224, 8, 64, 3, 1, torch.bfloat16),
    "unet1x64": (B, 224, 224, 1, 64, 3, 1, torch.float32),
    "loc16x20": (B, 224, 224, 16, 20, 5, 0, torch.float32),
    "loc20x20a": (B, 110, 110, 20, 20, 5, 0, torch.float32),
    "loc20x20b": (B, 53, 53, 20, 20, 5, 0, torch.float32),
    "head64x8": (B, 224, 224, 64, 8, 1, 0, torch.float32),
    "head64x5": (B, 224, 224, 64, 5, 1, 0, torch.float32),
    "out8x1": (B, 224, 224, 8, 1, 1, 0, torch.float32),
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
    for name, (N, H, W, Cin, Cout, k, pad, dt) in CASES.items():
        if sel and sel not in name:
            continue
        x = torch.randn(N, H, W, Cin, device="cuda").to(dt)
        w = torch.randn(k, k, Cin, Cout, device="cuda") * 0.1
        b = torch.zeros(Cout, device="cuda")
        Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
        dy = torch.randn(N, Ho, Wo, Cout, device="cuda").to(dt)
        wp, wpd = ops.pack_conv_nc(w, 0), ops.pack_conv_nc(w, 1)
        dw, db = ops.zeros(k, k, Cin, Cout), ops.zeros(Cout)
        es = x.element_size()
        res = []
        import ctypes
        from multimodal_segmentation_b200 import _lib
        L = _lib.lib().fn
        S = _lib.stream_ptr()
        dtc = 0 if dt == torch.float32 else 1

        def raw(name, *args):
            f = L["dafk_" + name]
            a = [x_.data_ptr() if hasattr(x_, "data_ptr") else x_ for x_ in args]
            return lambda: f(*a)
        y = torch.empty_like(dy)
        dx = torch.empty_like(x)
        t = timeit(raw("conv_nc_fwd", x, dtc, wp, b, y, dtc, N, H, W, Cin, Cout, k, k, pad, 2, 0.3, S), reps)
        nb = (x.numel() + dy.numel()) * es
        res.append("fwd %7.1f us %5.0f GB/s" % (t, nb / t / 1e3))
        if ops.nc_supported(Cin, Cout, k, k, W, pad, 1):
            t = timeit(raw("conv_nc_fwd", dy, dtc, wpd, None, dx, dtc, N, Ho, Wo, Cout, Cin, k, k, k - 1 - pad, 0, 0.0, S), reps)
            res.append("dgrad %7.1f us %5.0f GB/s" % (t, nb / t / 1e3))
        if ops.nc_supported(Cin, Cout, k, k, W, pad, 2):
            t = timeit(raw("conv_nc_wgrad", x, dtc, dy, dtc, dw, db, N, H, W, Cin, Cout, k, k, pad, S), reps)
            res.append("wgrad %7.1f us %5.0f GB/s" % (t, nb / t / 1e3))
        print("%-14s %s" % (name, " | ".join(res)), flush=True)


if __name__ == "__main__":
    main()
