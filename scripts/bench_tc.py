"""micro-benchmark of the tcgen05 3x3 convolution at the UNet layer shapes (B=32 @ 224^2 input).
usage: python scripts/bench_tc.py [substring] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from multimodal_segmentation_b200 import _lib, ops  # noqa: E402

B = int(os.environ.get("TC_B", "32"))
LAYERS = [  # name, H, C0, C1, Cout
    ("d0c2_64x64@224", 224, 64, 0, 64), ("d1c1_64x128@112", 112, 64, 0, 128), ("d1c2_128x128@112", 112, 128, 0, 128),
    ("d2c1_128x256@56", 56, 128, 0, 256), ("d2c2_256x256@56", 56, 256, 0, 256), ("d3c1_256x512@28", 28, 256, 0, 512),
    ("d3c2_512x512@28", 28, 512, 0, 512), ("bt1_512x1024@14", 14, 512, 0, 1024), ("bt2_1024x1024@14", 14, 1024, 0, 1024),
    ("u3up_1024x512@28", 28, 1024, 0, 512), ("u3cat_512+512x512@28", 28, 512, 512, 512),
    ("u2up_512x256@56", 56, 512, 0, 256), ("u2cat_256+256x256@56", 56, 256, 256, 256),
    ("u1up_256x128@112", 112, 256, 0, 128), ("u1cat_128+128x128@112", 112, 128, 128, 128),
    ("u0up_128x64@224", 224, 128, 0, 64), ("u0cat_64+64x64@224", 224, 64, 64, 64),
]


def timeit(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    sel = sys.argv[1] if len(sys.argv) > 1 else ""
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    L = _lib.lib().fn
    S = _lib.stream_ptr()
    tot = {"fwd": [0.0, 0.0], "wgrad": [0.0, 0.0]}
    for name, H, C0, C1, Cout in LAYERS:
        if sel and sel not in name:
            continue
        Cin = C0 + C1
        x0 = torch.randn(B, H, H, C0, device="cuda").to(torch.bfloat16)
        x1 = torch.randn(B, H, H, C1, device="cuda").to(torch.bfloat16) if C1 else None
        w = torch.randn(3, 3, Cin, Cout, device="cuda") * 0.05
        wp = ops.pack_conv(w, 0)
        bias = torch.zeros(Cout, device="cuda")
        ybf = os.environ.get("TC_YBF16", "1") == "1"       # output stored in bf16 as in the benched step (engine.RAW_BF16)
        y = torch.empty(B, H, H, Cout, device="cuda", dtype=torch.bfloat16 if ybf else torch.float32)
        dy = torch.randn(B, H, H, Cout, device="cuda").to(torch.bfloat16)
        dw = ops.zeros(3, 3, Cin, Cout)
        fl = 2.0 * B * H * H * Cout * 9 * Cin
        f = L["dafk_conv_tc_fwd"]
        args = [x0.data_ptr(), C0, x1.data_ptr() if C1 else None, C1, wp.data_ptr(), Cout, 0, bias.data_ptr(), y.data_ptr(), 1 if ybf else 0,
                B, H, H, Cout, 3, 3, 1, 1, H, H, H * H * Cout, H * Cout, Cout, S]
        t = timeit(lambda: f(*args), reps)
        g = L["dafk_conv_tc_wgrad"]
        wargs = [x0.data_ptr(), C0, 0, Cin, dy.data_ptr(), Cout, dw.data_ptr(), B, H, H, 3, 3, 1, 1, H, H, S]
        flw = 2.0 * B * H * H * Cout * 9 * C0
        tw = timeit(lambda: g(*wargs), reps)
        th = None
        if L["dafk_conv3x3_tc_wgrad_halo_supported"](C0, Cout):
            gh = L["dafk_conv3x3_tc_wgrad_halo"]
            hargs = [x0.data_ptr(), C0, 0, Cin, dy.data_ptr(), Cout, dw.data_ptr(), B, H, H, S]
            th = timeit(lambda: gh(*hargs), reps)
        tot["fwd"][0] += t; tot["fwd"][1] += fl
        tot["wgrad"][0] += tw; tot["wgrad"][1] += flw
        print("%-24s fwd %8.1f us %7.1f TF/s | wgrad(src0) %8.1f us %7.1f TF/s | wgrad halo %s" %
              (name, t, fl / t / 1e6, tw, flw / tw / 1e6, "-" if th is None else "%8.1f us %7.1f TF/s" % (th, flw / th / 1e6)), flush=True)
    for k, (t, fl) in tot.items():
        if t:
            print("total %s: %.1f us, %.1f TF/s" % (k, t, fl / t / 1e6))


if __name__ == "__main__":
    main()
