"""Sweep of the raster-strip geometry (rows per strip R, ring depth S) on the narrow-channel layers of one DAFNet train_batch:
times forward / data gradient / weight gradient for each forced R against the geometry search's own choice.
usage: python scripts/bench_nc_sweep.py [case-substring] [reps]     (diagnostic; knobs in csrc/conv_nc.cu nc_tune)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from multimodal_segmentation_b200 import ops  # noqa: E402
from multimodal_segmentation_b200._lib import ACT_LRELU  # noqa: E402
from scripts.bench_nc import CASES, timeit  # noqa: E402

RS = [int(v) for v in os.environ.get("NC_SWEEP_R", "0,1,2,3,4,6,8,12,16,24,32").split(",")]


def main():
    sel = sys.argv[1] if len(sys.argv) > 1 else ""
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    print("# us per launch for forced rows-per-strip R (0 = the search's choice); '-' = geometry does not fit")
    print("%-22s %-6s %s" % ("case", "pass", " ".join("%7s" % ("R=%d" % r) for r in RS)))
    for name, (N, Hh, W, Cin, Cout, k, pad, xdt, ydt) in CASES.items():
        if sel and sel not in name:
            continue
        os.environ.pop("DAFK_NC_R", None)
        os.environ.pop("DAFK_NC_WG_R", None)
        x = torch.randn(N, Hh, W, Cin, device="cuda").to(xdt)
        w = torch.randn(k, k, Cin, Cout, device="cuda") * 0.1
        b = torch.zeros(Cout, device="cuda")
        Ho, Wo = Hh + 2 * pad - k + 1, W + 2 * pad - k + 1
        dy = torch.randn(N, Ho, Wo, Cout, device="cuda")
        dw, db = ops.zeros(k, k, Cin, Cout), ops.zeros(Cout)
        wp = ops.pack_conv_nc(w, 0) if ops.nc_supported(Cin, Cout, k, k, W, pad, 0) else None
        wpd = ops.pack_conv_nc(w, 1) if ops.nc_supported(Cin, Cout, k, k, W, pad, 1) else None
        wg = ops.nc_supported(Cin, Cout, k, k, W, pad, 2)
        rows = {"fwd": [], "dgrad": [], "wgrad": []}
        for R in RS:
            for var in ("DAFK_NC_R", "DAFK_NC_WG_R"):
                if R:
                    os.environ[var] = str(R)
                else:
                    os.environ.pop(var, None)
            for label, fn, ok in (
                    ("fwd", lambda: ops.conv_nc_fwd(x, wp, b, Cout, k, k, pad, ACT_LRELU, 0.3, ydt), wp is not None),
                    ("dgrad", lambda: ops.conv_nc_fwd(dy, wpd, None, Cin, k, k, k - 1 - pad), wpd is not None),
                    ("wgrad", lambda: ops.conv_nc_wgrad(x, dy, dw, db, pad), wg)):
                t = None
                if ok:
                    try:
                        t = timeit(fn, reps)
                    except Exception:
                        t = None
                rows[label].append("%7s" % ("-" if t is None else "%.1f" % t))
        for label in ("fwd", "dgrad", "wgrad"):
            print("%-22s %-6s %s" % (name, label, " ".join(rows[label])), flush=True)
        del x, dy
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
