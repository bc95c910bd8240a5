"""diagnostic: which stage bounds the raster-strip forward kernel?  DAFK_NC_DEBUG bits: 1 = one MMA pair per tile instead of all,
2 = epilogue without global stores, 4 = no staging loads (register-staged kernel only), 8 = accumulators dropped."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_b200 import ops
from scripts.bench_nc import CASES, timeit
from multimodal_segmentation_b200._lib import ACT_LRELU

for name in sys.argv[1:] or ["film8x8 B192 f32"]:
    N, Hh, W, Cin, Cout, k, pad, xdt, ydt = CASES[name]
    x = torch.randn(N, Hh, W, Cin, device="cuda").to(xdt)
    w = torch.randn(k, k, Cin, Cout, device="cuda") * 0.1
    b = torch.zeros(Cout, device="cuda")
    wp = ops.pack_conv_nc(w, 0)
    for raw in ("1", "0"):
        os.environ["DAFK_NC_RAW"] = raw
        row = []
        for dbg in (0, 1, 2, 3, 4, 8, 9, 12, 13):
            os.environ["DAFK_NC_DEBUG"] = str(dbg)
            row.append("dbg%d %.1f" % (dbg, timeit(lambda: ops.conv_nc_fwd(x, wp, b, Cout, k, k, pad, ACT_LRELU, 0.3, ydt), 10)))
        print(name, "RAW=" + raw, " | ".join(row), flush=True)
os.environ.pop("DAFK_NC_DEBUG")
