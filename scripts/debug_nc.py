"""debug helper: narrow-channel tcgen05 conv vs torch fp64, both descriptor conventions (DAFK_NC_SWAP)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [(2, 16, 16, 8, 8, 3, 1), (2, 24, 40, 8, 64, 3, 1), (2, 40, 36, 16, 20, 5, 0), (1, 224, 224, 8, 8, 3, 1)]


def child():
    import numpy as np
    import torch
    from multimodal_segmentation_b200 import ops
    from tests.util import cpu, gpu, rel_l2, t
    from oracle import ref_ops as R

    def bf(a):
        return torch.as_tensor(a).to(torch.bfloat16).float().numpy()
    for case in CASES:
        N, H, W, Cin, Cout, k, pad = case
        r = np.random.RandomState(1)
        x = bf(r.normal(size=(N, H, W, Cin)).astype(np.float32))
        w = bf((r.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(np.float32))
        wt = t(w, torch.float64, grad=True)
        xt = t(x, torch.float64, grad=True)
        yr = R.conv2d(xt, wt, None, 1, "same" if pad else "valid")
        dy = bf(r.normal(size=tuple(yr.shape)).astype(np.float32))
        (yr * t(dy, torch.float64)).sum().backward()
        try:
            y = ops.conv_nc_fwd(gpu(x), ops.pack_conv_nc(gpu(w), 0), None, Cout, k, k, pad)
            print(case, "fwd err", rel_l2(cpu(y), yr.detach().numpy()), flush=True)
            dx = ops.conv_nc_fwd(gpu(dy), ops.pack_conv_nc(gpu(w), 1), None, Cin, k, k, k - 1 - pad)
            print(case, "dgrad err", rel_l2(cpu(dx), xt.grad.numpy()), flush=True)
            dw = ops.zeros(k, k, Cin, Cout)
            db = ops.zeros(Cout)
            ops.conv_nc_wgrad(gpu(x), gpu(dy), dw, db, pad)
            print(case, "wgrad err", rel_l2(cpu(dw), wt.grad.numpy()), "db err", rel_l2(cpu(db), dy.sum((0, 1, 2))), flush=True)
            if case == CASES[0]:
                a, b = cpu(dw), wt.grad.numpy()
                print("dw[tap 0] gpu\n", a[0, 0, :3, :4], "\nref\n", b[0, 0, :3, :4], flush=True)
        except Exception as e:
            print(case, "EXC", repr(e)[:300], flush=True)
            break


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child()
    else:
        for swap in ("0", "1"):
            print("=== DAFK_NC_SWAP=%s" % swap, flush=True)
            env = dict(os.environ, DAFK_NC_SWAP=swap)
            try:
                out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, timeout=240)
                print(out.stdout[-4000:], out.stderr[-1500:], flush=True)
            except subprocess.TimeoutExpired:
                print("TIMEOUT", flush=True)
