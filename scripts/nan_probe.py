"""diagnostic: the 40 supervised steps of tests/test_config_shapes_gpu.py::_inference_net at 512^2, B = 2, lr 1e-3, repeated; prints
the loss curve and the first step at which a loss or a weight stops being finite."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multimodal_segmentation_b200 import engine as E
from tests.test_models_gpu import build_net, make_batch, product_step

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H = int(sys.argv[2]) if len(sys.argv) > 2 else 512
for rep in range(reps):
    net, conf = build_net(H=H, filters=64, rounding=True, use_tc=True, lr=1e-3)
    fixed = make_batch(conf, 2, seed=9)
    E.BatchNorm.MOMENTUM = 0.9
    curve, bad = [], None
    for step in range(40):
        tr = product_step(net, fixed, True)
        vals = tr.book.buf.cpu().numpy()
        curve.append(float(vals.sum()))
        g = torch.cat([p.grad.reshape(-1) for p in net.generator_params()])
        gfin = bool(torch.isfinite(g).all())
        if bad is None and (not np.all(np.isfinite(vals)) or not gfin):
            bad = (step, [round(float(v), 3) for v in vals], gfin)
            names = [p.name for p in net.generator_params() if not bool(torch.isfinite(p.grad).all())]
            print("  first non-finite at step %d: losses %s grads finite %s; non-finite gradients in %s" % (bad[0], bad[1], bad[2], names[:8]))
            break
        tr.apply_gradients()
    print("run %d: %s | %s" % (rep, "NON-FINITE" if bad else "ok", " ".join("%.1f" % c for c in curve[::3])), flush=True)
