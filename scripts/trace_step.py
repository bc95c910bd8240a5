"""Ordered trace of every C-ABI call of ONE host-launched train_batch at the benched configuration (diagnostic).
usage: python scripts/trace_step.py [out.txt] [workload] [batch]
Each line: microseconds, entry point, tensor operands (shape dtype).  Used to find glue passes to fold into their
producer / consumer kernels; the numbers are CUDA-event times on the launching stream."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from multimodal_segmentation_b200 import engine as E, instrument  # noqa: E402


class A:
    pass


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_step.txt"
    args = A()
    args.workload = sys.argv[2] if len(sys.argv) > 2 else "dafnet_film"
    args.batch = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    args.size, args.l_mix = 224, 1.0
    if args.workload == "mmsdnet":
        from multimodal_segmentation_b200.models.mmsdnet import MMSDNet as Net
        from multimodal_segmentation_b200.model_executors.mmsdnet_executor import MMSDNetExecutor as Executor
    else:
        from multimodal_segmentation_b200.models.dafnet import DAFNet as Net
        from multimodal_segmentation_b200.model_executors.dafnet_executor import DAFNetExecutor as Executor
    E.USE_TC = True
    conf = bench.conf_for(args)
    conf.seed = 10
    os.environ["DAFK_TRAIN_PAIRS"] = str(max(4 * args.batch, 64))
    np.random.seed(conf.seed)
    net = Net(conf)
    net.build()
    ex = Executor(conf, net)
    ex.init_train_data()
    pool = [ex.stage_step_inputs() for _ in range(2)]
    for i in range(2):
        ex.train_batch_on(pool[i % 2])
    torch.cuda.synchronize()
    instrument.trace = []
    ex.train_batch_on(pool[0])
    torch.cuda.synchronize()
    tr, instrument.trace = instrument.trace, None
    tot = 0.0
    with open(out, "w") as f:
        for name, ts, e0, e1 in tr:
            us = e0.elapsed_time(e1) * 1000.0
            tot += us
            f.write("%9.1f  %-26s %s\n" % (us, name, "  ".join("%s %s" % ("x".join(map(str, s)), d) for s, d in ts)))
        f.write("# %d calls, %.1f ms\n" % (len(tr), tot / 1000.0))
    print("%d calls, %.1f ms -> %s" % (len(tr), tot / 1000.0, out))


if __name__ == "__main__":
    main()
