#!/usr/bin/env python
"""bench.py -- DAFNet `train_batch` throughput on B200 (metric of BASELINE.json).

    python bench.py --gpus N --steps K --warmup W
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one full ``DAFNetExecutor.train_batch`` (model_executors/dafnet_executor.py:369-387 of the
reference: generator update + 2 mask-discriminator updates + 2 image-discriminator updates) on a
batch of B=32 paired synthetic CHAOS-shaped 224x224 T1/T2 slices per GPU.  One slice = one pair.

  value : slices/s with the inputs already resident in HBM (device-timed, max over ranks)
  e2e   : the same metric through the public executor API (pinned host batches -> H2D every step ->
          train_batch -> D2H of the loss slots every step)
  roofline : the dominant kernel family (tcgen05 implicit-GEMM convolution), algorithmic FLOPs per
          launch / CUDA-event duration of that launch, summed over the timed region
  cpu_baseline : the CPU oracle (torch-CPU restatement of the reference graph; TF 1.4 cannot be
          installed here) on a bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dafnet_film", choices=["dafnet_film", "dafnet_spade", "mmsdnet", "inference"],
                    help="dafnet_film = BASELINE config 2 (the metric), dafnet_spade = config 3, mmsdnet = config 1 (MMSDNet "
                         "train_batch, B = 4; the reference's CPU-runnable case), inference = config 5 (predict_mask "
                         "'simple': anatomy encoder + segmentor, use --size 512 --batch 128); config 4 = --l_mix 0.5")
    ap.add_argument("--batch", type=int, default=None, help="pairs per GPU (default 32; 4 for --workload mmsdnet)")
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--l_mix", type=float, default=1.0)
    ap.add_argument("--cpu-batch", type=int, default=8, help="pairs per CPU train_batch sample (8 pairs ~ 15 s on 16 host threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-tc", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying one CUDA graph per step")
    ap.add_argument("--profile-all", action="store_true", help="CUDA-event time of every entry point (diagnostic)")
    a = ap.parse_args()
    if a.batch is None:
        a.batch = 4 if a.workload == "mmsdnet" else 32
    return a


# ------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi, DURING the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# algorithmic flops (reference layer list; SURVEY.md 8d / BASELINE.md 3)
# ------------------------------------------------------------------------------------------------
GF_PER_PAIR = {"dafnet_film": 1241.9, "dafnet_spade": 2192.0, "mmsdnet": 1124.9}


def conf_for(args):
    from multimodal_segmentation_b200.configuration import dafnet_config_chaos, mmsdnet_config_chaos
    from multimodal_segmentation_b200.keras_like import EasyDict
    if args.workload == "mmsdnet":
        conf = EasyDict(mmsdnet_config_chaos.get((args.size, args.size, 1)))
    else:
        conf = EasyDict(dafnet_config_chaos.get((args.size, args.size, 1),
                                                decoder_type="spade" if args.workload == "dafnet_spade" else "film"))
    conf.batch_size = args.batch
    conf.l_mix = args.l_mix
    conf.n_pairs = 1
    conf.folder = "/tmp/dafk_bench_run"
    return conf


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle graph on the host cores
# ------------------------------------------------------------------------------------------------
def host_threads():
    """all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: undo that for the CPU arm)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_train_batch_seconds(conf, B, seed=0, workload="dafnet_film"):
    """one train_batch on the CPU oracle, on B pairs: DAFNet = generator fwd+bwd+Adam, D_Mask x2, D_Image x2 with their
    inference passes; MMSDNet = supervised generator update, Z-regressor update, one D_Mask update
    (model_executors/mmsdnet_executor.py:238-331).  Returns (seconds, threads)."""
    import torch
    from oracle import ref_step
    torch.set_num_threads(host_threads())
    torch.manual_seed(seed)
    t0 = time.perf_counter()
    if workload == "mmsdnet":
        ref_step.mmsdnet_train_batch_cpu(conf, B, seed)
    else:
        ref_step.dafnet_train_batch_cpu(conf, B, seed)
    return time.perf_counter() - t0, torch.get_num_threads()


def metric_name(args):
    return ("MMSDNet" if args.workload == "mmsdnet" else "DAFNet") + " train slices/s @%d^2" % args.size


def run_reference(args):
    """--impl reference: the reference's own CPU path cannot be run (TF 1.4 / Keras 2.1.6 are not
    installable: py3.12, no wheels, no network), so this arm times the CPU oracle = the restatement of the
    reference graph, with ALL host threads of the box (torchrun's OMP_NUM_THREADS=1 is overridden), on a bounded sample
    (cpu-batch pairs per step).  Under torchrun rank 0 alone runs it; the host is the same at every N, so the value does
    not depend on --gpus."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(host_threads())
    os.environ["MKL_NUM_THREADS"] = str(host_threads())
    import torch
    torch.set_num_threads(host_threads())
    conf = conf_for(args)
    B = min(args.cpu_batch, args.batch) if args.workload == "mmsdnet" else args.cpu_batch
    times = []
    # bounded sample: one warm-up call (builds the weights, warms the allocator) whatever --warmup says, and at
    # most --steps timed calls within a 4-minute budget (at least one)
    warm = min(args.warmup, 1)
    budget = time.perf_counter() + 240.0
    for i in range(warm + args.steps):
        dt, threads = cpu_train_batch_seconds(conf, B, seed=i, workload=args.workload)
        if i >= warm:
            times.append(dt)
            if time.perf_counter() + dt > budget:
                break
    ms = 1000.0 * float(np.mean(times))
    val = B / (ms / 1000.0)
    line = {
        "impl": "reference", "metric": metric_name(args), "value": val, "unit": "slices/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": warm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s train_batch l_mix=%g %dx%d, CPU sample of %d pairs per step" %
                   (args.workload, args.l_mix, args.size, args.size, B)},
        "cpu_baseline": {"value": val, "unit": "slices/s", "cores": threads, "kind": "port",
                         "sample": "%d pairs per train_batch (the GPU arm uses %d per GPU); CPU restatement of the "
                                   "reference graph (torch-CPU fp32, %d threads = every host thread of this box; the host "
                                   "does not grow with --gpus, so this value is the same at every N); TF 1.4 is not "
                                   "installable" % (B, args.batch, threads)},
        "e2e": {"value": val, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def nccl_env():
    """stdout carries exactly one JSON line; NCCL's INFO log (the driver counts the ranks of the communicator in it) must
    still be visible.  NCCL writes it to the C-level stdout, and NCCL_DEBUG_FILE=/dev/stderr truncates the log when stderr
    is a regular file, so file descriptor 1 is pointed at stderr for native code and Python's sys.stdout keeps the real
    stdout for the JSON line."""
    # forced, not setdefault: the boxes export NCCL_DEBUG=VERSION, which prints the version line only
    os.environ["NCCL_DEBUG"] = os.environ.get("DAFK_NCCL_DEBUG", "INFO")
    os.environ["NCCL_DEBUG_SUBSYS"] = os.environ.get("DAFK_NCCL_DEBUG_SUBSYS", "INIT")
    os.environ.pop("NCCL_DEBUG_FILE", None)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not getattr(nccl_env, "done", False):
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
        nccl_env.done = True


def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    nccl_env()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from multimodal_segmentation_b200 import _lib, engine as E
    from multimodal_segmentation_b200 import parallel
    from multimodal_segmentation_b200 import instrument
    if args.workload == "mmsdnet":
        from multimodal_segmentation_b200.models.mmsdnet import MMSDNet as Net
        from multimodal_segmentation_b200.model_executors.mmsdnet_executor import MMSDNetExecutor as Executor
    else:
        from multimodal_segmentation_b200.models.dafnet import DAFNet as Net
        from multimodal_segmentation_b200.model_executors.dafnet_executor import DAFNetExecutor as Executor

    E.USE_TC = not args.no_tc
    _lib.PROFILE_ALL = args.profile_all
    conf = conf_for(args)
    conf.seed = 10 + rank                     # per-rank data / sampling seed; weights are broadcast from rank 0
    os.environ["DAFK_TRAIN_PAIRS"] = str(max(4 * args.batch, 64))
    np.random.seed(conf.seed)
    net = Net(conf)
    net.build()
    if world > 1:
        parallel.enable_data_parallel(net)
    ex = Executor(conf, net)
    ex.init_train_data()

    # ---- resident-input mode: pre-stage a pool of step inputs in HBM
    pool = [ex.stage_step_inputs() for _ in range(2)]
    torch.cuda.synchronize()

    def step_resident(i):
        ex.train_batch_on(pool[i % len(pool)])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- per-kernel CUDA-event timing + launch count: an eager (host-launched) pass over the same step, because
    #      kernels inside a replayed CUDA graph cannot be bracketed by events individually
    for i in range(min(args.warmup, 2)):
        step_resident(i)
    barrier()
    probe_steps = 1 if not args.profile_all else args.steps
    instrument.reset()
    instrument.enabled = True
    launches0 = _lib.launch_count()
    for i in range(probe_steps):
        step_resident(i)
    barrier()
    instrument.enabled = False
    launches_per_step = (_lib.launch_count() - launches0) / probe_steps
    kern = instrument.summary()
    kern_steps = probe_steps
    _lib.PROFILE_ALL = False
    ex._pending = []

    use_graph = not args.no_graph and not args.profile_all
    if use_graph:
        ex.enable_cuda_graph()
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        step_resident(i)
    ev1.record()
    host_ms = (time.perf_counter() - t_host0) * 1000.0 / args.steps      # host time per step WITH back-pressure (below)
    barrier()
    # host cost of enqueueing ONE step into an idle queue (copies into the graph's static buffers + one graph launch).
    # The per-step average above is larger because the launch queue is finite: once a few replays are queued the
    # host blocks in cudaGraphLaunch until the device drains them, i.e. it converges to the DEVICE time per step.
    t_idle0 = time.perf_counter()
    step_resident(args.steps)
    host_idle_ms = (time.perf_counter() - t_idle0) * 1000.0
    barrier()
    launches = launches_per_step * args.steps
    ms_total = ev0.elapsed_time(ev1)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    pairs_per_step = args.batch * (2 if 0 < args.l_mix < 1 else 1)
    value = world * pairs_per_step * args.steps / (ms_total / 1000.0)

    # ---- end to end through the executor API (pinned host -> H2D -> train_batch -> D2H losses)
    e2e = None
    if not args.no_e2e:
        ex.h2d_bytes = ex.d2h_bytes = 0
        losses = {n: [] for n in ex.get_loss_names()}
        for i in range(max(1, args.warmup // 2)):
            ex.train_batch(losses)
            ex.flush_losses(losses)
        barrier()
        ex.h2d_bytes = ex.d2h_bytes = 0
        t0 = time.perf_counter()
        for i in range(args.steps):
            ex.train_batch(losses)
            ex.flush_losses(losses)       # D2H read of the step's losses (synchronises)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * pairs_per_step * args.steps / dt, "unit": "slices/s",
               "h2d_bytes_per_step": int(ex.h2d_bytes / args.steps), "d2h_bytes_per_step": int(ex.d2h_bytes / args.steps),
               "last_loss": float(np.mean(losses["loss"][-1:])) if losses.get("loss") else None}

    def finish():
        """multi-rank teardown: a captured CUDA graph that contains NCCL collectives makes
        destroy_process_group() hang, so every rank leaves through a barrier and a hard exit"""
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        finish()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_burst = peaks.get("bf16_tflops", 1650.0)
    peak_src = "measured (sustained, kernel timed inside a long step)" if peaks else "fallback 1.4 PFLOP/s sustained"
    roof = None
    if kern:
        dom = max(kern.values(), key=lambda k: k["ms"])
        ach = dom["flops"] / (dom["ms"] / 1000.0) / 1e12 if dom["ms"] > 0 else 0.0
        # DRAM traffic cannot be read without a profiler: it comes from the committed ncu launch list of THIS workload
        # (scripts/ncu_traffic.py: dram__bytes_read.sum + dram__bytes_write.sum over EVERY launch of the family in
        # host-launched train_batch calls, mean per launch) -- the same population of launches as
        # algorithmic_bytes_per_launch_step_mean, so the two are comparable; other workloads report null
        traffic, traffic_detail = None, {"note": "no ncu capture of this workload under profiles/"}
        if args.workload == "dafnet_film" and args.batch == 32 and args.size == 224 and args.l_mix == 1.0 and E.USE_TC:
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "conv_tc_traffic.json")))
                traffic = tj["dram_bytes_per_launch"]
                traffic_detail = {k: tj[k] for k in tj if k != "dram_bytes_per_launch"}
            except Exception:
                pass
        roof = {"bound": "tensor", "kernel": dom["name"], "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": ach / peak_tf, "frac_of_burst_peak": ach / peak_burst, "peak_burst": peak_burst,
                "traffic": traffic, "traffic_detail": traffic_detail,
                "algorithmic_bytes_per_launch_step_mean": dom["bytes"] / max(dom["n"], 1),
                "peak_source": peak_src + "; the family runs inside an 80 ms step, so the sustained figure applies; "
                                          "frac_of_burst_peak is given for comparison",
                "launches": dom["n"], "share_of_step": (dom["ms"] / kern_steps) / ms_step,
                "timing": "CUDA events around every launch of the family in a host-launched pass over the same step "
                          "(%d step); the headline value replays the step as one CUDA graph" % kern_steps,
                "all_kernels": {k: {"ms_per_step": v["ms"] / kern_steps, "launches_per_step": v["n"] / kern_steps,
                                    "TFLOP/s": (v["flops"] / (v["ms"] / 1000.0) / 1e12) if v["ms"] > 0 and v["flops"] else None,
                                    "GB/s": (v["bytes"] / (v["ms"] / 1000.0) / 1e9) if v["ms"] > 0 and v["bytes"] else None}
                                for k, v in kern.items()}}
    cpu = None
    if not args.no_cpu_baseline and world == 1:      # rank 0 at N=1 only
        cb = min(args.cpu_batch, args.batch)
        dtc, threads = cpu_train_batch_seconds(conf, cb, workload=args.workload)
        cpu = {"value": cb / dtc, "unit": "slices/s", "cores": threads, "kind": "port",
               "sample": "one train_batch on %d pairs (%.1f s); CPU restatement of the reference graph (torch-CPU fp32), "
                         "TF 1.4 / Keras 2.1.6 are not installable here" % (cb, dtc)}
    algo_tf = GF_PER_PAIR[args.workload] * pairs_per_step / 1000.0
    line = {
        "metric": metric_name(args), "value": value, "unit": "slices/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if E.USE_TC else "f32", "data": "synthetic",
        "config": {"workload": "%s train_batch (%s), l_mix=%g, %dx%d, %d pairs per GPU"
                               % (args.workload, "generator + Z-regressor + D_Mask updates" if args.workload == "mmsdnet"
                                  else "generator + 2x D_Mask + D_Image1 + D_Image2 updates", args.l_mix, args.size,
                                  args.size, args.batch),
                   "parallelism": "dp%d" % world,
                   "l2": "inputs+activations per step are tens of GB, far larger than the 126 MB L2",
                   "launch": "one CUDA graph per step" if use_graph else "host-launched kernels",
                   "algorithmic_tflop_per_step_per_gpu": algo_tf},
        "step_tflops_per_gpu": algo_tf / (ms_step / 1000.0), "host_enqueue_ms_per_step": host_ms,
        "host_enqueue_ms_one_step_idle_queue": host_idle_ms,
        "host_enqueue_note": "host_enqueue_ms_per_step is averaged over the timed loop and includes back-pressure: the "
                             "launch queue is finite, so after a few queued replays cudaGraphLaunch blocks until the "
                             "device catches up (the average tends to the device time per step); the idle-queue figure "
                             "is the real host cost of one step (static-buffer copies + one graph launch)",
        "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    finish()


def run_inference(args):
    """BASELINE config 5: predict_mask(type='simple') = Segmentor(Enc_Anatomy(x)) in inference phase (models/mmsdnet.py:
    210-224), device-timed with the images resident in HBM; replicas only at N > 1 (no collective)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    nccl_env()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from multimodal_segmentation_b200 import _lib, engine as E
    from multimodal_segmentation_b200.models.dafnet import DAFNet
    E.USE_TC = not args.no_tc
    conf = conf_for(args)
    np.random.seed(10 + rank)
    net = DAFNet(conf)
    net.build()
    B, S = args.batch, args.size
    x = [torch.rand(B, S, S, 1, device="cuda") * 2 - 1 for _ in range(2)]
    for _ in range(args.warmup):
        net.predict_mask_device(1, "simple", x[0], x[1])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = net.predict_mask_device(1, "simple", x[0], x[1])
    ev1.record()
    torch.cuda.synchronize()
    launches = int(_lib.launch_count() - l0)
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # ---- end to end: pinned host images -> H2D -> predict -> D2H of the soft masks (what Segmentor.predict returns)
    e2e = None
    if not args.no_e2e:
        hx = torch.empty((B, S, S, 1), dtype=torch.float32).pin_memory()
        hx.uniform_(-1, 1)
        hy = torch.empty(tuple(out.shape), dtype=out.dtype).pin_memory()
        nst = max(1, min(args.steps, 5))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(nst):
            xd = hx.cuda(non_blocking=True)
            o = net.predict_mask_device(1, "simple", x[0], xd)
            hy.copy_(o, non_blocking=True)
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * nst / float(tt.item()), "unit": "slices/s", "h2d_bytes_per_step": int(hx.numel() * 4),
               "d2h_bytes_per_step": int(hy.numel() * hy.element_size()), "steps": nst}
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    gf = 104.2 * (S / 224.0) ** 2      # algorithmic GFLOP per slice (104.2 @224^2, 544.3 @512^2)
    if rank == 0:
        print(json.dumps({
            "metric": "segmentor-only inference slices/s @%d^2" % S, "value": world * B * args.steps / (ms / 1000.0),
            "unit": "slices/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if E.USE_TC else "f32",
            "data": "synthetic", "config": {"workload": "predict_mask('simple'): anatomy encoder + segmentor, %dx%d, %d per GPU"
                                                      % (S, S, B), "parallelism": "replicas x%d" % world,
                                            "l2": "every 64-channel map is %.1f GB, far larger than the 126 MB L2"
                                                  % (B * S * S * 64 * 2 / 1e9)},
            "step_tflops_per_gpu": gf * B / 1000.0 / (ms / args.steps / 1000.0), "clocks": sampler.summary(), "e2e": e2e,
            "gpu_launches": launches, "out_shape": list(out.shape)}))
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "inference":
        run_inference(a)
    else:
        run_b200(a)
