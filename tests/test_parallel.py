"""Batch-sharded data parallelism (multimodal_segmentation_b200/parallel.py).

* CPU, gloo, world_size 2: the host-side logic -- every rank ends up with rank 0's weights and BatchNorm state
  (one broadcast per flat arena), the gradient buckets are summed across ranks in place, per-rank loaders draw
  different shards, and the reference arm of bench.py prints on rank 0 only.
* GPU (two ranks sharing cuda:0 over gloo, so it runs on a one-GPU box; the production launch uses NCCL): one generator
  update on two different shards must leave BOTH ranks with the weights a single process gets from
  "gradients of shard A and of shard B, each with its own BatchNorm / class-weight statistics, averaged, one Adam
  step" -- the reference's semantics for a 2x batch, because Keras' fit() splits it into mini-batches of 32
  (model_executors/dafnet_executor.py:404 passes no batch_size) -- and with identical weights on both ranks.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _tiny_conf(seed):
    from multimodal_segmentation_b200.configuration import dafnet_config_chaos
    from multimodal_segmentation_b200.keras_like import EasyDict
    conf = EasyDict(dafnet_config_chaos.get((64, 64, 1)))
    conf.anatomy_encoder.filters = 16
    conf.n_pairs = 1
    conf.seed = seed
    conf.folder = "/tmp/dafk_test_no_such_folder"
    return conf


def _arenas(net):
    seen, out = set(), []
    models = list(net.Encoders_Anatomy) + [net.Enc_Modality, net.Anatomy_Fuser, net.Segmentor, net.Decoder, net.D_Mask,
                                           net.D_Image1, net.D_Image2]
    for m in models:
        for a in (m._scope.arena, m._scope.state):
            if id(a) not in seen and a.flat is not None:
                seen.add(id(a))
                out.append(a)
    return out


# ------------------------------------------------------------------------------------------------ CPU / gloo
def _cpu_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    _init(rank, world, port)
    from multimodal_segmentation_b200 import parallel
    from multimodal_segmentation_b200.models.dafnet import DAFNet
    from multimodal_segmentation_b200.models.trainers import Trainer
    from multimodal_segmentation_b200.loaders.synthetic_chaos import make_pairs
    np.random.seed(100 + rank)
    net = DAFNet(_tiny_conf(10 + rank))          # different initial weights on every rank
    net.build()
    before = [a.flat.clone() for a in _arenas(net)]
    d = parallel.enable_data_parallel(net)
    after = [a.flat.clone() for a in _arenas(net)]
    versions = [a.version for a in _arenas(net)]
    # gradient buckets: summed in place across ranks
    buckets = [torch.full((5,), float(rank + 1)), torch.arange(4, dtype=torch.float32) * (rank + 1)]
    d.allreduce_(buckets)
    # per-rank shards: seed = conf.seed + rank (bench.py, experiment.py)
    x1 = make_pairs(2, (64, 64, 1), 4, seed=10 + rank)[0]
    q.put((rank, [float(b.double().sum()) for b in before], [float(a.double().sum()) for a in after],
           [b.tolist() for b in buckets], d.world_size, d.rank, float(np.abs(x1).sum()), versions,
           Trainer.dist is d))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_broadcast_and_bucket_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cpu_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0, r1 = res
    assert r0[1] != r1[1]                       # the ranks really started from different weights
    assert r0[2] == r0[1]                       # rank 0 is the source: unchanged
    assert r1[2] == r0[1]                       # rank 1 now holds rank 0's parameters and BatchNorm state
    assert r0[3] == r1[3] == [[3.0] * 5, [0.0, 3.0, 6.0, 9.0]]     # (1 + 2) * base, identical on both ranks
    assert (r0[4], r0[5], r1[4], r1[5]) == (2, 0, 2, 1)
    assert r0[6] != r1[6]                       # different data shards
    assert all(v >= 1 for v in r1[7])           # packed bf16 weight copies are invalidated by the broadcast
    assert r0[8] and r1[8]                      # the trainers see the process group


def test_reference_arm_prints_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""


# ------------------------------------------------------------------------------------------------ GPU, 2 ranks
def _gpu_setup(seed):
    sys.path.insert(0, ROOT)
    from multimodal_segmentation_b200 import engine as E
    from multimodal_segmentation_b200.models.dafnet import DAFNet
    E.USE_TC = False                            # strict fp32 kernels: the comparison is then at round-off level
    np.random.seed(seed)
    net = DAFNet(_tiny_conf(seed))
    net.build()
    return net


def _shard(rank):
    from tests.test_models_gpu import make_batch
    return make_batch(_tiny_conf(0), 2, seed=21 + rank)


def _gen_weights(net):
    return torch.cat([p.data.reshape(-1) for p in net.generator_params()]).double().cpu().numpy()


def _gpu_worker(rank, world, port, q):
    torch.cuda.set_device(0)
    net = _gpu_setup(3 + rank)
    _init(rank, world, port)
    from multimodal_segmentation_b200 import parallel
    parallel.enable_data_parallel(net)
    tr = net.supervised_trainer
    dev = [torch.from_numpy(a).cuda() for a in _shard(rank)]
    tr.train_on_device(*dev)                    # forward/backward on the local shard, all-reduce, Adam(1/world)
    torch.cuda.synchronize()
    q.put((rank, _gen_weights(net), float(tr.book.buf.sum().item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_rank_step_equals_averaged_shard_gradients():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    w_rank0, w_rank1 = res[0][1], res[1][1]
    assert res[0][2] != res[1][2]               # different shards -> different local losses
    assert np.array_equal(w_rank0, w_rank1)     # same reduced gradient, same update: replicas stay bit-identical

    # single-process restatement: rank 0's initial weights, per-shard gradients averaged, one Adam step
    net = _gpu_setup(3)
    w_before = _gen_weights(net)
    tr = net.supervised_trainer
    grads = []
    moving = []
    for r in range(2):
        state0 = [a.flat.clone() for a in _arenas(net) if not a.with_grad]
        tr.forward_backward(*[torch.from_numpy(a).cuda() for a in _shard(r)])
        grads.append([b.clone() for b in tr.opt.grad_buckets()])
        if r == 0:
            # rank 1 never sees rank 0's BatchNorm moving-average update: restore the state between the shards
            moving = [a for a in _arenas(net) if not a.with_grad]
            for a, s0 in zip(moving, state0):
                a.flat.copy_(s0)
    for b, g0, g1 in zip(tr.opt.grad_buckets(), grads[0], grads[1]):
        b.copy_(g0 + g1)
    tr.opt.step(grad_scale=0.5)
    torch.cuda.synchronize()
    w_ref = _gen_weights(net)
    assert np.abs(w_ref - w_before).max() > 0
    # Adam's first step moves every weight by ~lr * sign(g): compare the UPDATES (atomics reorder fp32 sums, and a
    # gradient that is ~0 can flip the sign of its update, so bound the fraction of such entries)
    upd_ref, upd_dp = w_ref - w_before, w_rank0 - w_before
    close = np.abs(upd_ref - upd_dp) <= 1e-2 * np.abs(upd_ref).max()
    assert close.mean() > 0.999, close.mean()


def test_reference_arm_line_contract():
    """bench.py --impl reference on rank 0: one JSON line with the keys the driver reads (the CPU restatement of the
    reference graph on a bounded sample; a tiny geometry here so that the CPU suite stays fast)"""
    import json
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "0", "--size", "64", "--cpu-batch", "1"], env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().split("\n") if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "slices/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["metric"].startswith("DAFNet train slices/s") and "workload" in d["config"]


# ------------------------------------------------------------------------------------------------ overlapped all-reduce
def test_tape_last_writers_maps_arena_ranges_to_backward_nodes():
    """host logic of the overlapped all-reduce: a range of the gradient arena is final once the FIRST recorded node (in
    forward order) that owns a parameter inside it has run in the backward pass; ranges nobody writes report None"""
    sys.path.insert(0, ROOT)
    from multimodal_segmentation_b200 import engine as E
    arena, other = E.Arena(), E.Arena()
    ps = [arena.add("p%d" % i, (10,), np.zeros(10, np.float32)) for i in range(4)]      # offsets 0, 12, 24, 36
    q = other.add("q", (10,), np.zeros(10, np.float32))
    for p in ps + [q]:
        p.requires_grad = True
    ps[3].requires_grad = False                                                          # frozen: never recorded
    tape = E.Tape()
    ctx = E.Ctx(tape, training=True)
    x = E.Var(torch.zeros(1), requires_grad=True)
    for used in ([ps[0]], [ps[1], q], [ps[0], ps[2]], [ps[3]]):                          # p0 is used by nodes 0 and 2
        assert ctx.rec(x, *used)
        tape.record(lambda: None)
    assert [len(n) for n in tape.node_params] == [1, 2, 2, 0]
    ranges = [(arena, 0, 12), (arena, 12, 24), (arena, 24, 36), (arena, 36, 48), (arena, 0, 48), (other, 0, 12), (arena, 8, 14)]
    assert tape.last_writers(ranges) == [0, 1, 2, None, 0, 1, 0]
    order = []
    tape.backward(order.append)
    assert order == [3, 2, 1, 0] and tape.nodes == []


@pytest.mark.gpu
def test_overlapped_allreduce_fires_each_piece_after_its_last_write():
    """one generator forward/backward with a recording stand-in for the process group: every piece of the gradient arena
    handed to the asynchronous all-reduce during the backward pass must already hold its FINAL value (nothing writes into
    it afterwards), most of the arena must go out before the backward pass ends, and the pieces + the remainder that
    apply_gradients() reduces cover the bucket exactly once"""
    from multimodal_segmentation_b200.models.trainers import Trainer
    from tests.test_models_gpu import make_batch

    class Recorder(object):
        world_size, rank, overlap = 2, 0, True

        def __init__(self):
            self.fired, self.rest, self.waited = [], [], 0

        def allreduce_async(self, t):
            self.fired.append((t, t.clone()))
            return None

        def allreduce_(self, buckets):
            self.rest.extend(buckets)

        def wait(self, works):
            self.waited += len(works)

    net = _gpu_setup(5)
    from multimodal_segmentation_b200 import engine as E
    E.USE_TC = True
    tr = net.supervised_trainer
    old_chunk, old_dist = Trainer.AR_CHUNK, Trainer.dist
    Trainer.AR_CHUNK, Trainer.dist = 1 << 15, Recorder()
    try:
        dev = [torch.from_numpy(a).cuda() for a in make_batch(_tiny_conf(0), 2, seed=21)]
        tr.forward_backward(*dev)
        torch.cuda.synchronize()
        rec = Trainer.dist
        total = sum(b.numel() for b in tr.opt.grad_buckets())
        assert len(rec.fired) > 4 and sum(t.numel() for t, _ in rec.fired) > 0.9 * total
        for t, snap in rec.fired:
            assert torch.equal(t, snap)                       # final when it was handed to the collective
        assert any(float(snap.abs().sum()) > 0 for _, snap in rec.fired)
        tr.apply_gradients()
        torch.cuda.synchronize()
        assert rec.waited == len(rec.fired)
        assert sum(t.numel() for t, _ in rec.fired) + sum(b.numel() for b in rec.rest) == total
    finally:
        Trainer.AR_CHUNK, Trainer.dist = old_chunk, old_dist
        E.USE_TC = True
