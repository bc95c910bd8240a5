"""Optional per-kernel-family CUDA-event timing used by bench.py (roofline numbers measured live inside
the timed region, on the launching stream).  Off by default: zero overhead on the product path."""
import torch

import os

enabled = False
by_shape = bool(os.environ.get("DAFK_PROFILE_SHAPES"))    # diagnostic: one record per (family, layer shape)
_records = {}
trace = None        # a list while an ordered trace of every C-ABI call is being taken (scripts/trace_step.py)


def reset():
    _records.clear()


def timed(name, flops, nbytes, fn, tag=None):
    if not enabled:
        return fn()
    if by_shape and tag is not None:
        name = "%s %s" % (name, tag)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    rec = _records.setdefault(name, {"name": name, "events": [], "flops": 0.0, "bytes": 0.0})
    rec["events"].append((e0, e1))
    rec["flops"] += float(flops)
    rec["bytes"] += float(nbytes)
    return r


def summary():
    torch.cuda.synchronize()
    out = {}
    for name, rec in _records.items():
        ms = sum(a.elapsed_time(b) for a, b in rec["events"])
        out[name] = {"name": name, "ms": ms, "n": len(rec["events"]), "flops": rec["flops"], "bytes": rec["bytes"]}
    return out
