"""Launch plans of the raster-strip convolutions (csrc/conv_nc.cu, dafk_conv_nc_plan: host code only, no GPU): the invariants
the kernels' barrier protocol and shared-memory carve-up rely on, over the shapes of the benched step and a random sweep.

The one that matters most: with bulk-copy staging the ring's slot count is a multiple of the six converter warps, so that
every slot has ONE owner warp.  Slot barriers are waited on by parity; with 8 or 16 slots a warp could reach its wait for
wrap k of a slot while another warp's wrap k-1 segment was still in flight, pass on the stale parity and convert a segment
that had not arrived (a rare trapped launch, found by scripts/stress_nc.py in round 2)."""
import ctypes

import numpy as np
import pytest

F32, BF16 = 0, 1
SMEM_MAX = 227 * 1024


def _plan(kind, N, H, W, Cin, Cout, k, pad, x_dt, o_dt):
    from multimodal_segmentation_b200 import _lib
    L = _lib.lib()
    out = (ctypes.c_int64 * 10)()
    rc = L.fn["dafk_conv_nc_plan"](kind, N, H, W, Cin, Cout, k, k, pad, x_dt, o_dt, ctypes.cast(out, ctypes.c_void_p))
    if rc != 0:
        return None
    keys = ("raw", "R", "S", "slots", "slot_bytes", "seg_px", "nseg", "seg_px_y", "nseg_y", "smem")
    return dict(zip(keys, [int(v) for v in out]))


def _check(p, W, Cin, esz, Wo=None, Cout=None, esz_y=None):
    assert 1 <= p["R"] <= 32 and p["smem"] <= SMEM_MAX
    if not p["raw"]:
        assert p["S"] >= 2 and p["slots"] == 0
        return
    assert p["S"] in (1, 2)
    assert p["slots"] in (6, 12), p                      # one owner warp per slot
    assert p["slot_bytes"] % 128 == 0
    for w, c, e, seg, nseg in ((W, Cin, esz, p["seg_px"], p["nseg"]),) + \
            (((Wo, Cout, esz_y, p["seg_px_y"], p["nseg_y"]),) if Wo is not None else ()):
        px = c * e
        assert 0 < seg <= w and (nseg - 1) * seg < w <= nseg * seg
        assert (seg * px) % 16 == 0 and ((w - (nseg - 1) * seg) * px) % 16 == 0      # bulk copies: 16-byte multiples
        assert seg * px <= p["slot_bytes"] <= 8192 + 127


STEP_SHAPES = [
    # N, H, W, Cin, Cout, k, pad, x dtype, y / dy dtype      (one DAFNet train_batch at B = 32 pairs, 224^2; scripts/bench_nc.py)
    (192, 224, 224, 8, 8, 3, 1, F32, F32), (192, 224, 224, 8, 8, 3, 1, BF16, BF16), (32, 224, 224, 8, 64, 3, 1, F32, BF16),
    (32, 224, 224, 1, 64, 3, 1, F32, BF16), (32, 224, 224, 64, 8, 3, 1, F32, F32), (32, 220, 220, 20, 16, 5, 4, F32, F32),
    (32, 110, 110, 20, 20, 5, 0, F32, F32), (32, 112, 112, 16, 64, 2, 0, BF16, F32), (32, 111, 111, 64, 16, 2, 1, F32, BF16),
    (32, 56, 56, 64, 32, 2, 0, BF16, F32), (192, 224, 224, 8, 1, 1, 0, F32, F32), (192, 224, 224, 1, 8, 1, 0, F32, F32),
    (128, 512, 512, 1, 64, 3, 1, F32, BF16), (128, 512, 512, 8, 64, 3, 1, F32, BF16), (2, 512, 512, 8, 8, 3, 1, F32, F32),
]


@pytest.mark.parametrize("force_raw", [None, "1", "0"])
def test_plans_of_the_benched_shapes(monkeypatch, force_raw):
    if force_raw is None:
        monkeypatch.delenv("DAFK_NC_RAW", raising=False)
    else:
        monkeypatch.setenv("DAFK_NC_RAW", force_raw)
    seen_raw = 0
    for (N, H, W, Cin, Cout, k, pad, xd, od) in STEP_SHAPES:
        esz = 4 if xd == F32 else 2
        p = _plan(0, N, H, W, Cin, Cout, k, pad, xd, od)
        assert p is not None, (N, H, W, Cin, Cout)
        _check(p, W, Cin, esz)
        seen_raw += p["raw"]
        if force_raw == "0":
            assert not p["raw"]
        Wo = W + 2 * pad - k + 1
        q = _plan(2, N, H, W, Cin, Cout, k, pad, xd, F32)
        if q is not None:
            _check(q, W, Cin, esz, Wo, Cout, 4)
            seen_raw += q["raw"]
    if force_raw != "0":
        assert seen_raw > 0
    if force_raw is None:
        # the default selects bulk-copy staging for the big FiLM-decoder maps and keeps register staging for small ones
        assert _plan(0, 192, 224, 224, 8, 8, 3, 1, F32, F32)["raw"] == 1
        assert _plan(2, 192, 224, 224, 8, 8, 3, 1, F32, F32)["raw"] == 1
        assert _plan(0, 32, 56, 56, 64, 32, 2, 0, BF16, F32)["raw"] == 0


def test_plans_of_a_random_sweep(monkeypatch):
    monkeypatch.setenv("DAFK_NC_RAW", "1")
    r = np.random.RandomState(7)
    n_raw = 0
    for _ in range(400):
        k = int(r.choice([1, 2, 3, 5]))
        pad = int(r.choice([0, k // 2, k - 1]))
        W = int(r.randint(max(k, 8), 600))
        H = int(r.randint(max(k, 8), 300))
        Cin, Cout = int(r.choice([1, 4, 8, 16, 20, 36, 64])), int(r.choice([1, 5, 8, 16, 20, 64]))
        xd = int(r.choice([F32, BF16]))
        N = int(r.randint(1, 200))
        esz = 4 if xd == F32 else 2
        p = _plan(0, N, H, W, Cin, Cout, k, pad, xd, F32)
        if p is not None:
            _check(p, W, Cin, esz)
            n_raw += p["raw"]
        q = _plan(2, N, H, W, Cin, Cout, k, pad, xd, F32)
        if q is not None:
            _check(q, W, Cin, esz, W + 2 * pad - k + 1, Cout, 4)
            n_raw += q["raw"]
    assert n_raw > 50
