#!/usr/bin/env python
"""Kernel-only timing sweep of the chained kernels over their env knobs (one process:
the library reads the environment at every call).  usage: python scripts/gpu_time_v3.py [fwd|bwd|both]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepblast_b200 import ops  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "both"
dev = torch.device("cuda:0")
KNOBS = ["B200DP_V3", "B200DP_NCH", "B200DP_RING", "B200DP_DBG", "B200DP_CTAS", "B200DP_BRING"]


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def setenv(kv):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in kv.items()})


def sweep(B, N, M, combos):
    g = torch.Generator(device=dev).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=dev)
    A = -torch.rand(B, N, M, generator=g, device=dev)
    Et = torch.ones(B, device=dev)
    cells = B * N * M
    print(f"== B={B} N={N} M={M}", flush=True)
    for kv in combos:
        setenv(kv)
        try:
            f_ms = b_ms = float("nan")
            if what in ("fwd", "both"):
                f_ms = timeit(lambda: ops.forward_pass(theta, A, "nw"))
            if what in ("bwd", "both") and not kv.get("B200DP_DBG"):
                setenv({})
                Vt, Q = ops.forward_pass(theta, A, "nw")
                setenv(kv)
                b_ms = timeit(lambda: ops.backward_pass(Et, Q, "nw", N=N))
            print("%-60s fwd %.3f ms (%4.0f GB/s)  bwd %.3f ms (%4.0f GB/s)  %.1f Gcell/s" % (
                " ".join("%s=%s" % (k[7:], v) for k, v in kv.items()) or "default",
                f_ms, cells * 20 / f_ms / 1e6, b_ms, cells * 16 / b_ms / 1e6, cells / (f_ms + b_ms) / 1e6), flush=True)
        except Exception as e:  # noqa: BLE001
            print("%s failed: %s" % (kv, e), flush=True)
    setenv({})


KNOBS += ["B200DP_PFW", "B200DP_PFD"]
base = [{"B200DP_V3": 0}, {}]
if what == "fwd":
    pf = [{"B200DP_NCH": n, "B200DP_RING": r, "B200DP_PFW": w, "B200DP_PFD": d}
          for n in (1, 2) for r in (3,) for w in (64, 128, 256) for d in (8, 16)]
    pf += [{"B200DP_NCH": 1, "B200DP_RING": 6, "B200DP_PFW": 128, "B200DP_PFD": 12}]
    sweep(1024, 256, 256, base + [{"B200DP_NCH": 1, "B200DP_RING": 3}] + pf)
    sweep(1024, 512, 512, base + [{"B200DP_NCH": 1, "B200DP_RING": 3}] +
          [{"B200DP_NCH": 1, "B200DP_RING": 3, "B200DP_PFW": w, "B200DP_PFD": 8} for w in (64, 128, 256)])
    sweep(4096, 256, 256, base + [{"B200DP_NCH": 1, "B200DP_RING": 3}] +
          [{"B200DP_NCH": n, "B200DP_RING": 3, "B200DP_PFW": w, "B200DP_PFD": 8} for n in (1, 2) for w in (64, 128)])
else:
    br = [{"B200DP_BRING": r} for r in (2, 3, 4, 6)] + [{"B200DP_BRING": r, "B200DP_CTAS": 512} for r in (4, 6)]
    sweep(1024, 256, 256, base + br)
    sweep(1024, 512, 512, base + br[:4])
    sweep(4096, 256, 256, base + br[:4])
