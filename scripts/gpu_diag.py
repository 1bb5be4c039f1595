#!/usr/bin/env python
"""Diagnostic sweep: where does the chained kernels' time go?  DBG bit0 = no Q stores,
bit1 = no theta/A staging (fwd only).  usage: python scripts/gpu_diag.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepblast_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
KNOBS = ["B200DP_V3", "B200DP_NCH", "B200DP_RING", "B200DP_DBG", "B200DP_CTAS", "B200DP_BRING", "B200DP_PFW",
         "B200DP_PFD", "B200DP_WARPS"]


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


def sweep(B, N, M, combos, bwd_combos):
    g = torch.Generator(device=dev).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=dev)
    A = -torch.rand(B, N, M, generator=g, device=dev)
    Et = torch.ones(B, device=dev)
    cells = B * N * M
    print(f"== B={B} N={N} M={M}", flush=True)
    for kv in combos:
        setenv(kv)
        try:
            f_ms = timeit(lambda: ops.forward_pass(theta, A, "nw"))
            print("fwd %-50s %.3f ms (%4.0f GB/s, %.1f Gcell/s)" % (
                " ".join("%s=%s" % (k[7:], v) for k, v in kv.items()) or "default",
                f_ms, cells * 20 / f_ms / 1e6, cells / f_ms / 1e6), flush=True)
        except Exception as e:  # noqa: BLE001
            print("%s failed: %s" % (kv, e), flush=True)
    setenv({})
    Vt, Q = ops.forward_pass(theta, A, "nw")
    for kv in bwd_combos:
        setenv(kv)
        try:
            b_ms = timeit(lambda: ops.backward_pass(Et, Q, "nw", N=N))
            print("bwd %-50s %.3f ms (%4.0f GB/s, %.1f Gcell/s)" % (
                " ".join("%s=%s" % (k[7:], v) for k, v in kv.items()) or "default",
                b_ms, cells * 16 / b_ms / 1e6, cells / b_ms / 1e6), flush=True)
        except Exception as e:  # noqa: BLE001
            print("%s failed: %s" % (kv, e), flush=True)
    setenv({})


fw = [{}] + [{"B200DP_DBG": d} for d in (1, 2, 3)] + [{"B200DP_NCH": 2}] + \
     [{"B200DP_NCH": 2, "B200DP_DBG": d} for d in (1, 2, 3)] + [{"B200DP_RING": 4}, {"B200DP_RING": 6}] + \
     [{"B200DP_V3": 0, "B200DP_WARPS": w} for w in (1, 2, 4)]
bw = [{}] + [{"B200DP_BRING": r} for r in (2, 4, 6)] + [{"B200DP_V3": 0, "B200DP_WARPS": w} for w in (1, 2, 4)]
sweep(1024, 256, 256, fw, bw)
sweep(2048, 256, 256, fw[:8], bw[:2])
sweep(4096, 256, 256, fw[:8], bw[:2])
sweep(1024, 512, 512, fw[:5], bw[:2])
