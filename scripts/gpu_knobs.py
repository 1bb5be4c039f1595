#!/usr/bin/env python
"""Kernel-only timing of fwd / bwd at one shape under several env-knob settings.
usage: python scripts/gpu_knobs.py B N M "K1=V1,K2=V2" "K1=V3" ..."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepblast_b200 import ops  # noqa: E402

B, N, M = (int(x) for x in sys.argv[1:4])
dev = torch.device("cuda:0")


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


g = torch.Generator(device=dev).manual_seed(2)
theta = torch.rand(B, N, M, generator=g, device=dev)
A = -torch.rand(B, N, M, generator=g, device=dev)
Et = torch.ones(B, device=dev)
cells = B * N * M
Vt, Q = ops.forward_pass(theta, A, "nw")
for combo in [""] + sys.argv[4:]:
    kv = dict(x.split("=") for x in combo.split(",") if x)
    for k in list(os.environ):
        if k.startswith("B200DP_"):
            del os.environ[k]
    os.environ.update({"B200DP_" + k: v for k, v in kv.items()})
    f = timeit(lambda: ops.forward_pass(theta, A, "nw"))
    b = timeit(lambda: ops.backward_pass(Et, Q, "nw", N=N))
    print("%-28s fwd %.3f ms | bwd %.3f ms | %.1f Gcell/s" % (combo or "default", f, b, cells / (f + b) / 1e6), flush=True)
