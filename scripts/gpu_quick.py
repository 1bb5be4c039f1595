#!/usr/bin/env python
"""Kernel-only timing of the forward / backward passes at the bench shapes (CUDA events,
3 warm-ups, 10 iterations).  usage: python scripts/gpu_quick.py [KEY=VAL ...env knobs]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepblast_b200 import ops  # noqa: E402

for kv in sys.argv[1:]:
    k, v = kv.split("=")
    os.environ[k] = v
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


for mode, B, N, M in (("nw", 1024, 256, 256), ("sw", 1024, 256, 256), ("nw", 1024, 512, 512), ("nw", 4096, 256, 256),
                      ("nw", 2048, 128, 128), ("nw", 512, 1024, 1024)):
    g = torch.Generator(device=dev).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=dev)
    A = -torch.rand(B, N, M, generator=g, device=dev)
    Et = torch.ones(B, device=dev)
    cells = B * N * M
    f = timeit(lambda: ops.forward_pass(theta, A, mode))
    Vt, Q = ops.forward_pass(theta, A, mode)
    b = timeit(lambda: ops.backward_pass(Et, Q, mode, N=N))
    print("%s B=%d %dx%d  fwd %.3f ms %4.0f GB/s | bwd %.3f ms %4.0f GB/s | %.1f Gcell/s (%.2f of 182)" % (
        mode, B, N, M, f, cells * 20 / f / 1e6, b, cells * 16 / b / 1e6, cells / (f + b) / 1e6,
        cells / (f + b) / 1e6 / 182.0), flush=True)
    del theta, A, Q
