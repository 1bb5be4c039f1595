#!/usr/bin/env python
"""Where should the chained kernels take over from the hand-off kernels?  Forward + backward
(and the adjoint pair) at mid-size batches, default dispatch vs chained forced (B200DP_V3MIN=1).
usage: python scripts/gpu_threshold.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepblast_b200 import ops  # noqa: E402

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


for N, M in ((256, 256), (512, 512)):
    for B in (64, 96, 148, 200, 256, 296):
        g = torch.Generator(device=dev).manual_seed(2)
        theta = torch.rand(B, N, M, generator=g, device=dev)
        A = -torch.rand(B, N, M, generator=g, device=dev)
        Et = torch.ones(B, device=dev)
        Zt = torch.randn(B, N + 2, M + 2, generator=g, device=dev)
        row = []
        for v3min in ("100000", "1"):
            os.environ["B200DP_V3MIN"] = v3min
            Vt, Q = ops.forward_pass(theta, A, "nw")
            E = ops.backward_pass(Et, Q, "nw", N=N)
            f = timeit(lambda: ops.forward_pass(theta, A, "nw"))
            b = timeit(lambda: ops.backward_pass(Et, Q, "nw", N=N))
            if v3min == "1":
                ad = timeit(lambda: ops.adjoint_pair_fast(Q, E, Zt, None))
            else:
                def gen():
                    Vtd, Qd = ops.adjoint_forward_pass(Q, Zt, torch.zeros(B, N, M, device=dev))
                    ops.adjoint_backward_pass(E, Q, Qd)
                ad = timeit(gen, 3)
            row.append((f, b, ad))
        os.environ.pop("B200DP_V3MIN")
        (f0, b0, a0), (f1, b1, a1) = row
        print("%dx%d B=%3d  hand-off fwd %.3f bwd %.3f adj %.3f | chained fwd %.3f bwd %.3f adj %.3f ms" % (
            N, M, B, f0, b0, a0, f1, b1, a1), flush=True)
