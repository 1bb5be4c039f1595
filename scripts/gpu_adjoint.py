#!/usr/bin/env python
"""Kernel-only timing of the adjoint pair (double backward: nw.py:178-199, 251-267) next to
forward / backward at the bench shapes.  usage: python scripts/gpu_adjoint.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepblast_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for B, N, M in ((1024, 256, 256), (1024, 512, 512), (64, 256, 256)):
    g = torch.Generator(device=dev).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=dev)
    A = -torch.rand(B, N, M, generator=g, device=dev)
    Et = torch.ones(B, device=dev)
    Zt = torch.randn(B, N + 2, M + 2, generator=g, device=dev)
    ZA = torch.zeros(B, N, M, device=dev)
    cells = B * N * M
    Vt, Q = ops.forward_pass(theta, A, "nw")
    E = ops.backward_pass(Et, Q, "nw", N=N)
    Vtd, Qd = ops.adjoint_forward_pass(Q, Zt, ZA)
    f = timeit(lambda: ops.forward_pass(theta, A, "nw"))
    b = timeit(lambda: ops.backward_pass(Et, Q, "nw", N=N))
    af = timeit(lambda: ops.adjoint_forward_pass(Q, Zt, ZA))
    ab = timeit(lambda: ops.adjoint_backward_pass(E, Q, Qd))
    fast = ops.adjoint_pair_fast(Q, E, Zt, None)
    fp = timeit(lambda: ops.adjoint_pair_fast(Q, E, Zt, None)) if fast is not None else float("nan")
    print("B=%d %dx%d  fwd %.3f  bwd %.3f  adj_fwd %.3f (%4.0f GB/s of 32 B/cell)  adj_bwd %.3f (%4.0f GB/s)  ms | "
          "chained adjoint pair (incl. the two interior copies) %.3f ms | training step (4 sweeps) %.1f -> %.1f Gcell/s" % (
              B, N, M, f, b, af, cells * 32 / af / 1e6, ab, cells * 32 / ab / 1e6, fp,
              cells / (f + b + af + ab) / 1e6, cells / (f + b + (fp if fp == fp else af + ab)) / 1e6), flush=True)
    if fast is not None:
        from deepblast_b200 import _lib
        L = _lib.lib()
        st = torch.cuda.current_stream().cuda_stream
        zt = Zt[:, 1:-1, 1:-1].contiguous()
        e = E[:, 1:-1, 1:-1].contiguous()
        QdE = ops.q_empty(B, N, M, dev)
        Vtd2 = torch.empty(B, device=dev)
        Ed = torch.empty(B, N + 2, M + 2, device=dev)
        tc = timeit(lambda: (Zt[:, 1:-1, 1:-1].contiguous(), E[:, 1:-1, 1:-1].contiguous()))
        tf = timeit(lambda: L.b200dp_adj_fwd3(Q.data_ptr(), zt.data_ptr(), None, e.data_ptr(), Vtd2.data_ptr(),
                                              QdE.data_ptr(), B, N, M, 0, st))
        tb = timeit(lambda: L.b200dp_adj_bwd3(Q.data_ptr(), QdE.data_ptr(), Ed.data_ptr(), None, B, N, M, 0, st))
        print("    pieces: interior copies %.3f  adj_fwd3 %.3f  adj_bwd3 %.3f ms" % (tc, tf, tb), flush=True)
    del fast
    del theta, A, Q, Qd, E, Zt, ZA
