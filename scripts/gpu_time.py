#!/usr/bin/env python
"""Kernel-only timing sweep (CUDA events around back-to-back launches) + host overhead
per call.  usage: python scripts/gpu_time.py [B N M]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepblast_b200 import ops  # noqa: E402

B, N, M = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (1024, 256, 256)))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(2)
theta = torch.rand(B, N, M, generator=g, device=dev)
A = -torch.rand(B, N, M, generator=g, device=dev)
Et = torch.ones(B, device=dev)
cells = B * N * M


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / iters
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, host * 1e3


print(f"B={B} N={N} M={M}")
configs = [(0, 0), (1, 0), (2, 0), (1, 148 * 4), (4, 148 * 4)]
for W, G in configs:
    fl = (W << 4) | (G << 8)
    try:
        Vt, Q = ops.forward_pass(theta, A, "nw", flags=fl)
        f_ms, f_host = timeit(lambda: ops.forward_pass(theta, A, "nw", flags=fl))
        b_ms, b_host = timeit(lambda: ops.backward_pass(Et, Q, "nw", flags=fl, N=N))
        print("W=%d grid=%-5d fwd %.3f ms (%.0f GB/s, host %.3f ms)  bwd %.3f ms (%.0f GB/s, host %.3f ms)  fwd+bwd %.1f Gcell/s"
              % (W, G, f_ms, cells * 20 / f_ms / 1e6, f_host, b_ms, cells * 16 / b_ms / 1e6, b_host,
                 cells / (f_ms + b_ms) / 1e6), flush=True)
    except Exception as e:  # noqa: BLE001
        print("W=%d grid=%d failed: %s" % (W, G, e), flush=True)

# ---- host-side cost of the autograd path (no sync inside the timed parts) ----------
from deepblast_b200.nw_cuda import NeedlemanWunschFunction as Fn  # noqa: E402
th = theta.clone().requires_grad_()
for _ in range(3):
    v = Fn.apply(th, A, 'softmax'); g, = torch.autograd.grad(v.sum(), th)
torch.cuda.synchronize()
tf = ts = tg = 0.0
n_it = 20
for _ in range(n_it):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); v = Fn.apply(th, A, 'softmax'); t1 = time.perf_counter()
    s = v.sum(); t2 = time.perf_counter()
    g, = torch.autograd.grad(s, th); t3 = time.perf_counter()
    tf += t1 - t0; ts += t2 - t1; tg += t3 - t2
print("host ms per call: Function.apply %.3f  sum %.3f  autograd.grad %.3f" % (tf / n_it * 1e3, ts / n_it * 1e3, tg / n_it * 1e3))

# ---- do the kernels slow down when they alternate (as in a training step)? -------------
def alt(iters=10):
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(iters)]
    for it in range(iters):
        ev[it][0].record()
        Vt, Q = ops.forward_pass(theta, A, "nw")
        ev[it][1].record()
        E = ops.backward_pass(Et, Q, "nw", N=N)
        ev[it][2].record()
    torch.cuda.synchronize()
    f = sum(e[0].elapsed_time(e[1]) for e in ev[2:]) / (iters - 2)
    b = sum(e[1].elapsed_time(e[2]) for e in ev[2:]) / (iters - 2)
    return f, b
alt(4)
f, b = alt(12)
print("alternating C-ABI calls: fwd %.3f ms  bwd %.3f ms  -> %.1f Gcell/s" % (f, b, cells / (f + b) / 1e6))
