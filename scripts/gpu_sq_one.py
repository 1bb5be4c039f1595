"""GPU-box helper for ncu: a few launches of the strip-queue sweeps on one workload.
usage: gpu_sq_one.py c2|c3|c4|c5p|c5|b32 [flags] [iters]"""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops, plan as P
from gpu_sq_perf import zipf_lengths

d = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
flags = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mode = "sw" if which == "c3" else "nw"
xl, yl = zipf_lengths(1024, np.random.default_rng(0))
pl = {"c2": lambda: P.Plan(1024, 256, 256, device=d), "c3": lambda: P.Plan(1024, 256, 256, device=d),
      "c4": lambda: P.Plan(1024, 512, 512, device=d),
      "c5p": lambda: P.Plan(1024, 1024, 1024, xl, yl, packed=True, device=d),
      "c5": lambda: P.Plan(1024, 1024, 1024, xl, yl, device=d),
      "b32": lambda: P.Plan(32, 1024, 1024, device=d)}[which]()
g = torch.Generator(device=d).manual_seed(2)
shape = (pl.packed_floats,) if pl.packed else (pl.B, pl.N, pl.M)
theta = torch.rand(shape, generator=g, device=d)
A = -torch.rand(shape, generator=g, device=d)
Zt = torch.randn(shape, generator=g, device=d)
Et = torch.ones(pl.B, device=d)
for _ in range(iters):
    Vt, Q = ops.sq_forward(pl, theta, A, mode, flags=flags)
    E = ops.sq_backward(pl, Et, Q, mode, flags=flags)
    Vtd, QdE = ops.sq_adjoint_forward(pl, Q, Zt, None, E, flags=flags)
    Ed = ops.sq_adjoint_backward(pl, Q, QdE, flags=flags)
torch.cuda.synchronize()
print("ok", which, float(Vt.sum()))
