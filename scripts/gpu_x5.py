"""GPU-box experiment: grid sizes that divide the ticket count (equal strips: the last round of a
non-dividing grid is partly idle)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops, plan as P
from gpu_x1 import timeit
d = torch.device("cuda:0")
for B, N, M, grids in ((1024, 256, 256, (1171, 1366, 1480, 1639, 1776, 1924, 2048, 2368)), (1024, 512, 512, (1366, 1490, 1639, 1821, 1924, 2048, 2341))):
    g = torch.Generator(device=d).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=d)
    A = -torch.rand(B, N, M, generator=g, device=d)
    Et = torch.ones(B, device=d)
    for W in grids:
        pl = P.Plan(B, N, M, device=d, resident_warps=W)
        Vt, Q = ops.sq_forward(pl, theta, A)
        for ring in (3, 4):
            fl = (ring << 24) | (W << 8)
            tf = timeit(lambda: ops.sq_forward(pl, theta, A, flags=fl))
            print(json.dumps({"B": B, "M": M, "W": W, "ring": ring, "fwd_ms": round(tf, 4)}), flush=True)
        for ring in (2, 3):
            fl = (ring << 24) | (W << 8)
            tb = timeit(lambda: ops.sq_backward(pl, Et, Q, flags=fl))
            print(json.dumps({"B": B, "M": M, "W": W, "ring": ring, "bwd_ms": round(tb, 4)}), flush=True)
