"""GPU-box experiment: strip-queue backward (Q slot re-armed inside the block), resident warps per SM x ring at C2 / C4."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops, plan as P
from gpu_x16 import timeit
d = torch.device("cuda:0")
for B, N, M in ((1024, 256, 256), (1024, 512, 512)):
    g = torch.Generator(device=d).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=d)
    A = -torch.rand(B, N, M, generator=g, device=d)
    Et = torch.ones(B, device=d)
    for per_sm in (7, 8, 9, 10, 11, 12):
        W = 148 * per_sm
        pl = P.Plan(B, N, M, device=d, resident_warps=W)
        Vt, Q = ops.sq_forward(pl, theta, A)
        for ring in (2, 3):
            fl = (ring << 24) | (W << 8)
            try:
                tb = timeit(lambda: ops.sq_backward(pl, Et, Q, flags=fl))
                print(json.dumps({"B": B, "N": N, "M": M, "per_sm": per_sm, "ring": ring, "bwd_ms": round(tb, 4)}), flush=True)
            except Exception as e:
                print("ERR", per_sm, ring, repr(e)[:100], flush=True)
