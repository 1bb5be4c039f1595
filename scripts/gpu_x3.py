"""GPU-box experiment: strip-queue forward (TMA operand staging) ring x grid sweep at C2 / C4, against fwd3."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops, plan as P
from gpu_x1 import timeit
d = torch.device("cuda:0")
for B, N, M in ((1024, 256, 256), (1024, 512, 512)):
    g = torch.Generator(device=d).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=d)
    A = -torch.rand(B, N, M, generator=g, device=d)
    f3 = timeit(lambda: ops.forward_pass(theta, A, "nw"))
    print(json.dumps({"B": B, "N": N, "M": M, "fwd3_ms": round(f3, 4)}), flush=True)
    for per_sm in (10, 13, 16, 18, 20, 24):
        W = 148 * per_sm
        pl = P.Plan(B, N, M, device=d, resident_warps=W)
        for ring in (3, 4, 6):
            fl = (ring << 24) | (W << 8)
            try:
                tf = timeit(lambda: ops.sq_forward(pl, theta, A, flags=fl))
                ts = timeit(lambda: ops.sq_forward(pl, theta, A, need_q=False, flags=fl))
                print(json.dumps({"B": B, "N": N, "M": M, "per_sm": per_sm, "ring": ring, "fwd_ms": round(tf, 4), "score_ms": round(ts, 4)}), flush=True)
            except Exception as e:
                print("ERR", per_sm, ring, repr(e)[:100], flush=True)
