"""GPU-box helper: ring depth x grid size sweep of the strip-queue forward / backward (the plan's list schedule
is rebuilt for each grid size)."""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops, plan as P
from gpu_sq_perf import timeit, zipf_lengths
d = torch.device("cuda:0")

PER_SM = (3, 4, 5, 6, 7, 8, 10, 12, 13)


def run(name, mk):
    best = {}
    for per_sm in PER_SM:
        W = 148 * per_sm
        pl = mk(W)
        g = torch.Generator(device=d).manual_seed(2)
        shape = (pl.packed_floats,) if pl.packed else (pl.B, pl.N, pl.M)
        theta = torch.rand(shape, generator=g, device=d)
        A = -torch.rand(shape, generator=g, device=d)
        Et = torch.ones(pl.B, device=d)
        Vt, Q = ops.sq_forward(pl, theta, A)
        for ring in (4,):
            fl = (ring << 24) | (W << 8)
            tf = timeit(lambda: ops.sq_forward(pl, theta, A, flags=fl), it=5, warm=2)
            print(json.dumps({"name": name, "pass": "fwd", "per_sm": per_sm, "ring": ring, "ms": round(tf, 4)}), flush=True)
        for ring in (2, 3):
            if ring == 3 and per_sm > 10:
                continue
            fl = (ring << 24) | (W << 8)
            tb = timeit(lambda: ops.sq_backward(pl, Et, Q, flags=fl), it=5, warm=2)
            print(json.dumps({"name": name, "pass": "bwd", "per_sm": per_sm, "ring": ring, "ms": round(tb, 4)}), flush=True)

xl, yl = zipf_lengths(1024, np.random.default_rng(0))
run("c5p", lambda W: P.Plan(1024, 1024, 1024, xl, yl, packed=True, device=d, resident_warps=W))
run("b32", lambda W: P.Plan(32, 1024, 1024, device=d, resident_warps=W))
tl_x, tl_y = zipf_lengths(32, np.random.default_rng(1))
run("train32", lambda W: P.Plan(32, 1024, 1024, tl_x, tl_y, packed=True, device=d, resident_warps=W))
run("b64x512", lambda W: P.Plan(64, 512, 512, device=d, resident_warps=W))
run("b256x256", lambda W: P.Plan(256, 256, 256, device=d, resident_warps=W))
run("c2", lambda W: P.Plan(1024, 256, 256, device=d, resident_warps=W))
run("c4", lambda W: P.Plan(1024, 512, 512, device=d, resident_warps=W))
