"""GPU-box helper: ring depth of the strip-queue kernels in the latency-bound regime (few long pairs, ragged)."""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops, plan as P
from gpu_sq_perf import timeit, zipf_lengths
d = torch.device("cuda:0")

def run(name, pl):
    g = torch.Generator(device=d).manual_seed(2)
    shape = (pl.packed_floats,) if pl.packed else (pl.B, pl.N, pl.M)
    theta = torch.rand(shape, generator=g, device=d)
    A = -torch.rand(shape, generator=g, device=d)
    Et = torch.ones(pl.B, device=d)
    Vt, Q = ops.sq_forward(pl, theta, A)
    for ring in (3, 4, 6, 8):
        tf = timeit(lambda: ops.sq_forward(pl, theta, A, flags=ring << 24), it=5, warm=2)
        print(json.dumps({"name": name, "pass": "fwd", "ring": ring, "ms": round(tf, 4)}), flush=True)
    for ring in (2, 3, 4, 6):
        tb = timeit(lambda: ops.sq_backward(pl, Et, Q, flags=ring << 24), it=5, warm=2)
        print(json.dumps({"name": name, "pass": "bwd", "ring": ring, "ms": round(tb, 4)}), flush=True)

xl, yl = zipf_lengths(1024, np.random.default_rng(0))
run("b32x1024", P.Plan(32, 1024, 1024, device=d))
run("b32x512", P.Plan(32, 512, 512, device=d))
run("c5p", P.Plan(1024, 1024, 1024, xl, yl, packed=True, device=d))
run("c2", P.Plan(1024, 256, 256, device=d))
run("c4", P.Plan(1024, 512, 512, device=d))
