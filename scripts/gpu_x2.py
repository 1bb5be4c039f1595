"""GPU-box experiment: where the strip-queue kernels' time goes (diagnostic flags; results are wrong when set)."""
import os, sys, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops, plan as P
from gpu_sq_perf import zipf_lengths
d = torch.device("cuda:0")


def timeit(fn, it=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def run(name, pl):
    g = torch.Generator(device=d).manual_seed(2)
    shape = (pl.packed_floats,) if pl.packed else (pl.B, pl.N, pl.M)
    theta = torch.rand(shape, generator=g, device=d)
    A = -torch.rand(shape, generator=g, device=d)
    Et = torch.ones(pl.B, device=d)
    Vt, Q = ops.sq_forward(pl, theta, A)
    for dbg in (0, 1, 2, 4, 6):
        fl = dbg << 28
        tf = timeit(lambda: ops.sq_forward(pl, theta, A, flags=fl))
        print(json.dumps({"name": name, "pass": "fwd", "dbg": dbg, "ms": round(tf, 4)}), flush=True)
    tf = timeit(lambda: ops.sq_forward(pl, theta, A, need_q=False))
    print(json.dumps({"name": name, "pass": "score", "ms": round(tf, 4)}), flush=True)
    for dbg in (0, 4, 8, 12):
        fl = (dbg << 28) & 0xFFFFFFFF
        if fl >= 1 << 31:
            fl -= 1 << 32
        tb = timeit(lambda: ops.sq_backward(pl, Et, Q, flags=fl))
        print(json.dumps({"name": name, "pass": "bwd", "dbg": dbg, "ms": round(tb, 4)}), flush=True)


run("b32", P.Plan(32, 1024, 1024, device=d))
run("c2", P.Plan(1024, 256, 256, device=d))
xl, yl = zipf_lengths(1024, np.random.default_rng(0))
run("c5p", P.Plan(1024, 1024, 1024, xl, yl, packed=True, device=d))
run("b1", P.Plan(1, 1024, 1024, device=d))
