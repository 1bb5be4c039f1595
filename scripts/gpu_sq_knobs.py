"""GPU-box helper: ring depth / grid / diagnostic sweeps of the strip-queue kernels at C2 and C5 packed."""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops, plan as P
from gpu_sq_perf import timeit, zipf_lengths

d = torch.device("cuda:0")


def run(name, pl, mode="nw"):
    g = torch.Generator(device=d).manual_seed(2)
    shape = (pl.packed_floats,) if pl.packed else (pl.B, pl.N, pl.M)
    theta = torch.rand(shape, generator=g, device=d)
    A = -torch.rand(shape, generator=g, device=d)
    Et = torch.ones(pl.B, device=d)
    Vt, Q = ops.sq_forward(pl, theta, A, mode)
    for ring in (3, 4, 6, 8):
        for dbg in (0, 1, 2, 4, 6):
            for ctas in (0,):
                fl = (ring << 24) | (dbg << 28) | (ctas << 8)
                tf = timeit(lambda: ops.sq_forward(pl, theta, A, mode, flags=fl), it=5, warm=2)
                print(json.dumps({"name": name, "pass": "fwd", "ring": ring, "dbg": dbg, "ms": round(tf, 4)}), flush=True)
    for ctas in (148 * 4, 148 * 6, 148 * 8, 148 * 10):
        fl = (4 << 24) | (ctas << 8)
        tf = timeit(lambda: ops.sq_forward(pl, theta, A, mode, flags=fl), it=5, warm=2)
        print(json.dumps({"name": name, "pass": "fwd", "ring": 4, "ctas": ctas, "ms": round(tf, 4)}), flush=True)
    Vt, Q = ops.sq_forward(pl, theta, A, mode)
    for ring in (2, 3, 4, 6):
        for dbg in (0, 4):
            fl = (ring << 24) | (dbg << 28)
            tb = timeit(lambda: ops.sq_backward(pl, Et, Q, mode, flags=fl), it=5, warm=2)
            print(json.dumps({"name": name, "pass": "bwd", "ring": ring, "dbg": dbg, "ms": round(tb, 4)}), flush=True)
    for ctas in (148 * 4, 148 * 6, 148 * 8):
        fl = (3 << 24) | (ctas << 8)
        tb = timeit(lambda: ops.sq_backward(pl, Et, Q, mode, flags=fl), it=5, warm=2)
        print(json.dumps({"name": name, "pass": "bwd", "ring": 3, "ctas": ctas, "ms": round(tb, 4)}), flush=True)


if __name__ == "__main__":
    run("c2", P.Plan(1024, 256, 256, device=d))
    xl, yl = zipf_lengths(1024, np.random.default_rng(0))
    run("c5p", P.Plan(1024, 1024, 1024, xl, yl, packed=True, device=d))
