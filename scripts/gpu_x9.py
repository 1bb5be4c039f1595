"""GPU-box experiment: alternative builds of the library (LIB=path) on the strip-queue timing set."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import _lib
if os.environ.get("LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["LIB"])
import numpy as np
import torch
from deepblast_b200 import ops, plan as P
from gpu_x1 import timeit
from gpu_sq_perf import zipf_lengths
d = torch.device("cuda:0")
ops.CLUSTER = False
xl, yl = zipf_lengths(1024, np.random.default_rng(0))
out = {}
for name, mk in (("b32", lambda: P.Plan(32, 1024, 1024, device=d)), ("c2", lambda: P.Plan(1024, 256, 256, device=d)),
                 ("c4", lambda: P.Plan(1024, 512, 512, device=d)),
                 ("c5p", lambda: P.Plan(1024, 1024, 1024, xl, yl, packed=True, device=d)), ("b1", lambda: P.Plan(1, 1024, 1024, device=d))):
    pl = mk()
    g = torch.Generator(device=d).manual_seed(2)
    shape = (pl.packed_floats,) if pl.packed else (pl.B, pl.N, pl.M)
    theta = torch.rand(shape, generator=g, device=d)
    A = -torch.rand(shape, generator=g, device=d)
    Et = torch.ones(pl.B, device=d)
    Vt, Q = ops.sq_forward(pl, theta, A)
    out[name] = (round(timeit(lambda: ops.sq_forward(pl, theta, A)), 4), round(timeit(lambda: ops.sq_backward(pl, Et, Q)), 4))
print(os.environ.get("LIB", "default"), json.dumps(out))
