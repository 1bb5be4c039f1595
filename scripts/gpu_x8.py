"""GPU-box experiment: cost of the ramp blocks.  One strip per pair (N = 32), lone warps: the sweep takes
(M + 31) steps of which 64 are in ramp blocks; the slope over M is the steady step time, the intercept the ramps."""
import os, sys, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops, plan as P
from gpu_x1 import timeit
d = torch.device("cuda:0")
ops.CLUSTER = False
res = {"fwd": [], "bwd": [], "score": []}
Ms = (64, 128, 256, 512, 1024, 2048)
for M in Ms:
    B, N = 32, 32
    g = torch.Generator(device=d).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=d)
    A = -torch.rand(B, N, M, generator=g, device=d)
    Et = torch.ones(B, device=d)
    pl = P.Plan(B, N, M, device=d)
    Vt, Q = ops.sq_forward(pl, theta, A)
    res["fwd"].append(timeit(lambda: ops.sq_forward(pl, theta, A), it=20) * 1e3)
    res["score"].append(timeit(lambda: ops.sq_forward(pl, theta, A, need_q=False), it=20) * 1e3)
    res["bwd"].append(timeit(lambda: ops.sq_backward(pl, Et, Q), it=20) * 1e3)
for k, v in res.items():
    x = np.array(Ms, float)
    y = np.array(v)
    slope, icpt = np.polyfit(x[2:], y[2:], 1)
    print(k, "us:", [round(t, 1) for t in v], " steady ns/step %.1f, intercept %.1f us (= %.0f steady steps for the 64 ramp + 31 skew steps and launch)" % (
        slope * 1e3, icpt, icpt / slope))
