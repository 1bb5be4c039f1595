#!/usr/bin/env python
"""Forward + backward on a ragged batch (Zipf lengths, multiples of 64) with a device
synchronisation after each pass: the quick check for the hand-off kernels on many short
pairs per CTA.  usage: python scripts/gpu_ragged_check.py B kmax seed [flags]"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepblast_b200 import ops
B, kmax, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = torch.device("cuda:0")
rng = np.random.default_rng(seed)
k = np.arange(1, kmax + 1); pk = (1.0 / k) / (1.0 / k).sum()
xl = 64 * rng.choice(k, size=B, p=pk); yl = 64 * rng.choice(k, size=B, p=pk)
N = M = 64 * kmax
g = torch.Generator(device=dev).manual_seed(2)
theta = torch.rand(B, N, M, generator=g, device=dev); A = -torch.rand(B, N, M, generator=g, device=dev)
xlen = torch.tensor(xl, dtype=torch.int32, device=dev); ylen = torch.tensor(yl, dtype=torch.int32, device=dev)
try:
    Vt, Q = ops.forward_pass(theta, A, "nw", xlen, ylen, flags=flags)
    torch.cuda.synchronize()
    print("fwd ok", float(Vt.sum()))
except Exception as e:
    print("FWD FAIL", str(e)[:80]); sys.exit(1)
try:
    E = ops.backward_pass(torch.ones(B, device=dev), Q, "nw", xlen, ylen, N=N, flags=flags)
    torch.cuda.synchronize()
    print("bwd ok", float(E.sum()))
except Exception as e:
    print("BWD FAIL", str(e)[:80]); sys.exit(2)
