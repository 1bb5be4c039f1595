"""GPU-box experiment: would two warps per pair (each half of the columns, chained) beat one warp per pair?
Emulated by the chained forward on 2048 pairs of 256 x 128 (the per-warp work of such a split of C2) and
on 2048 pairs of 512 x 256 (C4)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops
from gpu_x1 import timeit
d = torch.device("cuda:0")
for B, N, M in ((1024, 256, 256), (2048, 256, 128), (4096, 256, 64), (1024, 512, 512), (2048, 512, 256)):
    g = torch.Generator(device=d).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=d)
    A = -torch.rand(B, N, M, generator=g, device=d)
    for ring in (0, 3, 4):
        t = timeit(lambda: ops.forward_pass(theta, A, "nw", flags=ring << 24), it=20)
        print(json.dumps({"B": B, "N": N, "M": M, "ring": ring, "fwd_ms": round(t, 4)}), flush=True)
