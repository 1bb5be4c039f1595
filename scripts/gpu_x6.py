"""GPU-box experiment: L2 prefetch of wide row pieces ahead of the chained forward's 16 x 16 boxes
(B200DP_EXPERIMENTS build, env B200DP_X_PF=tiles,dist)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops
from gpu_x1 import timeit
d = torch.device("cuda:0")
for B, N, M in ((1024, 256, 256), (1024, 512, 512)):
    g = torch.Generator(device=d).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=d)
    A = -torch.rand(B, N, M, generator=g, device=d)
    ref = None
    for pf in ("", "4,4", "4,8", "8,8", "8,16", "16,16", "16,24", "16,32", "2,4"):
        if pf:
            os.environ["B200DP_X_PF"] = pf
        else:
            os.environ.pop("B200DP_X_PF", None)
        f = timeit(lambda: ops.forward_pass(theta, A, "nw"))
        Vt, Q = ops.forward_pass(theta, A, "nw")
        if ref is None:
            ref = Q.clone()
        print(json.dumps({"B": B, "M": M, "pf": pf, "fwd_ms": round(f, 4), "dQ": (Q - ref).abs().max().item()}), flush=True)
