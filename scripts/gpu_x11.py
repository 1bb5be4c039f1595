"""GPU-box experiment: alternative builds (LIB=path): chained forward at C2 / C3 / C4 and the adjoint pair."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import _lib
if os.environ.get("LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["LIB"])
import torch
from deepblast_b200 import ops
from gpu_x1 import timeit
d = torch.device("cuda:0")
out = {}
for mode, B, N, M in (("nw", 1024, 256, 256), ("sw", 1024, 256, 256), ("nw", 1024, 512, 512)):
    g = torch.Generator(device=d).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=d)
    A = -torch.rand(B, N, M, generator=g, device=d)
    out[f"{mode}{N}"] = round(timeit(lambda: ops.forward_pass(theta, A, mode), it=20), 4)
print(os.environ.get("LIB", "default"), json.dumps(out))
