"""GPU-box experiment: the chained forward's memory skeleton (LIB=scripts/_bin/libb200dp_dbg8.so, see gpu_x16.py)
and the full kernel against the depth of the operand ring and the number of CTAs at C2 / C4."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import _lib
if os.environ.get("LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["LIB"])
import torch
from deepblast_b200 import ops
from gpu_x16 import timeit
d = torch.device("cuda:0")
for B, N, M in ((1024, 256, 256), (1024, 512, 512), (2048, 256, 128)):
    g = torch.Generator(device=d).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=d)
    A = -torch.rand(B, N, M, generator=g, device=d)
    out = {}
    for ring in (3, 4, 6, 8):
        out["r%d" % ring] = round(timeit(lambda: ops.forward_pass(theta, A, "nw", flags=ring << 24), it=20), 4)
    print(os.path.basename(os.environ.get("LIB", "default")), B, N, M, json.dumps(out), flush=True)
