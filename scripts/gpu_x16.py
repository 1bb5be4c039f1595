"""GPU-box experiment: what bounds the chained forward at C2?  Diagnostic builds of the library
(-DB200DP_FWD3_DBG=d, loaded with LIB=path): 8 = memory skeleton (loads, shuffle, stores, one FMA per
cell), 9 = skeleton without the Q stores, 10 = skeleton without the theta / A loads, 1 = full arithmetic
without the Q stores, 2 = full arithmetic without the loads.  Results of those builds are wrong by design."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import _lib
if os.environ.get("LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["LIB"])
import torch
from deepblast_b200 import ops


def timeit(fn, it=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


if __name__ == "__main__":
    d = torch.device("cuda:0")
    out = {}
    for B, N, M, ring in ((1024, 256, 256, 0), (1024, 512, 512, 0), (2048, 256, 128, 3), (2048, 256, 128, 4)):
        g = torch.Generator(device=d).manual_seed(2)
        theta = torch.rand(B, N, M, generator=g, device=d)
        A = -torch.rand(B, N, M, generator=g, device=d)
        t = timeit(lambda: ops.forward_pass(theta, A, "nw", flags=ring << 24), it=20)
        out["%dx%dx%d r%d" % (B, N, M, ring)] = round(t, 4)
    print(os.path.basename(os.environ.get("LIB", "default")), json.dumps(out), flush=True)
