"""GPU-box experiment: chunk size of the host pipelines (b200dp_decode_host / b200dp_align_host) at C2."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops
from deepblast_b200.align import HostAligner
d = torch.device("cuda:0")
B, N, M = 1024, 256, 256
th = torch.rand(B, N, M).pin_memory()
A = (-torch.rand(B, N, M)).pin_memory()
Vt_h = torch.empty(B, pin_memory=True)
E_h = torch.empty(B, N + 2, M + 2, pin_memory=True)


def timeit(fn, it=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


for cp in (16, 24, 32, 48, 64, 96, 128, 256):
    t = timeit(lambda: ops.decode_host_async(th, A, "nw", out=(Vt_h, E_h), chunk_pairs=cp))
    al = HostAligner(B, N, M, "nw", chunk_pairs=cp)
    ta = timeit(lambda: al.align(th, A))
    print(json.dumps({"chunk_pairs": cp, "decode_host_ms": round(t, 3), "align_host_ms": round(ta, 3)}), flush=True)
    del al
