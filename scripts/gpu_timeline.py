"""Compact GPU/CPU timeline of a few pipelined autograd steps (torch.profiler / CUPTI)."""
import os, sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepblast_b200.nw_cuda import NeedlemanWunschFunction as Fn
dev = torch.device("cuda:0")
B, N, M = 1024, 256, 256
theta = torch.rand(B, N, M, device=dev, requires_grad=True); A = -torch.rand(B, N, M, device=dev)
def step():
    v = Fn.apply(theta, A, 'softmax'); g, = torch.autograd.grad(v.sum(), theta); return g
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(4): step()
    torch.cuda.synchronize()
evs = prof.events()
rows = []
for e in evs:
    dt = str(e.device_type)
    if "CUDA" in dt or e.name.startswith("cuda") or "softdp" in e.name or e.name in ("cudaLaunchKernel",):
        rows.append((e.time_range.start, e.time_range.end - e.time_range.start, dt.split(".")[-1], e.name[:70]))
rows.sort()
t0 = rows[0][0] if rows else 0
for s, d, k, n in rows:
    print("%9.1f us  +%8.1f  %-5s %s" % (s - t0, d, k, n))
