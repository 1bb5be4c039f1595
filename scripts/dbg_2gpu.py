import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepblast_b200 import ops
from deepblast_b200.nw_cuda import NeedlemanWunschFunction as Fn
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
use_pg = os.environ.get("NO_PG") is None and world > 1
if use_pg:
    dist.init_process_group("nccl", device_id=dev)
B, N, M = 1024, 256, 256
theta = torch.rand(B, N, M, device=dev, requires_grad=True); A = -torch.rand(B, N, M, device=dev)
acc = {}
def wrap(name, fn):
    def f(*a, **k):
        t0 = time.perf_counter(); r = fn(*a, **k); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; return r
    return f
ops.backward_pass = wrap("ops.backward_pass", ops.backward_pass)
ops.forward_pass = wrap("ops.forward_pass", ops.forward_pass)
import deepblast_b200._functions as F
for _ in range(3):
    v = Fn.apply(theta, A, 'softmax'); g, = torch.autograd.grad(v.sum(), theta)
torch.cuda.synchronize(); acc.clear()
n = 10; tf = ts = tg = 0.0
for _ in range(n):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); v = Fn.apply(theta, A, 'softmax'); t1 = time.perf_counter()
    s = v.sum(); t2 = time.perf_counter()
    g, = torch.autograd.grad(s, theta); t3 = time.perf_counter()
    tf += t1 - t0; ts += t2 - t1; tg += t3 - t2
print(f"rank {rank} pg={use_pg} threads={torch.get_num_threads()} host ms: apply {tf/n*1e3:.3f} sum {ts/n*1e3:.3f} grad {tg/n*1e3:.3f} | inside ops: " + " ".join(f"{k} {v/n*1e3:.3f}" for k, v in acc.items()), flush=True)
if use_pg: dist.destroy_process_group()
