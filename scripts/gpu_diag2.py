import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepblast_b200 import ops
dev = torch.device("cuda:0")
KN = ["B200DP_PROMO", "B200DP_DIAG", "B200DP_PFW", "B200DP_PFD", "B200DP_DBG", "B200DP_RING", "B200DP_NCH"]
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
def run(B, N, M, combos):
    g = torch.Generator(device=dev).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=dev); A = -torch.rand(B, N, M, generator=g, device=dev)
    print(f"== B={B} {N}x{M}", flush=True)
    for kv in combos:
        for k in KN: os.environ.pop(k, None)
        os.environ.update({k: str(v) for k, v in kv.items()})
        ms = timeit(lambda: ops.forward_pass(theta, A, "nw"))
        print("fwd %-40s %.3f ms  %.0f GB/s" % (" ".join(f"{k[7:]}={v}" for k, v in kv.items()) or "default", ms, B*N*M*20/ms/1e6), flush=True)
    for k in KN: os.environ.pop(k, None)
c = [{}] + [{"B200DP_PROMO": v} for v in (0, 1, 2, 3)] + [{"B200DP_PROMO": 3, "B200DP_RING": 4}, {"B200DP_PROMO": 3, "B200DP_DIAG": 2}, {"B200DP_PROMO": 0, "B200DP_DIAG": 2}]
run(1024, 256, 256, c)
run(4096, 256, 256, c[:5])
run(1024, 512, 512, c[:5])
