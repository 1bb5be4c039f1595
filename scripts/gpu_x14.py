"""GPU-box experiment: a large equal-size batch split between the chained forward (one warp per pair) and the
strip-queue forward on a second stream -- the two are latency-bound and leave each other issue slots."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops, plan as P
d = torch.device("cuda:0")


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


side = torch.cuda.Stream()
for mode, B, N, M in (("nw", 1024, 256, 256), ("nw", 1024, 512, 512)):
    g = torch.Generator(device=d).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=d)
    A = -torch.rand(B, N, M, generator=g, device=d)
    base = timeit(lambda: ops.forward_pass(theta, A, mode))
    for B3 in (592, 640, 704, 768, 832, 896):
        pl = P.Plan(B - B3, N, M, device=d)
        t3, a3, ts, as_ = theta[:B3], A[:B3], theta[B3:].contiguous(), A[B3:].contiguous()
        for per_sm in (0, 4, 6, 8):
            fl = (148 * per_sm) << 8 if per_sm else 0

            def both():
                ev = torch.cuda.Event()
                ev.record()
                with torch.cuda.stream(side):
                    side.wait_event(ev)
                    ops.sq_forward(pl, ts, as_, mode, flags=fl)
                    ev2 = torch.cuda.Event()
                    ev2.record(side)
                ops.forward_pass(t3, a3, mode)
                torch.cuda.current_stream().wait_event(ev2)
            t = timeit(both)
            print(json.dumps({"B": B, "M": M, "B3": B3, "sq_per_sm": per_sm, "ms": round(t, 4), "chained_alone_ms": round(base, 4)}), flush=True)
