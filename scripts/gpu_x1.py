"""GPU-box experiment: fwd3 with NCH = 1 / 2 (B200DP_EXPERIMENTS build, env B200DP_X_NCH) and the strip-queue
backward grid sweep at C2 / C3 / C4."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops, plan as P
d = torch.device("cuda:0")


def timeit(fn, it=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


for mode, B, N, M in (("nw", 1024, 256, 256), ("sw", 1024, 256, 256), ("nw", 1024, 512, 512), ("nw", 592, 256, 256)):
    g = torch.Generator(device=d).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=d)
    A = -torch.rand(B, N, M, generator=g, device=d)
    Et = torch.ones(B, device=d)
    ref = None
    for nch in (1, 2):
        for ring in (3, 4):
            os.environ["B200DP_X_NCH"] = str(nch)
            fl = ring << 24
            try:
                f = timeit(lambda: ops.forward_pass(theta, A, mode, flags=fl))
                Vt, Q = ops.forward_pass(theta, A, mode, flags=fl)
                if ref is None:
                    ref = (Vt.clone(), Q.clone())
                dv = (Vt - ref[0]).abs().max().item()
                dq = (Q - ref[1]).abs().max().item()
                print(json.dumps({"mode": mode, "B": B, "N": N, "M": M, "nch": nch, "ring": ring, "fwd_ms": round(f, 4),
                                  "dVt": dv, "dQ": dq}), flush=True)
            except Exception as e:
                print("ERR", mode, B, N, M, nch, ring, repr(e)[:200], flush=True)
    os.environ["B200DP_X_NCH"] = "1"
    Vt, Q = ops.forward_pass(theta, A, mode)
    b3 = timeit(lambda: ops.backward_pass(Et, Q, mode, N=N))
    print(json.dumps({"mode": mode, "B": B, "N": N, "M": M, "bwd_default_ms": round(b3, 4)}), flush=True)
    if mode == "nw":
        for per_sm in (9, 10, 11, 12, 13):
            W = 148 * per_sm
            pl = P.Plan(B, N, M, device=d, resident_warps=W)
            Vs, Qs = ops.sq_forward(pl, theta, A)
            for ring in (2, 3):
                if ring == 3 and per_sm > 9:
                    continue
                fl = (ring << 24) | (W << 8)
                tb = timeit(lambda: ops.sq_backward(pl, Et, Qs, flags=fl), it=10, warm=3)
                print(json.dumps({"B": B, "N": N, "M": M, "sq_bwd": 1, "per_sm": per_sm, "ring": ring, "ms": round(tb, 4)}), flush=True)
    del theta, A, Q
