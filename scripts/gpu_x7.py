"""GPU-box experiment: cluster kernels (DSMEM hand-off) against the strip-queue kernels on small batches of
long pairs: correctness (bit-identical Q / E expected: same arithmetic) and time."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops, plan as P, _lib
from gpu_x1 import timeit
d = torch.device("cuda:0")
L = _lib.lib()
for mode, B, N, M in (("nw", 2, 96, 128), ("nw", 3, 70, 200), ("sw", 2, 130, 260), ("nw", 32, 1024, 1024), ("sw", 32, 1024, 1024),
                      ("nw", 1, 1024, 1024), ("nw", 32, 512, 512), ("nw", 64, 256, 256), ("nw", 8, 2047 - 3, 2044)):
    g = torch.Generator(device=d).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=d)
    A = -torch.rand(B, N, M, generator=g, device=d)
    Et = torch.linspace(0.5, 1.5, B, device=d)
    pl = P.Plan(B, N, M, device=d)
    ops.CLUSTER = False
    Vt0, Q0 = ops.sq_forward(pl, theta, A, mode)
    E0 = ops.sq_backward(pl, Et, Q0, mode)
    Vs0, _ = ops.sq_forward(pl, theta, A, mode, need_q=False)
    tf0 = timeit(lambda: ops.sq_forward(pl, theta, A, mode))
    tb0 = timeit(lambda: ops.sq_backward(pl, Et, Q0, mode))
    torch.cuda.synchronize()
    for cs, W in ((0, 0), (8, 1), (8, 2), (8, 4), (4, 4)):
        fl = (cs << 4) | (W << 8)
        Q = torch.empty_like(Q0)
        Vt = torch.empty_like(Vt0)
        E = torch.full_like(E0, float("nan"))
        st = torch.cuda.current_stream().cuda_stream
        rc = L.b200dp_cl_fwd(theta.data_ptr(), A.data_ptr(), Q.data_ptr(), Vt.data_ptr(), B, N, M, ops.MODES[mode], fl, st)
        assert rc == 0, L.b200dp_last_error()
        rc = L.b200dp_cl_bwd(Et.data_ptr(), Et.stride(0), Q.data_ptr(), E.data_ptr(), B, N, M, ops.MODES[mode], fl, st)
        assert rc == 0, L.b200dp_last_error()
        Vs = torch.empty_like(Vt0)
        rc = L.b200dp_cl_fwd(theta.data_ptr(), A.data_ptr(), None, Vs.data_ptr(), B, N, M, ops.MODES[mode], fl, st)
        assert rc == 0, L.b200dp_last_error()
        torch.cuda.synchronize()
        # compare Q through the reference-layout conversion (storage between streams is never written)
        dq = max((ops.sq_q_to_reference(pl, Q, b) - ops.sq_q_to_reference(pl, Q0, b)).abs().max().item() for b in range(min(B, 4)))
        tf = timeit(lambda: L.b200dp_cl_fwd(theta.data_ptr(), A.data_ptr(), Q.data_ptr(), Vt.data_ptr(), B, N, M, ops.MODES[mode], fl, st))
        tb = timeit(lambda: L.b200dp_cl_bwd(Et.data_ptr(), Et.stride(0), Q.data_ptr(), E.data_ptr(), B, N, M, ops.MODES[mode], fl, st))
        print(json.dumps({"mode": mode, "B": B, "N": N, "M": M, "cluster": cs or L.b200dp_cl_applicable(B, N, M), "W": W,
                          "dVt": (Vt - Vt0).abs().max().item(), "dVs": (Vs - Vt0).abs().max().item(), "dQ": dq,
                          "dE": (E - E0).abs().max().item(), "fwd_ms": round(tf, 4), "bwd_ms": round(tb, 4),
                          "sq_fwd_ms": round(tf0, 4), "sq_bwd_ms": round(tb0, 4)}), flush=True)
