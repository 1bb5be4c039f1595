"""GPU-box experiment: the adjoint pair at C2 / C4: chained kernels (adjoint_pair_fast) against the strip-queue
kernels (sq_adjoint_forward / sq_adjoint_backward)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops, plan as P, _lib
from gpu_x1 import timeit
d = torch.device("cuda:0")
L = _lib.lib()
for B, N, M in ((1024, 256, 256), (1024, 512, 512)):
    g = torch.Generator(device=d).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=d)
    A = -torch.rand(B, N, M, generator=g, device=d)
    Zt = torch.randn(B, N, M, generator=g, device=d)
    Et = torch.ones(B, device=d)
    pl = P.Plan(B, N, M, device=d)
    Vt, Q5 = ops.forward_pass(theta, A, "nw")                 # strip-major 5-D view (chained forward)
    E = ops.sq_backward(pl, Et, Q5, "nw")                     # interior [B, N, M]
    Vtd0, Ed0 = ops.adjoint_pair_fast(Q5, None, Zt, None, interior=True, Ei=E, interior_out=True, dims=(B, N, M))
    Vtd1, QdE = ops.sq_adjoint_forward(pl, Q5, Zt, None, E)
    Ed1 = ops.sq_adjoint_backward(pl, Q5, QdE)
    torch.cuda.synchronize()
    print("max |dVtd|", (Vtd0 - Vtd1).abs().max().item(), "max |dEd|", (Ed0 - Ed1).abs().max().item(), "scale", Ed0.abs().max().item())
    st = torch.cuda.current_stream().cuda_stream
    QdE3 = ops.q_empty(B, N, M, d)
    Vtd3 = torch.empty(B, device=d)
    Edi = torch.empty(B, N, M, device=d)
    t_f3 = timeit(lambda: L.b200dp_adj_fwd3(Q5.data_ptr(), Zt.data_ptr(), None, E.data_ptr(), Vtd3.data_ptr(), QdE3.data_ptr(), B, N, M, 0, st))
    t_b3 = timeit(lambda: L.b200dp_adj_bwd3(Q5.data_ptr(), QdE3.data_ptr(), None, Edi.data_ptr(), B, N, M, 0, st))
    t_fs = timeit(lambda: ops.sq_adjoint_forward(pl, Q5, Zt, None, E))
    t_bs = timeit(lambda: ops.sq_adjoint_backward(pl, Q5, QdE))
    print(json.dumps({"B": B, "N": N, "M": M, "adj_fwd_chained": round(t_f3, 4), "adj_bwd_chained": round(t_b3, 4),
                      "adj_fwd_sq": round(t_fs, 4), "adj_bwd_sq": round(t_bs, 4)}), flush=True)
