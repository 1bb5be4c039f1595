"""GPU-box timing of the batched traceback (B pairs N x M) next to the sweeps it follows."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import _lib
if os.environ.get('LIB'):
    _lib.LIB_PATH = os.path.abspath(os.environ['LIB'])
from deepblast_b200 import ops
from deepblast_b200.nw_cuda import NeedlemanWunschDecoder
from gpu_x1 import timeit
d = torch.device("cuda:0")
L = _lib.lib()
for B, N, M in ((64, 256, 256), (1, 1024, 1024), (32, 1024, 1024), (1024, 256, 256)):
    g = torch.Generator(device=d).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=d).requires_grad_()
    A = (-torch.rand(B, N, M, generator=g, device=d)).requires_grad_()
    aln = NeedlemanWunschDecoder('softmax').decode(theta, A).detach()
    cap = 2 * (N + M) + 8
    out = torch.empty((B, cap, 3), dtype=torch.int32, device=d)
    ln = torch.empty(B, dtype=torch.int32, device=d)
    st = torch.cuda.current_stream().cuda_stream
    f = lambda: L.b200dp_traceback(aln.data_ptr(), aln.stride(0), aln.stride(1), aln.stride(2), None, None, B, N, M, 1,
                                   out.data_ptr(), cap, ln.data_ptr(), st)
    t = timeit(f, it=20)
    print(json.dumps({"B": B, "N": N, "M": M, "traceback_ms": round(t, 4), "mean_len": float(ln.float().mean())}), flush=True)
