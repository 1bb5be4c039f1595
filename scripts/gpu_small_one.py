"""GPU-box helper for ncu: the small-batch inference path -- cluster forward, strip-queue backward, batched walk
(64 pairs 256 x 256 through the decoder API)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops
from deepblast_b200.nw_cuda import NeedlemanWunschDecoder
d = torch.device("cuda:0")
g = torch.Generator(device=d).manual_seed(2)
theta = torch.rand(64, 256, 256, generator=g, device=d).requires_grad_()
A = (-torch.rand(64, 256, 256, generator=g, device=d)).requires_grad_()
dec = NeedlemanWunschDecoder('softmax')
for _ in range(2):
    aln = dec.decode(theta, A)
    paths = ops.traceback_batch(aln.detach())
torch.cuda.synchronize()
print("ok", len(paths), len(paths[0]))
