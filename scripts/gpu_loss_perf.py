"""GPU-box helper: the fused MatrixCrossEntropy kernels alone at C2 (LIB=path for alternative builds)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import _lib
if os.environ.get("LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["LIB"])
import torch
from deepblast_b200.losses import MatrixCrossEntropy
from gpu_x16 import timeit
d = torch.device("cuda:0")
out = {}
for B, N, M in ((1024, 256, 256), (32, 512, 512)):
    g = torch.Generator(device=d).manual_seed(2)
    pred = torch.rand(B, N, M, generator=g, device=d).requires_grad_()
    Ytrue = (torch.rand(B, N, M, generator=g, device=d) < 0.01).float()
    G = torch.ones(B, N, M, device=d)
    xlen, ylen = [N] * B, [M] * B
    lossf = MatrixCrossEntropy()
    tf = timeit(lambda: lossf(Ytrue, pred, xlen, ylen, G))
    loss = lossf(Ytrue, pred, xlen, ylen, G)

    def bwd():
        pred.grad = None
        loss.backward(retain_graph=True)
    tb = timeit(bwd)
    out["%dx%dx%d" % (B, N, M)] = {"fwd_ms": round(tf, 4), "bwd_ms": round(tb, 4), "loss": float(loss)}
print(os.path.basename(os.environ.get("LIB", "default")), json.dumps(out))
