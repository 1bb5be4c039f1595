#!/usr/bin/env python
"""A training-shaped step through the public API (trainer.py:185-199): decode (forward +
backward under create_graph), a loss on the expected alignment, loss.backward() (the two
adjoint sweeps).  usage: python scripts/gpu_train_step.py [B N M]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepblast_b200 import _lib  # noqa: E402
if os.environ.get("LIB"):            # alternative build of the library
    _lib.LIB_PATH = os.path.abspath(os.environ["LIB"])
from deepblast_b200.nw_cuda import NeedlemanWunschDecoder  # noqa: E402
from deepblast_b200.losses import MatrixCrossEntropy  # noqa: E402

B, N, M = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (1024, 256, 256)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(2)
theta = torch.rand(B, N, M, generator=g, device=dev).requires_grad_()
A = (-torch.rand(B, N, M, generator=g, device=dev)).requires_grad_()
Ytrue = (torch.rand(B, N, M, generator=g, device=dev) < 0.01).float()
G = torch.ones(B, N, M, device=dev)
xlen = [N] * B
ylen = [M] * B
dec = NeedlemanWunschDecoder('softmax')
lossf = MatrixCrossEntropy()


def step():
    theta.grad = None
    predA = dec.decode(theta, A)
    loss = lossf(Ytrue, predA, xlen, ylen, G)
    loss.backward()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 10
for _ in range(K):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print("B=%d %dx%d training-shaped step (decode + MatrixCrossEntropy + backward): %.3f ms, %.1f G cell/s over the 4 sweeps, loss %.5f"
      % (B, N, M, ms, B * N * M / ms / 1e6, float(loss)))
