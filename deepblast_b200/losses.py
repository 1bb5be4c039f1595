"""MatrixCrossEntropy with the call signature of deepblast.losses.MatrixCrossEntropy
(deepblast/losses.py:9-48; used by trainer.py:154-171 on predA = decode(theta, A)), fused:
one kernel reduces every pair's masked mean in one pass (reading Ypred in place, e.g. the
strided view of the padded E that `decode` returns), one kernel writes the gradient
(SURVEY.md section 8f, row 2).  CUDA tensors only; no CPU path."""
import torch

from . import _lib
from .ops import _ptr


class _MatrixCrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Ypred, Ytrue, G, xlen, ylen):
        B, N, M = Ypred.shape
        dev = Ypred.device
        yp = Ypred.detach()
        if yp.stride(2) != 1:
            yp = yp.contiguous()
        with torch.cuda.device(dev):
            pair_loss = torch.empty(B, dtype=torch.float32, device=dev)
            pair_count = torch.empty(B, dtype=torch.float32, device=dev)
            rc = _lib.lib().b200dp_mxent_fwd(_ptr(Ytrue), _ptr(yp), yp.stride(0), yp.stride(1), _ptr(G),
                                             _ptr(xlen), _ptr(ylen), B, N, M, _ptr(pair_loss), _ptr(pair_count),
                                             torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(rc, "b200dp_mxent_fwd")
        ctx.save_for_backward(yp, Ytrue, G, xlen, ylen, pair_count)
        return pair_loss.sum()

    @staticmethod
    def backward(ctx, gout):
        yp, Ytrue, G, xlen, ylen, pair_count = ctx.saved_tensors
        B, N, M = yp.shape
        dev = yp.device
        gout = gout.detach().to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(dev):
            grad = torch.empty((B, N, M), dtype=torch.float32, device=dev)
            rc = _lib.lib().b200dp_mxent_bwd(_ptr(Ytrue), _ptr(yp), yp.stride(0), yp.stride(1), _ptr(G),
                                             _ptr(xlen), _ptr(ylen), B, N, M, _ptr(pair_count), _ptr(gout),
                                             _ptr(grad), torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(rc, "b200dp_mxent_bwd")
        return grad, None, None, None, None


class MatrixCrossEntropy:
    def __call__(self, Ytrue, Ypred, x_len, y_len, G):
        """Same arguments as deepblast.losses.MatrixCrossEntropy.__call__: Ytrue, Ypred, G
        [B, N, M]; x_len, y_len sequences / tensors of B lengths.  Returns the scalar loss
        (differentiable with respect to Ypred)."""
        if not Ypred.is_cuda:
            raise RuntimeError("Ypred must be a CUDA tensor (deepblast_b200 has no CPU path)")
        if Ypred.dtype != torch.float32:
            raise TypeError("CUDA variant only supports torch.float32 type")
        dev = Ypred.device
        B = Ypred.shape[0]
        Ytrue = Ytrue.to(device=dev, dtype=torch.float32).contiguous()
        G = None if G is None else G.to(device=dev, dtype=torch.float32).contiguous()
        xlen = torch.as_tensor(x_len, dtype=torch.int32).reshape(B).to(dev)
        ylen = torch.as_tensor(y_len, dtype=torch.int32).reshape(B).to(dev)
        return _MatrixCrossEntropyFn.apply(Ypred, Ytrue, G, xlen, ylen)
