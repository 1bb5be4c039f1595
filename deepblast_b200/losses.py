"""MatrixCrossEntropy with the call signature of deepblast.losses.MatrixCrossEntropy
(deepblast/losses.py:9-48; used by trainer.py:154-171 on predA = decode(theta, A)), fused:
one kernel reduces every pair's masked mean in one pass (reading Ypred in place, e.g. the
strided view of the padded E that `decode` returns), one kernel writes the gradient
(SURVEY.md section 8f, row 2).  CUDA tensors only; no CPU path."""
import torch

from . import _lib
from .ops import _ptr


def _fit(t, shape, name):
    if t.dim() != 3 or t.shape[0] != shape[0]:
        raise RuntimeError(f"{name} must be [B, n, m] with B = {shape[0]}, got {tuple(t.shape)}")
    if tuple(t.shape) == tuple(shape):
        return t.contiguous()
    out = torch.zeros(tuple(shape), dtype=t.dtype, device=t.device)
    n, m = min(shape[1], t.shape[1]), min(shape[2], t.shape[2])
    out[:, :n, :m] = t[:, :n, :m]
    return out


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
        if Ypred.dim() != 3:
            raise RuntimeError("Ypred must be [B, N, M]")
        B = Ypred.shape[0]
        # the kernels index all three tensors with Ypred's N and M: a differently padded Ytrue / G
        # (fine for the reference, which slices each tensor on its own, losses.py:28-40) is cut or
        # zero-padded to Ypred's shape first; cells beyond x_len / y_len are never read
        Ytrue = _fit(Ytrue.to(device=dev, dtype=torch.float32), Ypred.shape, "Ytrue")
        G = None if G is None else _fit(G.to(device=dev, dtype=torch.float32), Ypred.shape, "G")
        if len(x_len) != B or len(y_len) != B:
            raise RuntimeError(f"x_len / y_len must hold one length per pair ({B})")
        xlen = torch.as_tensor(x_len, dtype=torch.int32).reshape(B).to(dev)
        ylen = torch.as_tensor(y_len, dtype=torch.int32).reshape(B).to(dev)
        return _MatrixCrossEntropyFn.apply(Ypred, Ytrue, G, xlen, ylen)
