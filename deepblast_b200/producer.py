"""The step right before the DP: the match and gap score matrices of
`NeuralAligner.forward` (deepblast/alignment.py:122-123; also :134-135 and :162-163)

    theta = F.softplus(torch.einsum('bid,bjd->bij', zx, zy))
    A     = F.logsigmoid(torch.einsum('bid,bjd->bij', gx, gy))

as ONE batched tcgen05 GEMM launch with the activations fused into the epilogue
(csrc/softdp_gemm.cu; C ABI b200dp_theta_a), differentiable: the backward recovers the activations'
derivatives from the outputs (sigmoid(s) = 1 - exp(-theta), 1 - sigmoid(s) = 1 - exp(A)) and forms the
embedding gradients with two batched matrix products each.  CUDA tensors only; no CPU path."""
import torch

from . import _lib
from .ops import _ptr


# False: the hi/lo split happens inside the GEMM kernel (default).  True: version 1, a pre-pass writes
# bf16 copies of the embeddings into a workspace and the GEMM reads them through TMA.
PREPASS = False


def _check(name, t, shape=None):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (deepblast_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise TypeError("CUDA variant only supports torch.float32 type")
    if t.dim() != 3 or (shape is not None and tuple(t.shape) != tuple(shape)):
        raise RuntimeError(f"{name}: expected shape {shape or '[B, L, D]'}, got {tuple(t.shape)}")


def theta_a(zx, zy, gx, gy, xlen=None, ylen=None, plan=None):
    """zx, gx [B, Lx, D], zy, gy [B, Ly, D] -> (theta, A): dense [B, Lx, Ly], or -- with a packed
    plan.Plan -- flat packed buffers in the plan's layout (what the strip-queue kernels read).
    xlen / ylen (or the plan's lengths): only each pair's n_b x m_b corner is computed; the rest of a
    dense output is zero.  Not differentiable (see ThetaA for autograd)."""
    _check("zx", zx)
    B, Lx, D = zx.shape
    _check("zy", zy)
    Ly = zy.shape[1]
    _check("zy", zy, (B, Ly, D))
    _check("gx", gx, (B, Lx, D))
    _check("gy", gy, (B, Ly, D))
    if D % 64 != 0:
        raise RuntimeError("the embedding dimension must be a multiple of 64")
    dev = zx.device
    zx, zy, gx, gy = (t.detach().contiguous() for t in (zx, zy, gx, gy))
    packed = plan is not None and plan.packed
    if plan is not None:
        if plan.B != B or plan.N > Lx or plan.M > Ly:
            raise RuntimeError("the plan does not fit these embeddings")
        xlen, ylen = plan.xlen, plan.ylen
    with torch.cuda.device(dev):
        xl = None if xlen is None else torch.as_tensor(xlen, dtype=torch.int32).to(dev)
        yl = None if ylen is None else torch.as_tensor(ylen, dtype=torch.int32).to(dev)
        poff = None
        if packed:
            poff = torch.as_tensor(plan.pair_off[:B].copy(), dtype=torch.int64).to(dev)
            theta = torch.zeros(plan.packed_floats, dtype=torch.float32, device=dev)
            A = torch.zeros(plan.packed_floats, dtype=torch.float32, device=dev)
        else:
            alloc = torch.zeros if xl is not None or yl is not None else torch.empty
            theta = alloc((B, Lx, Ly), dtype=torch.float32, device=dev)
            A = alloc((B, Lx, Ly), dtype=torch.float32, device=dev)
        need = _lib.lib().b200dp_theta_a_workspace(B, Lx, Ly, D) if PREPASS else 0
        ws = torch.empty(need, dtype=torch.uint8, device=dev) if PREPASS else None
        rc = _lib.lib().b200dp_theta_a(_ptr(zx), _ptr(zy), _ptr(gx), _ptr(gy), B, Lx, Ly, D, _ptr(xl), _ptr(yl),
                                       _ptr(poff), _ptr(theta), _ptr(A), _ptr(ws), need,
                                       torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "b200dp_theta_a")
    return theta, A


class ThetaA(torch.autograd.Function):
    """(theta, A) = ThetaA.apply(zx, zy, gx, gy), dense [B, Lx, Ly], differentiable with respect to
    all four embeddings."""

    @staticmethod
    def forward(ctx, zx, zy, gx, gy):
        theta, A = theta_a(zx, zy, gx, gy)
        ctx.save_for_backward(zx, zy, gx, gy, theta, A)
        return theta, A

    @staticmethod
    def backward(ctx, gtheta, gA):
        zx, zy, gx, gy, theta, A = ctx.saved_tensors
        dzx = dzy = dgx = dgy = None
        if gtheta is not None:
            ds = gtheta * (1.0 - torch.exp(-theta))          # d softplus(s) / ds = sigmoid(s) = 1 - exp(-softplus(s))
            dzx, dzy = torch.bmm(ds, zy), torch.bmm(ds.transpose(1, 2), zx)
        if gA is not None:
            ds = gA * (1.0 - torch.exp(A))                    # d logsigmoid(s) / ds = 1 - sigmoid(s) = 1 - exp(logsigmoid(s))
            dgx, dgy = torch.bmm(ds, gy), torch.bmm(ds.transpose(1, 2), gx)
        return dzx, dzy, dgx, dgy
