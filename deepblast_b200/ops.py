"""Raw soft-DP passes on the GPU -- thin wrappers over the C ABI (include/b200dp.h).

One function per private pass of the reference (deepblast/nw.py:65-117, 138-175,
202-248, 270-312; deepblast/nw_cuda.py:46-165).  torch is used for device memory
and streams only.  Q / Qd are returned as strided VIEWS with the reference's
logical shape [B, N+2, M+2, 3] over anti-diagonal-major storage.
"""
import torch

from . import _lib

MODES = {"nw": 0, "sw": 1, 0: 0, 1: 1}
Q_ROW_BORDERS = 0x1
NO_TMA = 0x2


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _check_in(name, t, shape=None):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (deepblast_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise TypeError("CUDA variant only supports torch.float32 type")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")


def _lens(xlen, ylen, B, N, M, device):
    if xlen is None and ylen is None:
        return None, None
    if xlen is None:
        xlen = torch.full((B,), N, dtype=torch.int32, device=device)
    if ylen is None:
        ylen = torch.full((B,), M, dtype=torch.int32, device=device)
    xlen = xlen.to(device=device, dtype=torch.int32).contiguous()
    ylen = ylen.to(device=device, dtype=torch.int32).contiguous()
    if xlen.shape != (B,) or ylen.shape != (B,):
        raise RuntimeError("xlen / ylen must have shape [B]")
    return xlen, ylen


def q_empty(B, N, M, device):
    """Allocate anti-diagonal-major storage and return (storage, view[B,N+2,M+2,3])."""
    Lp, ND, ps, off = _lib.q_layout(N, M)
    storage = torch.empty(max(B, 1) * ps, dtype=torch.float32, device=device)
    view = storage.as_strided((B, N + 2, M + 2, 3), (ps, 3 * Lp + 1, 3 * Lp, Lp), off)
    return storage, view


def q_storage_ptr(Q, N, M):
    """Storage base pointer of a Q view produced by q_empty (validated)."""
    Lp, ND, ps, off = _lib.q_layout(N, M)
    B = Q.shape[0]
    want = (ps, 3 * Lp + 1, 3 * Lp, Lp)
    if tuple(Q.shape[1:]) != (N + 2, M + 2, 3) or (B > 1 and Q.stride(0) != ps) or \
            tuple(Q.stride()[1:]) != want[1:] or Q.dtype != torch.float32:
        raise RuntimeError(
            "Q must be the anti-diagonal-major view produced by deepblast_b200 "
            "(use deepblast_b200.ops.q_from_reference to convert a dense reference-layout Q)")
    return Q.data_ptr() - 4 * off


def q_from_reference(Qref):
    """Convert a dense reference-layout Q [B,N+2,M+2,3] to the engine's layout."""
    B, N2, M2, _ = Qref.shape
    storage, view = q_empty(B, N2 - 2, M2 - 2, Qref.device)
    view.copy_(Qref)
    return view


def forward_pass(theta, A, mode="nw", xlen=None, ylen=None, row_borders=False, flags=0):
    """theta, A [B,N,M] -> (Vt [B], Q view [B,N+2,M+2,3]).  nw.py:65-117."""
    _check_in("theta", theta)
    _check_in("A", A, theta.shape)
    B, N, M = theta.shape
    theta = theta.detach().contiguous()
    A = A.detach().contiguous()
    xlen, ylen = _lens(xlen, ylen, B, N, M, theta.device)
    with torch.cuda.device(theta.device):
        storage, Q = q_empty(B, N, M, theta.device)
        Vt = torch.empty(B, dtype=torch.float32, device=theta.device)
        fl = flags | (Q_ROW_BORDERS if row_borders else 0)
        rc = _lib.lib().b200dp_fwd(_ptr(theta), _ptr(A), _ptr(storage), _ptr(Vt), _ptr(xlen), _ptr(ylen),
                                   B, N, M, MODES[mode], fl, _stream(theta))
        _lib.check(rc, "b200dp_fwd")
    return Vt, Q


def backward_pass(Et, Q, mode="nw", xlen=None, ylen=None, flags=0):
    """Et [B] (any stride), Q view -> E [B,N+2,M+2].  nw.py:138-175, 347-352."""
    B, N2, M2, _ = Q.shape
    N, M = N2 - 2, M2 - 2
    _check_in("Et", Et, (B,))
    Et = Et.detach()
    xlen, ylen = _lens(xlen, ylen, B, N, M, Q.device)
    with torch.cuda.device(Q.device):
        alloc = torch.zeros if xlen is not None else torch.empty
        E = alloc((B, N + 2, M + 2), dtype=torch.float32, device=Q.device)
        rc = _lib.lib().b200dp_bwd(_ptr(Et), Et.stride(0) if B > 0 else 0, q_storage_ptr(Q, N, M), _ptr(E),
                                   _ptr(xlen), _ptr(ylen), B, N, M, MODES[mode], flags, _stream(Q))
        _lib.check(rc, "b200dp_bwd")
    return E


def adjoint_forward_pass(Q, Ztheta, ZA, xlen=None, ylen=None, flags=0):
    """Q view, Ztheta [B,N+2,M+2], ZA [B,N,M] -> (Vtd [B], Qd view).  nw.py:202-248."""
    B, N2, M2, _ = Q.shape
    N, M = N2 - 2, M2 - 2
    _check_in("Ztheta", Ztheta, (B, N2, M2))
    _check_in("ZA", ZA, (B, N, M))
    Ztheta = Ztheta.detach().contiguous()
    ZA = ZA.detach().contiguous()
    xlen, ylen = _lens(xlen, ylen, B, N, M, Q.device)
    with torch.cuda.device(Q.device):
        qd_storage, Qd = q_empty(B, N, M, Q.device)
        Vtd = torch.empty(B, dtype=torch.float32, device=Q.device)
        rc = _lib.lib().b200dp_adj_fwd(q_storage_ptr(Q, N, M), _ptr(Ztheta), _ptr(ZA), _ptr(Vtd),
                                       _ptr(qd_storage), _ptr(xlen), _ptr(ylen), B, N, M, flags, _stream(Q))
        _lib.check(rc, "b200dp_adj_fwd")
    return Vtd, Qd


def adjoint_backward_pass(E, Q, Qd, xlen=None, ylen=None, flags=0):
    """E [B,N+2,M+2], Q view, Qd view -> Ed [B,N+2,M+2].  nw.py:270-312."""
    B, N2, M2, _ = Q.shape
    N, M = N2 - 2, M2 - 2
    _check_in("E", E, (B, N2, M2))
    E = E.detach().contiguous()
    xlen, ylen = _lens(xlen, ylen, B, N, M, Q.device)
    with torch.cuda.device(Q.device):
        alloc = torch.zeros if xlen is not None else torch.empty
        Ed = alloc((B, N2, M2), dtype=torch.float32, device=Q.device)
        rc = _lib.lib().b200dp_adj_bwd(_ptr(E), q_storage_ptr(Q, N, M), q_storage_ptr(Qd, N, M), _ptr(Ed),
                                       _ptr(xlen), _ptr(ylen), B, N, M, flags, _stream(Q))
        _lib.check(rc, "b200dp_adj_bwd")
    return Ed


def traceback_batch(grad, xlen=None, ylen=None, variant="cuda"):
    """grad [B,N,M] (any strides) -> list of B lists of (i, j, state) tuples, exactly
    what NeedlemanWunschDecoder.traceback returns per pair (nw.py:401-444 for
    variant 'cpu', nw_cuda.py:273-317 for 'cuda').  One launch + one D2H copy."""
    if not grad.is_cuda:
        raise RuntimeError("grad must be a CUDA tensor")
    grad = grad.detach()
    if grad.dtype != torch.float32:
        grad = grad.float()      # the reference compares candidates as float32 (nw.py:427)
    B, N, M = grad.shape
    xlen, ylen = _lens(xlen, ylen, B, N, M, grad.device)
    cap = 2 * (N + M) + 8
    with torch.cuda.device(grad.device):
        out = torch.empty((B, cap, 3), dtype=torch.int32, device=grad.device)
        ln = torch.empty(B, dtype=torch.int32, device=grad.device)
        rc = _lib.lib().b200dp_traceback(_ptr(grad), grad.stride(0), grad.stride(1), grad.stride(2),
                                         _ptr(xlen), _ptr(ylen), B, N, M,
                                         {"cpu": 0, "cuda": 1}[variant], _ptr(out), cap, _ptr(ln),
                                         _stream(grad))
        _lib.check(rc, "b200dp_traceback")
        ln_h = ln.cpu().tolist()
        mx = max([l for l in ln_h if l > 0], default=0)
        out_h = out[:, :mx].cpu().numpy()
    res = []
    for b in range(B):
        if ln_h[b] == -2:
            raise IndexError("index out of range in traceback (negative wrap-around exhausted)")
        if ln_h[b] < 0:
            raise RuntimeError("b200dp_traceback: output capacity exceeded")
        res.append([tuple(int(v) for v in row) for row in out_h[b, :ln_h[b]]])
    return res
