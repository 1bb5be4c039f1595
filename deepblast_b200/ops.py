"""Raw soft-DP passes on the GPU -- thin wrappers over the C ABI (include/b200dp.h).

One function per private pass of the reference (deepblast/nw.py:65-117, 138-175,
202-248, 270-312; deepblast/nw_cuda.py:46-165).  torch is used for device memory
and streams only.  Q / Qd travel between the passes in the engine's STRIP-MAJOR layout
 as a 5-D strided view Q5[b, k, t, j-1, s] (strip k, lane t, i.e.
lattice row i = 32k + t + 1); `q_to_reference` / `q_from_reference` convert to and
from the reference's dense padded [B, N+2, M+2, 3].  Only the x and y states of a cell
are stored; the m state is implied (Q sums to 1 over the states, Qd to 0), and q_x = -1
marks a cell whose Q is identically zero (first row / column of the sw.py lattice).
"""
import ctypes
import functools
import threading

import torch

from . import _lib

MODES = {"nw": 0, "sw": 1, 0: 0, 1: 1}
NO_CHAINED = 0x1
NO_TMA = 0x2
V1_KERNELS = 0x4
FORCE_CHAINED = 0x8


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
    """Allocate strip-major storage; returns the 5-D view [B, K, 32, M, 2] (states x, y)."""
    K, ss, ps, pad = _lib.q_layout(N, M)
    storage = torch.empty(max(B, 1) * ps + pad, dtype=torch.float32, device=device)
    return storage.as_strided((B, K, 32, M, 2), (ps, ss, 65, 64, 32), 0)


def _is_engine_q(Q, N, M):
    K, ss, ps, pad = _lib.q_layout(N, M)
    return (Q.dim() == 5 and tuple(Q.shape[1:]) == (K, 32, M, 2) and Q.dtype == torch.float32
            and tuple(Q.stride()[1:]) == (ss, 65, 64, 32) and (Q.shape[0] <= 1 or Q.stride(0) == ps)
            and Q.storage_offset() == 0 and Q.is_cuda)


def _as_engine_q(Q, N=None, M=None, kind="q"):
    """Accept the engine's 5-D view or a dense reference-layout [B,N+2,M+2,3] tensor.
    Returns (Q5, N, M)."""
    if Q.dim() == 4 and Q.shape[-1] == 3:
        return q_from_reference(Q, kind), Q.shape[1] - 2, Q.shape[2] - 2
    if Q.dim() != 5 or N is None:
        raise RuntimeError("Q must be deepblast_b200's strip-major view (pass N) or a dense "
                           "[B, N+2, M+2, 3] reference-layout tensor")
    M = Q.shape[3]
    if not _is_engine_q(Q, N, M):
        raise RuntimeError("Q is not a strip-major view produced by deepblast_b200 for this N, M "
                           "(use deepblast_b200.ops.q_from_reference to convert a dense Q)")
    return Q, N, M


def q_from_reference(Qref, kind="q"):
    """Dense reference-layout Q (kind 'q') or Qd (kind 'qd') [B,N+2,M+2,3] -> engine layout
    (5-D view).  The m state is dropped: it must be 1 - x - y for Q (a softmax; all-zero
    cells are kept as marks) and -(x + y) for Qd.  A pure layout conversion (torch indexing
    only): it also runs on CPU tensors, which is how the layout is tested without a GPU."""
    B, N2, M2, _ = Qref.shape
    N, M = N2 - 2, M2 - 2
    Q5 = q_empty(B, N, M, Qref.device)
    K = Q5.shape[1]
    rows = torch.zeros((B, K * 32, M, 2), dtype=torch.float32, device=Qref.device)
    inner = Qref[:, 1:N + 1, 1:M + 1].float()
    xy = inner[..., 0::2]                                    # states x (0) and y (2)
    if kind == "q":
        zero = (inner == 0).all(dim=-1, keepdim=True)        # sw.py first row / column
        # the implied m state (1 - x) - y must not go negative by a rounding of x + y
        xy = torch.stack([xy[..., 0], torch.minimum(xy[..., 1], 1.0 - xy[..., 0])], dim=-1)
        xy = torch.where(zero, torch.full_like(xy, -1.0), xy)
    rows[:, :N] = xy
    Q5.copy_(rows.view(B, K, 32, M, 2))
    return Q5


def q_to_reference(Q5, N, kind="q"):
    """Engine layout -> dense reference-layout [B,N+2,M+2,3] with the implied m state and
    the implicit borders made explicit: zeros, and Q[N+1, M+1, :] = 1 (nw.py:51) for
    kind 'q'; kind 'qd' (the adjoint's Qd) has m = -(x + y) and an all-zero border."""
    B, K, _, M, _ = Q5.shape
    out = torch.zeros((B, N + 2, M + 2, 3), dtype=torch.float32, device=Q5.device)
    xy = Q5.reshape(B, K * 32, M, 2)[:, :N]
    x, y = xy[..., 0], xy[..., 1]
    if kind == "q":
        zero = x < 0
        m = (1.0 - x) - y
        inner = torch.stack([x, m, y], dim=-1)
        inner = torch.where(zero.unsqueeze(-1), torch.zeros_like(inner), inner)
        out[:, N + 1, M + 1] = 1.0
    else:
        inner = torch.stack([x, -(x + y), y], dim=-1)
    out[:, 1:N + 1, 1:M + 1] = inner
    return out


def forward_pass(theta, A, mode="nw", xlen=None, ylen=None, flags=0):
    """theta, A [B,N,M] -> (Vt [B], Q strip-major view).  nw.py:65-117."""
    _check_in("theta", theta)
    _check_in("A", A, theta.shape)
    B, N, M = theta.shape
    theta = theta.detach().contiguous()
    A = A.detach().contiguous()
    xlen, ylen = _lens(xlen, ylen, B, N, M, theta.device)
    with torch.cuda.device(theta.device):
        Q = q_empty(B, N, M, theta.device)
        # (with lengths an empty pair has no strips: its score stays 0, nothing to sum)
        Vt = (torch.zeros if xlen is not None else torch.empty)(B, dtype=torch.float32, device=theta.device)
        rc = _lib.lib().b200dp_fwd(_ptr(theta), _ptr(A), _ptr(Q), _ptr(Vt), _ptr(xlen), _ptr(ylen),
                                   B, N, M, MODES[mode], flags, _stream(theta))
        _lib.check(rc, "b200dp_fwd")
    return Vt, Q


def backward_pass(Et, Q, mode="nw", xlen=None, ylen=None, flags=0, N=None, keep_interior=False):
    """Et [B] (any stride), Q (strip-major view + N, or dense reference layout)
    -> E [B,N+2,M+2].  nw.py:138-175, 347-352.  keep_interior=True returns (E, Ei) with Ei a
    contiguous copy of E[:, 1:-1, 1:-1] written by the sweep itself where the chained kernel
    takes the batch (None otherwise): the operand the chained adjoint forward sweep reads."""
    Q, N, M = _as_engine_q(Q, N)
    B = Q.shape[0]
    _check_in("Et", Et, (B,))
    Et = Et.detach()
    xlen, ylen = _lens(xlen, ylen, B, N, M, Q.device)
    if keep_interior and xlen is None and B > 0:
        with torch.cuda.device(Q.device):
            E = torch.empty((B, N + 2, M + 2), dtype=torch.float32, device=Q.device)
            Ei = None
            wrote = _lib._i(0)
            if _lib.lib().b200dp_adj3_applicable(B, N, M):
                Ei = torch.empty((B, N, M), dtype=torch.float32, device=Q.device)
            rc = _lib.lib().b200dp_bwd_keep_interior(_ptr(Et), Et.stride(0), _ptr(Q), _ptr(E), _ptr(Ei),
                                                     ctypes.byref(wrote), B, N, M, MODES[mode], flags, _stream(Q))
            _lib.check(rc, "b200dp_bwd_keep_interior")
        return E, (Ei if wrote.value else None)
    if keep_interior:
        return backward_pass(Et, Q, mode, xlen, ylen, flags, N), None
    with torch.cuda.device(Q.device):
        alloc = torch.zeros if xlen is not None else torch.empty
        E = alloc((B, N + 2, M + 2), dtype=torch.float32, device=Q.device)
        rc = _lib.lib().b200dp_bwd(_ptr(Et), Et.stride(0) if B > 0 else 0, _ptr(Q), _ptr(E),
                                   _ptr(xlen), _ptr(ylen), B, N, M, MODES[mode], flags, _stream(Q))
        _lib.check(rc, "b200dp_bwd")
    return E


def adjoint_forward_pass(Q, Ztheta, ZA, xlen=None, ylen=None, flags=0):
    """Q, Ztheta [B,N+2,M+2], ZA [B,N,M] -> (Vtd [B], Qd strip-major view).  nw.py:202-248."""
    B, N2, M2 = Ztheta.shape
    N, M = N2 - 2, M2 - 2
    Q, N, M = _as_engine_q(Q, N)
    _check_in("Ztheta", Ztheta, (B, N2, M2))
    _check_in("ZA", ZA, (B, N, M))
    Ztheta = Ztheta.detach().contiguous()
    ZA = ZA.detach().contiguous()
    xlen, ylen = _lens(xlen, ylen, B, N, M, Q.device)
    with torch.cuda.device(Q.device):
        Qd = q_empty(B, N, M, Q.device)
        Vtd = torch.empty(B, dtype=torch.float32, device=Q.device)
        rc = _lib.lib().b200dp_adj_fwd(_ptr(Q), _ptr(Ztheta), _ptr(ZA), _ptr(Vtd), _ptr(Qd),
                                       _ptr(xlen), _ptr(ylen), B, N, M, flags, _stream(Q))
        _lib.check(rc, "b200dp_adj_fwd")
    return Vtd, Qd


def adjoint_backward_pass(E, Q, Qd, xlen=None, ylen=None, flags=0):
    """E [B,N+2,M+2], Q, Qd -> Ed [B,N+2,M+2].  nw.py:270-312."""
    B, N2, M2 = E.shape
    N, M = N2 - 2, M2 - 2
    Q, _, _ = _as_engine_q(Q, N)
    Qd, _, _ = _as_engine_q(Qd, N, kind="qd")
    _check_in("E", E, (B, N2, M2))
    E = E.detach().contiguous()
    xlen, ylen = _lens(xlen, ylen, B, N, M, Q.device)
    with torch.cuda.device(Q.device):
        alloc = torch.zeros if xlen is not None else torch.empty
        Ed = alloc((B, N2, M2), dtype=torch.float32, device=Q.device)
        rc = _lib.lib().b200dp_adj_bwd(_ptr(E), _ptr(Q), _ptr(Qd), _ptr(Ed),
                                       _ptr(xlen), _ptr(ylen), B, N, M, flags, _stream(Q))
        _lib.check(rc, "b200dp_adj_bwd")
    return Ed


def adjoint_pair_fast(Q, E, Ztheta, ZA=None, N=None, flags=0, interior=False, Ei=None, interior_out=False, dims=None):
    """Both adjoint sweeps on the chained kernels (large batches of equal-size lattices):
    Q (strip-major), E [B,N+2,M+2], Ztheta [B,N+2,M+2] (or, with interior=True, its interior
    [B,N,M]; None = zeros), ZA [B,N,M] or None (= zeros) -> (Vtd [B], Ed [B,N+2,M+2]), or
    None when the shape is not taken (use adjoint_forward_pass / adjoint_backward_pass).
    The forward sweep multiplies Qd by E on the fly (it reads the interiors of Ztheta and E
    as contiguous [B,N,M] tensors through TMA), so the backward sweep needs Q and that
    product only."""
    # (dims = (B, N, M) with E = None: the caller holds the interior Ei only, no padded E exists)
    if dims is not None:
        B, N, M = dims
        N2, M2 = N + 2, M + 2
    else:
        B, N2, M2 = E.shape
        N, M = N2 - 2, M2 - 2
    if Q.dim() != 5 or not _is_engine_q(Q, N, M) or not Q.is_cuda or (flags & NO_CHAINED):
        return None
    with torch.cuda.device(Q.device):
        if not _lib.lib().b200dp_adj3_applicable(B, N, M):
            return None
        if E is not None:
            _check_in("E", E, (B, N2, M2))
        if Ztheta is None:
            zt = torch.zeros((B, N, M), dtype=torch.float32, device=Q.device)
        elif interior:
            _check_in("Ztheta", Ztheta, (B, N, M))
            zt = Ztheta.detach().contiguous()
        else:
            _check_in("Ztheta", Ztheta, (B, N2, M2))
            zt = Ztheta.detach()[:, 1:-1, 1:-1].contiguous()
        e = Ei if Ei is not None else E.detach()[:, 1:-1, 1:-1].contiguous()
        za = None
        if ZA is not None:
            _check_in("ZA", ZA, (B, N, M))
            za = ZA.detach().contiguous()
        if SQ_MODE != "never" and M % 4 == 0 and not flags:
            # the adjoint FORWARD sweep on the strip-queue kernel (more warps per SM than one per pair; measured
            # on B200: 1024 x 256^2 0.280 against 0.306 ms, 1024 x 512^2 1.03 against 1.12; the same values bit
            # for bit, the same Qd * E layout), the adjoint BACKWARD stays on the chained kernel (0.224 / 0.230)
            from . import plan as _plan
            Vtd, QdE = sq_adjoint_forward(_plan.get_plan(B, N, M, None, None, False, Q.device), Q, zt, za, e)
        else:
            QdE = q_empty(B, N, M, Q.device)
            Vtd = torch.empty(B, dtype=torch.float32, device=Q.device)
            rc = _lib.lib().b200dp_adj_fwd3(_ptr(Q), _ptr(zt), _ptr(za), _ptr(e), _ptr(Vtd), _ptr(QdE),
                                            B, N, M, flags, _stream(Q))
            _lib.check(rc, "b200dp_adj_fwd3")
        # interior_out: Ed[:, 1:-1, 1:-1] as a contiguous [B, N, M] tensor, no padded Ed at all
        Ed = None if interior_out else torch.empty((B, N2, M2), dtype=torch.float32, device=Q.device)
        Edi = torch.empty((B, N, M), dtype=torch.float32, device=Q.device) if interior_out else None
        rc = _lib.lib().b200dp_adj_bwd3(_ptr(Q), _ptr(QdE), _ptr(Ed), _ptr(Edi), B, N, M, flags, _stream(Q))
        _lib.check(rc, "b200dp_adj_bwd3")
    return Vtd, (Edi if interior_out else Ed)


# ---- strip-queue family: any batch shape, dense or packed (csrc/softdp_sq.cuh) ---------------
SQ_RING_SHIFT = 24
CTAS_SHIFT = 8


def _sq_flags(plan, flags, fwd):
    """The plan's grid size unless the caller forces one."""
    if (flags >> CTAS_SHIFT) & 0xFFFF:
        return flags
    return flags | ((plan.grid_fwd if fwd else plan.grid_bwd) << CTAS_SHIFT)


def _sq_ws(plan, t):
    from . import plan as _plan
    stream = _stream(t)
    return _plan.workspace(t.device, stream, plan.ws_bytes), stream


def _sq_check_operand(name, t, plan):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (deepblast_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise TypeError("CUDA variant only supports torch.float32 type")
    if t.numel() < plan.packed_floats:
        raise RuntimeError(f"{name}: {t.numel()} elements, the plan's layout needs {plan.packed_floats}")
    if t.device != plan.device:
        raise RuntimeError(f"{name} is on {t.device}, the plan was built for {plan.device}")


# Which batches the autograd functions send to the strip-queue kernels when the caller gave no
# plan: "auto" (ragged batches, and equal-size batches too small for the chained kernels),
# "always" (every shape the kernels take), "never" (the round-1 kernels only).
SQ_MODE = "auto"
SQ_TMA_OPERANDS = True       # dense plans: stage theta / A with TMA boxes (False: the LDGSTS path, as for packed plans)
CLUSTER = True               # small dense batches of equal-size pairs: the cluster FORWARD kernel (DSMEM hand-off)
CLUSTER_BWD = False          # the cluster backward is 5-20 % slower than the strip-queue backward on B200: not dispatched


@functools.lru_cache(maxsize=256)
def _cl_size(dev_index, B, N, M):
    with torch.cuda.device(dev_index):
        return int(_lib.lib().b200dp_cl_applicable(B, N, M))


def _use_cluster(plan, flags):
    """Cluster size for this plan, or 0: dense, equal-size, bound by one pair's dependency chain."""
    if not CLUSTER or plan.packed or plan.ragged or flags or plan.device.type != "cuda":
        return 0
    return _cl_size(plan.device.index, plan.B, plan.N, plan.M)
_sm_count = {}


def route_plan(theta, xlen=None, ylen=None):
    """The plan.Plan the batch should run with, or None for the chained / hand-off / general kernels."""
    from . import plan as _plan
    if SQ_MODE == "never" or not theta.is_cuda or theta.dim() != 3:
        return None
    B, N, M = theta.shape
    if B == 0 or M % 4 != 0:
        return None
    if SQ_MODE == "auto" and xlen is None and ylen is None:
        sms = _sm_count.get(theta.device.index)
        if sms is None:
            sms = _sm_count[theta.device.index] = torch.cuda.get_device_properties(theta.device).multi_processor_count
        if B >= 4 * sms and N >= 32 and M >= 64 and M % 32 == 0:
            return None                      # large equal-size batch: the chained kernels
        if B > 64 and N < 512:
            return None                      # many short pairs, too few for chaining: the hand-off kernels
                                             # (measured on B200: 256 x 256^2 104 against 92 G cell-updates/s;
                                             # 32 x 1024^2 37 against 64, 512 x 1024^2 160 against 177)
    return _plan.get_plan(B, N, M, xlen, ylen, False, theta.device)


def _sq_take_out(name, t, numel, device):
    if t.device != device or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() < numel:
        raise RuntimeError(f"out {name}: need a contiguous float32 tensor of at least {numel} elements on {device}")
    return t


def score_plan(theta, xlen=None, ylen=None):
    """The plan for a forward whose Q nobody will read, or None when the strip-queue kernels do not take
    the shape (M % 4 != 0 in the dense layout)."""
    from . import plan as _plan
    if SQ_MODE == "never" or not theta.is_cuda or theta.dim() != 3 or theta.shape[0] == 0 or theta.shape[2] % 4 != 0:
        return None
    B, N, M = theta.shape
    return _plan.get_plan(B, N, M, xlen, ylen, False, theta.device)


def sq_forward(plan, theta, A, mode="nw", need_q=True, flags=0, out=None):
    """theta, A in the plan's layout (dense [B,N,M] or a flat packed buffer) -> (Vt [B], Q flat
    strip-major buffer, or None with need_q=False: score only).  nw.py:65-117 per pair.
    out = (Vt, Q): caller-owned result buffers (pipelines that must not allocate inside their loop)."""
    theta = theta.detach().contiguous()
    A = A.detach().contiguous()
    _sq_check_operand("theta", theta, plan)
    _sq_check_operand("A", A, plan)
    with torch.cuda.device(theta.device):
        if out is not None:
            Vt = _sq_take_out("Vt", out[0], plan.B, theta.device)
            Q = _sq_take_out("Q", out[1], plan.q_floats, theta.device) if need_q else None
            if plan.has_empty:
                Vt[:plan.B].zero_()
        else:
            Q = torch.empty(plan.q_floats, dtype=torch.float32, device=theta.device) if need_q else None
            alloc = torch.zeros if plan.has_empty else torch.empty      # an empty pair scores 0 (nothing to sum)
            Vt = alloc(plan.B, dtype=torch.float32, device=theta.device)
        if _use_cluster(plan, flags):
            rc = _lib.lib().b200dp_cl_fwd(_ptr(theta), _ptr(A), _ptr(Q), _ptr(Vt), plan.B, plan.N, plan.M, MODES[mode],
                                          0, _stream(theta))
            _lib.check(rc, "b200dp_cl_fwd")
            return Vt, Q
        ws, stream = _sq_ws(plan, theta)
        if not plan.packed and SQ_TMA_OPERANDS:
            # dense [B, N, M] operands: TMA boxes instead of per-lane 16-byte copies
            rc = _lib.lib().b200dp_sq_fwd_dense(_ptr(plan.fwd_tab), plan.nstrips, _ptr(ws), _ptr(theta), _ptr(A),
                                                _ptr(Q), _ptr(Vt), plan.B, plan.N, plan.M, MODES[mode],
                                                _sq_flags(plan, flags, True), stream)
            _lib.check(rc, "b200dp_sq_fwd_dense")
        else:
            rc = _lib.lib().b200dp_sq_fwd(_ptr(plan.fwd_tab), plan.nstrips, _ptr(ws), _ptr(theta), _ptr(A),
                                          _ptr(Q), _ptr(Vt), MODES[mode], _sq_flags(plan, flags, True), stream)
            _lib.check(rc, "b200dp_sq_fwd")
    return Vt, Q


def _sq_out_like(plan, ref):
    """E / Ed in the plan's operand layout.  Dense ragged batches must read as zero outside each
    pair's corner (the gradient of cells the pair does not have); packed buffers have no such cells
    except the few pitch-padding floats per row, zeroed too so that sums over the buffer are exact."""
    if plan.packed:
        return torch.zeros(plan.packed_floats, dtype=torch.float32, device=ref.device)
    alloc = torch.zeros if plan.ragged else torch.empty
    return alloc((plan.B, plan.N, plan.M), dtype=torch.float32, device=ref.device)


def sq_backward(plan, Et, Q, mode="nw", flags=0, out=None):
    """Et [B] (any stride), Q (flat, from sq_forward) -> E in the plan's layout: the INTERIOR of
    the reference's padded E (nw.py:138-175, 339), i.e. dVt/dtheta.  out: a caller-owned buffer for E."""
    _check_in("Et", Et, (plan.B,))
    Et = Et.detach()
    with torch.cuda.device(Q.device):
        if out is not None:
            E = _sq_take_out("E", out, plan.packed_floats, Q.device)
            if plan.packed or plan.ragged:
                E[:plan.packed_floats].zero_()         # padding floats / cells outside a pair's corner read as zero
        else:
            E = _sq_out_like(plan, Q)
        if CLUSTER_BWD and _use_cluster(plan, flags):
            rc = _lib.lib().b200dp_cl_bwd(_ptr(Et), Et.stride(0) if plan.B > 0 else 0, _ptr(Q), _ptr(E), plan.B, plan.N,
                                          plan.M, MODES[mode], 0, _stream(Q))
            _lib.check(rc, "b200dp_cl_bwd")
            return E
        ws, stream = _sq_ws(plan, Q)
        rc = _lib.lib().b200dp_sq_bwd(_ptr(plan.bwd_tab), plan.nstrips, _ptr(ws), _ptr(Et),
                                      Et.stride(0) if plan.B > 0 else 0, _ptr(Q), _ptr(E), MODES[mode],
                                      _sq_flags(plan, flags, False), stream)
        _lib.check(rc, "b200dp_sq_bwd")
    return E


def sq_adjoint_forward(plan, Q, Zt, ZA=None, E=None, flags=0):
    """Q, Zt (the interior of Ztheta, plan layout), ZA (plan layout) or None, E (plan layout) or
    None -> (Vtd [B], QdE flat: Qd * E, or Qd itself without E).  nw.py:202-248 per pair."""
    Zt = Zt.detach().contiguous()
    _sq_check_operand("Ztheta", Zt, plan)
    if ZA is not None:
        ZA = ZA.detach().contiguous()
        _sq_check_operand("ZA", ZA, plan)
    if E is not None:
        E = E.detach().contiguous()
        _sq_check_operand("E", E, plan)
    with torch.cuda.device(Q.device):
        QdE = torch.empty(plan.q_floats, dtype=torch.float32, device=Q.device)
        alloc = torch.zeros if plan.has_empty else torch.empty
        Vtd = alloc(plan.B, dtype=torch.float32, device=Q.device)
        ws, stream = _sq_ws(plan, Q)
        rc = _lib.lib().b200dp_sq_adj_fwd(_ptr(plan.fwd_tab), plan.nstrips, _ptr(ws), _ptr(Q), _ptr(Zt),
                                          _ptr(ZA), _ptr(E), _ptr(Vtd), _ptr(QdE), _sq_flags(plan, flags, True), stream)
        _lib.check(rc, "b200dp_sq_adj_fwd")
    return Vtd, QdE


def sq_adjoint_backward(plan, Q, QdE, flags=0):
    """Q, QdE (= Qd * E from sq_adjoint_forward) -> Ed in the plan's layout (the interior of the
    reference's padded Ed, nw.py:270-312, 386)."""
    with torch.cuda.device(Q.device):
        Ed = _sq_out_like(plan, Q)
        ws, stream = _sq_ws(plan, Q)
        rc = _lib.lib().b200dp_sq_adj_bwd(_ptr(plan.bwd_tab), plan.nstrips, _ptr(ws), _ptr(Q), _ptr(QdE),
                                          _ptr(Ed), _sq_flags(plan, flags, False), stream)
        _lib.check(rc, "b200dp_sq_adj_bwd")
    return Ed


def sq_q_to_reference(plan, Q, b, kind="q"):
    """Pair b of a flat strip-major Q / Qd buffer -> the reference's dense padded
    [n_b+2, m_b+2, 3] (tests; pure indexing, also runs on CPU tensors)."""
    n, m = int(plan.xlen[b]), int(plan.ylen[b])
    K = (n + 31) // 32
    ss = m * 64
    v = Q.as_strided((1, K, 32, m, 2), (0, ss, 65, 64, 32), Q.storage_offset() + int(plan.q_off[b]))
    return q_to_reference(v, n, kind)[0]


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


_host_ws = {}
_host_ws_lock = threading.Lock()


def decode_host(theta_h, A_h, mode="nw", Et_h=None, chunk_pairs=None, out=None, device=None, flags=0):
    """Host-buffer form of `Decoder.decode` (nw_cuda.py:319-325): theta_h, A_h [B,N,M] fp32
    HOST tensors (pinned for full PCIe speed) -> (Vt_h [B], grad_h [B,N,M]) pinned host
    tensors, grad_h = dVt/dtheta being a view of the padded E the engine downloads.
    Uploads, sweeps and downloads of consecutive chunks overlap (b200dp_decode_host).
    Returns after the results have landed; `decode_host_async` only enqueues."""
    Vt_h, E_h = decode_host_async(theta_h, A_h, mode, Et_h, chunk_pairs, out, device, flags)
    torch.cuda.current_stream(device).synchronize()
    return Vt_h, E_h[:, 1:-1, 1:-1]


def decode_host_async(theta_h, A_h, mode="nw", Et_h=None, chunk_pairs=None, out=None, device=None, flags=0):
    """Enqueue only; returns (Vt_h, E_h padded [B,N+2,M+2]) pinned buffers that are valid
    after the current stream of `device` has been synchronised.  The copies are asynchronous: the
    CALLER keeps theta_h, A_h, Et_h and `out` alive and unmodified until that synchronisation
    (freeing pinned tensors earlier lets torch's pinned allocator hand the memory out again while
    the DMA is still in flight)."""
    for name, t in (("theta_h", theta_h), ("A_h", A_h)):
        if t.is_cuda:
            raise RuntimeError(f"{name} must be a host tensor (use Decoder.decode for CUDA tensors)")
        if t.dtype != torch.float32:
            raise TypeError("CUDA variant only supports torch.float32 type")
    if tuple(A_h.shape) != tuple(theta_h.shape) or theta_h.dim() != 3:
        raise RuntimeError("theta_h and A_h must both be [B, N, M]")
    if not torch.cuda.is_available():
        raise RuntimeError("deepblast_b200 needs a CUDA device: there is no CPU path")
    B, N, M = theta_h.shape
    theta_h = theta_h.contiguous()
    A_h = A_h.contiguous()
    if Et_h is not None:
        if Et_h.is_cuda or tuple(Et_h.shape) != (B,):
            raise RuntimeError("Et_h must be a host tensor of shape [B]")
        Et_h = Et_h.contiguous().float()
    if out is not None:
        # the C side copies B and B * (N+2) * (M+2) floats into these buffers: check before it does
        for name, t, shape in (("out[0] (Vt_h)", out[0], (B,)), ("out[1] (E_h)", out[1], (B, N + 2, M + 2))):
            if t.is_cuda or t.dtype != torch.float32 or tuple(t.shape) != shape or not t.is_contiguous():
                raise RuntimeError(f"{name} must be a contiguous float32 host tensor of shape {shape}")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if chunk_pairs is None:
        # enough chunks to hide the first upload and the last download, each still a few MB
        chunk_pairs = max(1, min(B, max(8, (6 << 20) // max(1, N * M)), (B + 9) // 10 if B >= 20 else B))
    with torch.cuda.device(dev):
        need = _lib.lib().b200dp_decode_host_workspace(N, M, chunk_pairs)
        key = (dev.index, N, M, chunk_pairs)
        with _host_ws_lock:
            ws = _host_ws.get(key)
            if ws is None or ws.numel() < need:
                _host_ws.clear()                   # one cached workspace per process is enough
                ws = torch.empty(need, dtype=torch.uint8, device=dev)
                _host_ws[key] = ws
        if out is None:
            Vt_h = torch.empty(B, dtype=torch.float32, pin_memory=True)
            E_h = torch.empty((B, N + 2, M + 2), dtype=torch.float32, pin_memory=True)
        else:
            Vt_h, E_h = out
        rc = _lib.lib().b200dp_decode_host(theta_h.data_ptr(), A_h.data_ptr(), _ptr(Et_h), Vt_h.data_ptr(),
                                           E_h.data_ptr(), B, N, M, MODES[mode], chunk_pairs, ws.data_ptr(),
                                           ws.numel(), flags, torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "b200dp_decode_host")
    return Vt_h, E_h
