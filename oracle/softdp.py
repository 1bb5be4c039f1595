"""ctypes front-end of oracle/softdp_oracle.c -- TEST INFRASTRUCTURE ONLY.

Restates deepblast/nw.py and deepblast/sw.py (reference @ ec661fa) in fp64 C.
Parity pinned against goldens generated from the imported reference
(tests/golden/make_golden.py) -- see tests/test_oracle.py.

The functions mirror the reference's private passes one to one:
  forward_pass            <- nw.py:65-117  / sw.py:65-96
  backward_pass           <- nw.py:138-175 / sw.py:118-139
  adjoint_forward_pass    <- nw.py:202-248
  adjoint_backward_pass   <- nw.py:270-312
  traceback               <- nw.py:401-444 (variant 'cpu'), nw_cuda.py:273-317 ('cuda')
and cast to the tensor dtype at the same places (nw.py:114-115, 347-352).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_softdp.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_fp = ctypes.POINTER(ctypes.c_float)
_ip = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    """Compile the oracle with gcc (few hundred ms)."""
    src = os.path.join(_HERE, "softdp_oracle.c")
    if force or not os.path.exists(_SO) or \
            os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B",
                               "liboracle_softdp.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.oracle_forward.restype = ctypes.c_double
        L.oracle_forward.argtypes = [_dp, _dp, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, _dp]
        L.oracle_backward.restype = None
        L.oracle_backward.argtypes = [ctypes.c_double, _dp, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, _dp]
        L.oracle_adjoint_forward.restype = ctypes.c_double
        L.oracle_adjoint_forward.argtypes = [_dp, _dp, _dp, ctypes.c_int,
                                             ctypes.c_int, _dp]
        L.oracle_adjoint_backward.restype = None
        L.oracle_adjoint_backward.argtypes = [_dp, _dp, _dp, ctypes.c_int,
                                              ctypes.c_int, _dp]
        L.oracle_traceback.restype = ctypes.c_int
        L.oracle_traceback.argtypes = [_fp, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, _ip, ctypes.c_int]
        L.oracle_fwd_bwd_batch_f32.restype = ctypes.c_int
        L.oracle_fwd_bwd_batch_f32.argtypes = [
            _fp, _fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            _ip, _ip, _fp, _fp, ctypes.c_int]
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=_dp):
    return a.ctypes.data_as(t)


def _i0(mode):
    return {"nw": 1, "sw": 2}[mode]


def forward_pass(theta, A, mode="nw"):
    """theta, A: [B,N,M] arrays (fp32 or fp64).  Returns (Vt[B], Q[B,N+2,M+2,3])
    in theta's dtype, computed in fp64 like nw.py:46-62."""
    theta = np.asarray(theta)
    A = np.asarray(A)
    B, N, M = theta.shape
    dt = theta.dtype
    Vt = np.zeros(B, dtype=dt)
    Q = np.zeros((B, N + 2, M + 2, 3), dtype=dt)
    L = lib()
    for b in range(B):
        th, aa = _d(theta[b]), _d(A[b])
        q = np.empty((N + 2, M + 2, 3), dtype=np.float64)
        vt = L.oracle_forward(_p(th), _p(aa), N, M, _i0(mode), _p(q))
        Vt[b] = vt
        Q[b] = q
    return Vt, Q


def backward_pass(Et, Q, mode="nw"):
    """Et: [B], Q: [B,N+2,M+2,3] (stored dtype).  Returns E[B,N+2,M+2]."""
    Q = np.asarray(Q)
    Et = np.broadcast_to(np.asarray(Et), (Q.shape[0],))
    B, N2, M2, _ = Q.shape
    N, M = N2 - 2, M2 - 2
    E = np.zeros((B, N2, M2), dtype=Q.dtype)
    L = lib()
    for b in range(B):
        q = _d(Q[b])
        e = np.empty((N2, M2), dtype=np.float64)
        L.oracle_backward(float(Et[b]), _p(q), N, M, _i0(mode), _p(e))
        E[b] = e
    return E


def adjoint_forward_pass(Q, Ztheta, ZA):
    """Q [B,N+2,M+2,3], Ztheta [B,N+2,M+2] (padded), ZA [B,N,M].
    Returns (Vtd[B], Qd[B,N+2,M+2,3]) in Ztheta's dtype."""
    Q, Ztheta, ZA = np.asarray(Q), np.asarray(Ztheta), np.asarray(ZA)
    B, N2, M2 = Ztheta.shape
    N, M = N2 - 2, M2 - 2
    dt = Ztheta.dtype
    Vtd = np.zeros(B, dtype=dt)
    Qd = np.zeros((B, N2, M2, 3), dtype=dt)
    L = lib()
    for b in range(B):
        q, zt, za = _d(Q[b]), _d(Ztheta[b]), _d(ZA[b])
        qd = np.empty((N2, M2, 3), dtype=np.float64)
        Vtd[b] = L.oracle_adjoint_forward(_p(q), _p(zt), _p(za), N, M, _p(qd))
        Qd[b] = qd
    return Vtd, Qd


def adjoint_backward_pass(E, Q, Qd):
    """Returns Ed[B,N+2,M+2] in E's dtype."""
    E, Q, Qd = np.asarray(E), np.asarray(Q), np.asarray(Qd)
    B, N2, M2 = E.shape
    N, M = N2 - 2, M2 - 2
    Ed = np.zeros((B, N2, M2), dtype=E.dtype)
    L = lib()
    for b in range(B):
        e, q, qd = _d(E[b]), _d(Q[b]), _d(Qd[b])
        ed = np.empty((N2, M2), dtype=np.float64)
        L.oracle_adjoint_backward(_p(e), _p(q), _p(qd), N, M, _p(ed))
        Ed[b] = ed
    return Ed


def traceback(grad, variant="cpu"):
    """grad: [N,M] array.  Returns list[(i, j, state)] exactly like
    NeedlemanWunschDecoder.traceback of deepblast.nw ('cpu') or
    deepblast.nw_cuda ('cuda')."""
    g = np.ascontiguousarray(np.asarray(grad), dtype=np.float32)
    N, M = g.shape
    cap = 2 * (N + M) + 8
    out = np.empty((cap, 3), dtype=np.int32)
    n = lib().oracle_traceback(_p(g, _fp), N, M,
                               {"cpu": 0, "cuda": 1}[variant],
                               _p(out, _ip), cap)
    if n == -2:
        raise IndexError("traceback walked past the wrap-around range")
    if n < 0:
        raise RuntimeError("oracle_traceback: output capacity exceeded")
    return [tuple(int(v) for v in row) for row in out[:n]]


def decode(theta, A, mode="nw"):
    """dVt/dtheta [B,N,M] with Et = 1, i.e. decoder.decode (nw.py:446-458)."""
    Vt, Q = forward_pass(theta, A, mode)
    E = backward_pass(np.ones(len(Vt)), Q, mode)
    return Vt, Q, E


def fwd_bwd_batch_f32(theta, A, mode="nw", xlen=None, ylen=None,
                      nthreads=1, want_E=True):
    """Batched fp32 forward+backward used by bench.py's CPU legs."""
    theta = np.ascontiguousarray(theta, dtype=np.float32)
    A = np.ascontiguousarray(A, dtype=np.float32)
    B, N, M = theta.shape
    Vt = np.empty(B, dtype=np.float32)
    E = np.empty((B, N + 2, M + 2), dtype=np.float32) if want_E else None
    xl = None if xlen is None else np.ascontiguousarray(xlen, dtype=np.int32)
    yl = None if ylen is None else np.ascontiguousarray(ylen, dtype=np.int32)
    lib().oracle_fwd_bwd_batch_f32(
        _p(theta, _fp), _p(A, _fp), B, N, M, _i0(mode),
        None if xl is None else _p(xl, _ip),
        None if yl is None else _p(yl, _ip),
        _p(Vt, _fp), None if E is None else _p(E, _fp), int(nthreads))
    return Vt, E
