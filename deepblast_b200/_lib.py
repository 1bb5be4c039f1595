"""ctypes binding of libb200dp.so (include/b200dp.h).  Fails loudly when the
library is missing: there is no CPU or PyTorch fallback on this path."""
import ctypes
import functools
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200dp.so")

_lib = None

_f = ctypes.c_void_p      # device pointers travel as integers
_i = ctypes.c_int
_ll = ctypes.c_longlong

EXPORTS = {
    "b200dp_version": (ctypes.c_int, []),
    "b200dp_last_error": (ctypes.c_char_p, []),
    "b200dp_q_layout": (ctypes.c_int, [_i, _i, ctypes.POINTER(_i), ctypes.POINTER(_ll),
                                       ctypes.POINTER(_ll), ctypes.POINTER(_ll)]),
    "b200dp_fwd": (ctypes.c_int, [_f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _f]),
    "b200dp_bwd": (ctypes.c_int, [_f, _ll, _f, _f, _f, _f, _i, _i, _i, _i, _i, _f]),
    "b200dp_bwd_keep_interior": (ctypes.c_int, [_f, _ll, _f, _f, _f, ctypes.POINTER(_i), _i, _i, _i, _i, _i, _f]),
    "b200dp_adj_fwd": (ctypes.c_int, [_f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _f]),
    "b200dp_adj_bwd": (ctypes.c_int, [_f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _f]),
    "b200dp_decode_host_workspace": (ctypes.c_size_t, [_i, _i, _i]),
    "b200dp_decode_host": (ctypes.c_int, [_f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _f, ctypes.c_size_t, _i, _f]),
    "b200dp_align_host_workspace": (ctypes.c_size_t, [_i, _i, _i]),
    "b200dp_align_host_path_cap": (ctypes.c_int, [_i, _i]),
    "b200dp_align_host": (ctypes.c_int, [_f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _f, ctypes.c_size_t, _i, _f]),
    "b200dp_adj3_applicable": (ctypes.c_int, [_i, _i, _i]),
    "b200dp_adj_fwd3": (ctypes.c_int, [_f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _f]),
    "b200dp_adj_bwd3": (ctypes.c_int, [_f, _f, _f, _f, _i, _i, _i, _i, _f]),
    "b200dp_mxent_fwd": (ctypes.c_int, [_f, _f, _ll, _ll, _f, _f, _f, _i, _i, _i, _f, _f, _f]),
    "b200dp_mxent_bwd": (ctypes.c_int, [_f, _f, _ll, _ll, _f, _f, _f, _i, _i, _i, _f, _f, _f, _f]),
    "b200dp_traceback": (ctypes.c_int, [_f, _ll, _ll, _ll, _f, _f, _i, _i, _i, _i, _f, _i, _f, _f]),
    # strip-queue family (plan builder is host code: works without a GPU)
    "b200dp_plan_build": (ctypes.c_int, [_f, _f, _i, _i, _i, _i, _i, _i, _f, _f, _f, _f, _f, _i]),
    "b200dp_sq_workspace_bytes": (ctypes.c_size_t, [_ll]),
    "b200dp_sq_resident_warps": (ctypes.c_int, [_i]),
    "b200dp_sq_set_trace": (None, [_f]),
    "b200dp_sq_fwd": (ctypes.c_int, [_f, _i, _f, _f, _f, _f, _f, _i, _i, _f]),
    "b200dp_sq_fwd_dense": (ctypes.c_int, [_f, _i, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _f]),
    "b200dp_sq_bwd": (ctypes.c_int, [_f, _i, _f, _f, _ll, _f, _f, _i, _i, _f]),
    "b200dp_sq_adj_fwd": (ctypes.c_int, [_f, _i, _f, _f, _f, _f, _f, _f, _f, _i, _f]),
    "b200dp_sq_adj_bwd": (ctypes.c_int, [_f, _i, _f, _f, _f, _f, _i, _f]),
    # cluster kernels (small batches of long pairs)
    "b200dp_cl_applicable": (ctypes.c_int, [_i, _i, _i]),
    "b200dp_cl_fwd": (ctypes.c_int, [_f, _f, _f, _f, _i, _i, _i, _i, _i, _f]),
    "b200dp_cl_bwd": (ctypes.c_int, [_f, _ll, _f, _f, _i, _i, _i, _i, _i, _f]),
    # theta / A producer (tcgen05 GEMM)
    "b200dp_theta_a_workspace": (ctypes.c_size_t, [_i, _i, _i, _i]),
    "b200dp_theta_a": (ctypes.c_int, [_f, _f, _f, _f, _i, _i, _i, _i, _f, _f, _f, _f, _f, _f, ctypes.c_size_t, _f]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m deepblast_b200.build` "
                "(nvcc, sm_100a).  deepblast_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().b200dp_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed ({rc}): {msg}")


@functools.lru_cache(maxsize=256)
def q_layout(N, M):
    """(K strips per pair, strip_stride, pair_stride, tail pad) in floats (a pure function
    of the lattice size, asked several times per call: cached)."""
    K = _i()
    ss, ps, pad = _ll(), _ll(), _ll()
    check(lib().b200dp_q_layout(N, M, ctypes.byref(K), ctypes.byref(ss), ctypes.byref(ps),
                                ctypes.byref(pad)), "b200dp_q_layout")
    return K.value, ss.value, ps.value, pad.value
