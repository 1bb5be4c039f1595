"""Work-queue plans for the strip-queue kernels (csrc/softdp_sq.cuh) and the PACKED pair layout.

A `Plan` describes one batch of pairs -- B lattices n_b x m_b, equal or ragged -- to the
engine: the two strip tables (forward and backward ticket order) that `b200dp_plan_build`
(host code of libb200dp.so) writes, uploaded once and reused by all four sweeps, plus the
offsets of every pair in the operand and Q buffers.

Layouts (include/b200dp.h):
  dense   theta / A / E are [B, N, M] tensors (M % 4 == 0); pair b lives in the top-left
          n_b x m_b corner -- what `NeuralAligner.forward` builds by padding every sequence to
          the global maximum (deepblast/dataset/utils.py:245, alignment.py:117-124);
  packed  one flat fp32 buffer, pair b an n_b x pitch_b row-major block at `pair_off[b]`,
          pitch_b = roundup(m_b, 4): nothing is padded to the longest pair, so H2D copies,
          HBM traffic and the memory of Q / E all scale with the useful cells
          (`pack_sequences` / `unpack_sequences`,
          dataset/utils.py:214-251, carried through to the DP operands).
"""
import ctypes
import threading
from collections import OrderedDict

import numpy as np
import torch

from . import _lib


class _PlanInfo(ctypes.Structure):
    _fields_ = [("nstrips", ctypes.c_int), ("max_m", ctypes.c_int), ("q_floats", ctypes.c_longlong),
                ("bnd_words", ctypes.c_longlong), ("packed_floats", ctypes.c_longlong),
                ("cells", ctypes.c_longlong), ("grid_fwd", ctypes.c_int), ("grid_bwd", ctypes.c_int)]


def _lens_host(x, B, name):
    if x is None:
        return None
    if torch.is_tensor(x):
        x = x.detach().cpu().numpy()          # a CUDA tensor costs one synchronising copy here
    a = np.ascontiguousarray(np.asarray(x).reshape(-1), dtype=np.int32)
    if a.shape != (B,):
        raise RuntimeError(f"{name} must have shape [B] = [{B}]")
    return a


def _ip(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def _lp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong))


class Plan:
    """Strip tables + offsets of one batch.  Build once per batch (`get_plan` caches by
    lengths), pass to ops.sq_* / the autograd functions."""

    def __init__(self, B, N, M, xlen=None, ylen=None, packed=False, device=None, resident_warps=None):
        self.B, self.N, self.M, self.packed = int(B), int(N), int(M), bool(packed)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        xl, yl = _lens_host(xlen, B, "xlen"), _lens_host(ylen, B, "ylen")
        self.ragged = xl is not None or yl is not None
        L = _lib.lib()
        if resident_warps is None:
            resident_warps = (0, 0)
            if self.device.type == "cuda":
                with torch.cuda.device(self.device):
                    # the ticket order is a list schedule for the warps each sweep keeps resident
                    resident_warps = (L.b200dp_sq_resident_warps(0), L.b200dp_sq_resident_warps(1))
        wf, wb = (resident_warps, resident_warps) if isinstance(resident_warps, int) else resident_warps
        info = _PlanInfo()
        self.pair_off = np.zeros(max(B, 1), dtype=np.int64)
        self.q_off = np.zeros(max(B, 1), dtype=np.int64)
        rc = L.b200dp_plan_build(_ip(xl), _ip(yl), B, N, M, int(packed), int(wf), int(wb), ctypes.byref(info),
                                 _lp(self.pair_off), _lp(self.q_off), None, None, 0)
        _lib.check(rc, "b200dp_plan_build")
        self.nstrips = info.nstrips
        tabs = np.zeros((2, max(1, info.nstrips), 64), dtype=np.uint8)
        rc = L.b200dp_plan_build(_ip(xl), _ip(yl), B, N, M, int(packed), int(wf), int(wb), ctypes.byref(info),
                                 _lp(self.pair_off), _lp(self.q_off), tabs[0].ctypes.data, tabs[1].ctypes.data,
                                 info.nstrips)
        _lib.check(rc, "b200dp_plan_build")
        self.q_floats, self.bnd_words = info.q_floats, info.bnd_words
        self.packed_floats, self.cells, self.max_m = info.packed_floats, info.cells, info.max_m
        self.ws_bytes = L.b200dp_sq_workspace_bytes(info.bnd_words)
        self.grid_fwd, self.grid_bwd = info.grid_fwd, info.grid_bwd
        self.xlen = np.full(B, N, np.int32) if xl is None else np.clip(xl, 0, N)
        self.ylen = np.full(B, M, np.int32) if yl is None else np.clip(yl, 0, M)
        empty = (self.xlen == 0) | (self.ylen == 0)
        self.xlen[empty] = 0
        self.ylen[empty] = 0
        self.has_empty = bool(empty.any())
        self.pitch = ((self.ylen + 3) & ~3) if packed else np.full(B, M, np.int32)
        self.tabs_host = tabs
        self._tabs_dev = None
        if self.device.type == "cuda":
            self._tabs_dev = torch.from_numpy(tabs).to(self.device, non_blocking=False)

    @property
    def fwd_tab(self):
        return self._tabs_dev[0]

    @property
    def bwd_tab(self):
        return self._tabs_dev[1]

    # ---- packed layout helpers (pure indexing: they also run on CPU tensors) ----------------
    def pair_view(self, flat, b):
        """The n_b x m_b matrix of pair b inside a flat operand buffer of this plan (a strided view)."""
        n, m, pitch, off = int(self.xlen[b]), int(self.ylen[b]), int(self.pitch[b]), int(self.pair_off[b])
        if not self.packed:
            return flat.reshape(self.B, self.N, self.M)[b, :n, :m]
        return flat.as_strided((n, m), (pitch, 1), flat.storage_offset() + off)

    def pack(self, dense):
        """[B, N, M] tensor (or a list of [n_b, m_b] matrices) -> flat packed buffer."""
        if not self.packed:
            raise RuntimeError("pack() needs a packed plan")
        ref = dense[0] if isinstance(dense, (list, tuple)) else dense
        flat = torch.zeros(self.packed_floats, dtype=torch.float32, device=ref.device)
        for b in range(self.B):
            n, m = int(self.xlen[b]), int(self.ylen[b])
            if n and m:
                self.pair_view(flat, b).copy_(dense[b][:n, :m])
        return flat

    def unpack(self, flat, fill=0.0):
        """flat packed buffer -> dense [B, N, M] tensor (`fill` outside each pair's corner)."""
        out = torch.full((self.B, self.N, self.M), fill, dtype=flat.dtype, device=flat.device)
        for b in range(self.B):
            n, m = int(self.xlen[b]), int(self.ylen[b])
            if n and m:
                out[b, :n, :m] = self.pair_view(flat, b)
        return out


_plans = OrderedDict()
_plans_lock = threading.Lock()
_PLAN_CACHE = 32


def get_plan(B, N, M, xlen=None, ylen=None, packed=False, device=None):
    """Cached `Plan` (keyed by shape, lengths, layout and device)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    xl, yl = _lens_host(xlen, B, "xlen"), _lens_host(ylen, B, "ylen")
    key = (dev.type, dev.index, B, N, M, bool(packed), None if xl is None else xl.tobytes(),
           None if yl is None else yl.tobytes())
    with _plans_lock:
        p = _plans.get(key)
        if p is not None:
            _plans.move_to_end(key)
            return p
    p = Plan(B, N, M, xl, yl, packed, dev)
    with _plans_lock:
        _plans[key] = p
        while len(_plans) > _PLAN_CACHE:
            _plans.popitem(last=False)
    return p


def packed_plan(xlen, ylen, device=None):
    """Plan of a packed batch from its lengths alone."""
    xl = _lens_host(xlen, len(xlen), "xlen")
    yl = _lens_host(ylen, len(ylen), "ylen")
    return get_plan(len(xl), max(1, int(xl.max(initial=1))), max(1, int(yl.max(initial=1))), xl, yl, True, device)


# ---- per-(device, stream) workspace: zeroed once, self-cleaning, one launch at a time ---------
_ws = {}
_ws_lock = threading.Lock()


def workspace(device, stream_ptr, nbytes):
    """The workspace tensor for strip-queue launches on this stream (grown when a plan needs more;
    a fresh one is zero-filled on the current stream, which is all the kernels require)."""
    key = (device.index, int(stream_ptr))
    with _ws_lock:
        ws = _ws.get(key)
        if ws is None or ws.numel() < nbytes:
            size = max(int(nbytes), 1 << 20)
            if ws is not None:
                size = size * 5 // 4
            ws = _ws[key] = torch.zeros((size + 255) & ~255, dtype=torch.uint8, device=device)
        return ws
