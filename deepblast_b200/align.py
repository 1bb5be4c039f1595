"""Batched inference on top of the decoder: the per-pair loop of the reference's
`NeuralAligner.traceback` (deepblast/alignment.py:139-171) and the state-string emission of
`DeepBLAST.align` (deepblast/trainer.py:80-88, `revstate_f` in dataset/utils.py:32-38) as
three launches for the whole batch.

The reference runs, for every pair b, a B = 1 `decode` on the slices
`match[b, :xlen[b], :ylen[b]]`, `gap[b, :xlen[b], :ylen[b]]` and then a Python walk with
three device->host synchronisations per step.  Here one forward and one backward launch
with per-pair lengths compute every pair on its own sub-lattice, and one traceback launch
walks all of them.
"""
import torch

from . import ops

# deepblast/constants.py: x, m, y = 0, 1, 2; dataset/utils.py:32-38 revstate_f
_STATE_CHARS = {0: '1', 1: ':', 2: '2'}


def state_string(decoded):
    """[(i, j, state), ...] -> the reference's alignment string ('1' gap in x, ':' match,
    '2' gap in y), i.e. ''.join(map(revstate_f, states)) of trainer.py:86-87."""
    return ''.join(_STATE_CHARS[s] for _, _, s in decoded)


def traceback_pairs(decoder, match, gap, xlen, ylen, variant="cuda"):
    """Generator with the contract of `NeuralAligner.traceback` (alignment.py:160-171):
    yields `(decoded, aln)` per pair, `decoded` the list of (i, j, state) tuples and `aln`
    the expected alignment matrix `[1, xlen[b], ylen[b]]` of that pair.

    decoder : deepblast_b200 NeedlemanWunschDecoder / SmithWatermanDecoder
    match, gap : CUDA fp32 [B, N, M] (theta and A of alignment.py:162-163)
    xlen, ylen : per-pair lengths (tensor, list or array of B ints)
    """
    B, N, M = match.shape
    xl = torch.as_tensor(xlen, dtype=torch.int32).reshape(B).cpu()
    yl = torch.as_tensor(ylen, dtype=torch.int32).reshape(B).cpu()
    if int(xl.min()) < 1 or int(yl.min()) < 1 or int(xl.max()) > N or int(yl.max()) > M:
        raise ValueError(f"lengths must satisfy 1 <= xlen <= {N}, 1 <= ylen <= {M}")
    with torch.enable_grad():
        th = match if match.requires_grad else match.detach().requires_grad_()
        a = gap if gap.requires_grad else gap.detach().requires_grad_()
        aln = decoder.decode(th, a, xl.to(match.device), yl.to(match.device))     # fwd + bwd, all pairs
    paths = ops.traceback_batch(aln.detach(), xl, yl, variant)                      # one launch + one D2H
    xl_h, yl_h = xl.tolist(), yl.tolist()
    for b in range(B):
        yield paths[b], aln[b:b + 1, :xl_h[b], :yl_h[b]]


def align_batch(decoder, match, gap, xlen, ylen, variant="cuda"):
    """All pairs at once: (state strings, decoded paths, aln [B, N, M]).  The strings are
    what `DeepBLAST.align` returns pair by pair (trainer.py:80-88)."""
    decoded, alns = [], []
    for d, a in traceback_pairs(decoder, match, gap, xlen, ylen, variant):
        decoded.append(d)
        alns.append(a)
    return [state_string(d) for d in decoded], decoded, alns


class HostAligner:
    """Inference for a caller whose theta / A live in pinned HOST memory (a data loader, a CPU
    embedding stage): what `DeepBLAST.align` needs per pair (deepblast/trainer.py:80-88 ->
    alignment.py:160-171) -- the alignment PATH -- for a whole batch.  Chunks of pairs flow
    upload | forward + backward + on-device traceback | download on three streams, and only the
    paths (int32 triples (i, j, state), at most n + m per pair) and the scores travel back: a few
    KB per pair instead of the (N+2) x (M+2) expected-alignment matrix `decode_host` returns.
    Equal-size batches run the whole pipeline inside the library (C ABI b200dp_align_host); with
    per-pair lengths the chunks go through the strip-queue kernels from here.

        al = HostAligner(B, N, M, mode="nw")
        paths_h, len_h, Vt_h = al.align(theta_h, A_h)      # enqueue; valid after a stream sync
        al.paths(b)                                        # [(i, j, state), ...] of pair b
    """

    def __init__(self, B, N, M, mode="nw", xlen=None, ylen=None, device=None, chunk_pairs=None, variant="cuda"):
        from . import plan as _plan, _lib
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.B, self.N, self.M, self.mode = int(B), int(N), int(M), mode
        self.variant = {"cpu": 0, "cuda": 1}[variant]
        if M % 4 != 0:
            raise RuntimeError("HostAligner needs M % 4 == 0 (16-byte rows)")
        if chunk_pairs is None:
            chunk_pairs = max(1, min(B, max(8, (6 << 20) // max(1, N * M)), (B + 9) // 10 if B >= 20 else B))
        self.cap = _lib.lib().b200dp_align_host_path_cap(N, M)
        self.chunk_pairs = int(chunk_pairs)
        self.native = xlen is None and ylen is None
        self._lib = _lib
        self.paths_h = torch.empty((B, self.cap, 3), dtype=torch.int32, pin_memory=True)
        self.len_h = torch.empty(B, dtype=torch.int32, pin_memory=True)
        self.Vt_h = torch.empty(B, dtype=torch.float32, pin_memory=True)
        if self.native:
            with torch.cuda.device(self.device):
                need = _lib.lib().b200dp_align_host_workspace(N, M, self.chunk_pairs)
                self.ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self.chunks = [(b0, min(B, b0 + self.chunk_pairs), None) for b0 in range(0, B, self.chunk_pairs)]
            return
        xl = None if xlen is None else torch.as_tensor(xlen, dtype=torch.int32).reshape(B).cpu()
        yl = None if ylen is None else torch.as_tensor(ylen, dtype=torch.int32).reshape(B).cpu()
        self.chunks = []
        for b0 in range(0, B, chunk_pairs):
            b1 = min(B, b0 + chunk_pairs)
            cx = None if xl is None else xl[b0:b1]
            cy = None if yl is None else yl[b0:b1]
            pl = _plan.get_plan(b1 - b0, N, M, cx, cy, False, self.device)
            self.chunks.append((b0, b1, pl))
        cb = min(chunk_pairs, B)
        with torch.cuda.device(self.device):
            self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(device=self.device) for _ in range(3))
            mk = lambda *s, dt=torch.float32: torch.empty(s, dtype=dt, device=self.device)      # noqa: E731
            qcap = max(int(pl.q_floats) for _, _, pl in self.chunks)
            # (results live in per-slot buffers too: nothing is allocated inside align())
            self.slots = [(mk(cb, N, M), mk(cb, N, M), mk(cb, self.cap, 3, dt=torch.int32), mk(cb, dt=torch.int32),
                           mk(qcap), mk(cb, N, M), mk(cb)) for _ in range(3)]
            self.cmp_done = [torch.cuda.Event() for _ in range(3)]
            self.out_done = [torch.cuda.Event() for _ in range(3)]
            self.ones = torch.ones(B, dtype=torch.float32, device=self.device)
            self.xl_d = None if xl is None else xl.to(self.device)
            self.yl_d = None if yl is None else yl.to(self.device)

    def align(self, theta_h, A_h):
        """Enqueue one batch; returns (paths_h [B, cap, 3] int32, len_h [B] int32, Vt_h [B]) pinned
        buffers, valid after the device's current stream has been synchronised."""
        for name, t in (("theta_h", theta_h), ("A_h", A_h)):
            if t.is_cuda or t.dtype != torch.float32 or tuple(t.shape) != (self.B, self.N, self.M) or not t.is_contiguous():
                raise RuntimeError(f"{name} must be a contiguous float32 host tensor [{self.B}, {self.N}, {self.M}]")
        dev = self.device
        L = self._lib.lib()
        if self.native:
            from .ops import MODES
            with torch.cuda.device(dev):
                rc = L.b200dp_align_host(theta_h.data_ptr(), A_h.data_ptr(), self.Vt_h.data_ptr(), self.paths_h.data_ptr(),
                                         self.len_h.data_ptr(), None, self.B, self.N, self.M, MODES[self.mode], self.variant,
                                         self.chunk_pairs, self.ws.data_ptr(), self.ws.numel(), 0,
                                         torch.cuda.current_stream(dev).cuda_stream)
                self._lib.check(rc, "b200dp_align_host")
            return self.paths_h, self.len_h, self.Vt_h
        cur = torch.cuda.current_stream(dev)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        for c, (b0, b1, pl) in enumerate(self.chunks):
            th_d, a_d, out_d, ln_d, q_d, e_d, vt_d = self.slots[c % 3]
            n = b1 - b0
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(self.cmp_done[c % 3])            # the slot's previous sweeps are done
                th_d[:n].copy_(theta_h[b0:b1], non_blocking=True)
                a_d[:n].copy_(A_h[b0:b1], non_blocking=True)
                up = torch.cuda.Event()
                up.record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(up)
                self.s_cmp.wait_event(self.out_done[c % 3])           # the slot's previous paths have left
                Vt, Q = ops.sq_forward(pl, th_d[:n], a_d[:n], self.mode, out=(vt_d, q_d))
                E = ops.sq_backward(pl, self.ones[b0:b1], Q, self.mode, out=e_d[:n])          # [n, N, M] interior
                rc = L.b200dp_traceback(E.data_ptr(), E.stride(0), E.stride(1), E.stride(2),
                                        None if self.xl_d is None else self.xl_d[b0:b1].data_ptr(),
                                        None if self.yl_d is None else self.yl_d[b0:b1].data_ptr(),
                                        n, self.N, self.M, self.variant, out_d.data_ptr(), self.cap, ln_d.data_ptr(),
                                        self.s_cmp.cuda_stream)
                self._lib.check(rc, "b200dp_traceback")
                self.cmp_done[c % 3].record(self.s_cmp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.cmp_done[c % 3])
                self.paths_h[b0:b1].copy_(out_d[:n], non_blocking=True)
                self.len_h[b0:b1].copy_(ln_d[:n], non_blocking=True)
                self.Vt_h[b0:b1].copy_(Vt[:n], non_blocking=True)
                self.out_done[c % 3].record(self.s_out)
        cur.wait_stream(self.s_out)
        return self.paths_h, self.len_h, self.Vt_h

    def paths(self, b):
        """[(i, j, state), ...] of pair b (after a synchronisation)."""
        n = int(self.len_h[b])
        if n == -2:
            raise IndexError("index out of range in traceback (negative wrap-around exhausted)")
        if n < 0:
            raise RuntimeError("b200dp_traceback: output capacity exceeded")
        return [tuple(int(v) for v in row) for row in self.paths_h[b, :n].numpy()]

    def state_strings(self):
        """The reference's alignment strings of all pairs (trainer.py:86-87), after a synchronisation."""
        import numpy as np
        chars = np.array(['1', ':', '2'])
        st = self.paths_h[:, :, 2].numpy()
        return [''.join(chars[st[b, :int(self.len_h[b])]]) for b in range(self.B)]
