"""Ragged batches end to end in the PACKED layout (plan.py): what `NeuralAligner.traceback`
(deepblast/alignment.py:160-171) does pair by pair -- decode on the slice
`match[b, :xlen[b], :ylen[b]]` -- for a whole batch whose operands never get padded to the
longest pair (`pack_sequences` / `unpack_sequences`,
deepblast/dataset/utils.py:214-251, carried through to theta / A / E).

    plan  = packed_plan(xlen, ylen)                  # offsets + work-queue tables, built on the host
    theta = plan.pack(theta_dense)                   # or write pair b at plan.pair_view(flat, b)
    aln   = decoder.decode(theta, A, plan=plan)      # flat packed dVt/dtheta, differentiable again
    paths = traceback_packed(plan, aln)              # one launch per pitch group

`PackedHostDecoder` is the host-buffer form (the caller's theta / A live in pinned HOST memory):
only the useful cells cross PCIe, and upload, sweeps and download of consecutive chunks of pairs
overlap on three streams.
"""
import numpy as np
import torch

from . import ops
from .plan import get_plan, packed_plan      # noqa: F401  (re-exported)


class PackedHostDecoder:
    """decode (forward + backward, Et = 1) of a ragged batch held in pinned host memory in the
    packed layout: (Vt_h [B], E_h packed) <- (theta_h, A_h packed)."""

    def __init__(self, xlen, ylen, mode="nw", device=None, nchunks=12):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.mode = mode
        xl = np.ascontiguousarray(np.asarray(xlen).reshape(-1), dtype=np.int32)
        yl = np.ascontiguousarray(np.asarray(ylen).reshape(-1), dtype=np.int32)
        self.B = len(xl)
        N, M = max(1, int(xl.max(initial=1))), max(1, int(yl.max(initial=1)))
        self.plan = get_plan(self.B, N, M, xl, yl, True, self.device)        # the whole batch: offsets
        self.packed_floats = int(self.plan.packed_floats)
        # chunks of consecutive pairs with about equal packed size; a chunk's sub-plan has the
        # same relative offsets (every pair starts on a multiple of 32 floats)
        off = np.append(self.plan.pair_off[:self.B], self.packed_floats)
        nchunks = max(1, min(nchunks, self.B))
        target = np.linspace(0, self.packed_floats, nchunks + 1)[1:-1]
        cuts = sorted(set([0] + [int(np.searchsorted(off, t)) for t in target] + [self.B]))
        self.chunks = []
        for b0, b1 in zip(cuts[:-1], cuts[1:]):
            if b1 > b0:
                sub = get_plan(b1 - b0, N, M, xl[b0:b1], yl[b0:b1], True, self.device)
                self.chunks.append((b0, b1, int(off[b0]), int(off[b1]), sub))
        self.nchunks = len(self.chunks)
        cap = max(o1 - o0 for _, _, o0, o1, _ in self.chunks)
        with torch.cuda.device(self.device):
            self.s_in = torch.cuda.Stream(device=self.device)
            self.s_cmp = torch.cuda.Stream(device=self.device)
            self.s_out = torch.cuda.Stream(device=self.device)
            # per slot: theta, A, E (packed floats of the largest chunk), Q (its strip-major floats), Vt -- nothing
            # is allocated inside decode(): an allocation there can fall through to cudaMalloc (a device-wide
            # synchronisation in the middle of the pipeline) whenever the caching allocator has no block ready
            qcap = max(int(sub.q_floats) for _, _, _, _, sub in self.chunks)
            bcap = max(b1 - b0 for b0, b1, _, _, _ in self.chunks)
            mk = lambda n: torch.empty(n, dtype=torch.float32, device=self.device)      # noqa: E731
            self.slots = [(mk(cap), mk(cap), mk(cap), mk(qcap), mk(bcap)) for _ in range(3)]
            self.out_done = [torch.cuda.Event() for _ in range(3)]
            self.cmp_done = [torch.cuda.Event() for _ in range(3)]
            self.ones = torch.ones(self.B, dtype=torch.float32, device=self.device)
        self.Vt_h = torch.empty(self.B, dtype=torch.float32, pin_memory=True)
        self.E_h = torch.empty(self.packed_floats, dtype=torch.float32, pin_memory=True)

    def decode(self, theta_h, A_h, out=None):
        """Enqueue one decode; returns (Vt_h, E_h) pinned buffers, valid after the current stream
        of the device has been synchronised."""
        Vt_h, E_h = (self.Vt_h, self.E_h) if out is None else out
        dev = self.device
        cur = torch.cuda.current_stream(dev)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        for c, (b0, b1, o0, o1, sub) in enumerate(self.chunks):
            th_d, a_d, e_d, q_d, vt_d = self.slots[c % 3]
            n = o1 - o0
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(self.cmp_done[c % 3])            # the slot's previous sweeps are done
                th_d[:n].copy_(theta_h[o0:o1], non_blocking=True)
                a_d[:n].copy_(A_h[o0:o1], non_blocking=True)
                up = torch.cuda.Event()
                up.record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(up)
                self.s_cmp.wait_event(self.out_done[c % 3])            # the slot's previous results have left
                Vt, Q = ops.sq_forward(sub, th_d[:n], a_d[:n], self.mode, out=(vt_d, q_d))
                E = ops.sq_backward(sub, self.ones[b0:b1], Q, self.mode, out=e_d)
                self.cmp_done[c % 3].record(self.s_cmp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.cmp_done[c % 3])
                E_h[o0:o1].copy_(E[:n], non_blocking=True)
                Vt_h[b0:b1].copy_(Vt[:b1 - b0], non_blocking=True)
                self.out_done[c % 3].record(self.s_out)
        cur.wait_stream(self.s_out)
        return Vt_h, E_h


def traceback_packed(plan, aln, variant="cuda"):
    """All pairs of a packed expected-alignment buffer -> list of B paths [(i, j, state), ...]
    (the per-pair walk of deepblast/nw_cuda.py:273-317).  Pairs that share a row pitch go to the
    device together (one launch per pitch)."""
    if not plan.packed:
        return ops.traceback_batch(aln, torch.as_tensor(plan.xlen), torch.as_tensor(plan.ylen), variant)
    out = [None] * plan.B
    by_pitch = {}
    for b in range(plan.B):
        if plan.xlen[b] and plan.ylen[b]:
            by_pitch.setdefault(int(plan.pitch[b]), []).append(b)
        else:
            out[b] = []
    flat = aln.detach()
    for pitch, bs in by_pitch.items():
        # a [len(bs), Nmax, pitch] view cannot express arbitrary offsets: walk them as a batch of
        # strided [1, n, m] views gathered into one dense tensor (the walk reads O(n + m) cells, the
        # copy is what it costs to keep one launch per pitch)
        nmax = max(int(plan.xlen[b]) for b in bs)
        dense = torch.zeros((len(bs), nmax, pitch), dtype=torch.float32, device=flat.device)
        for i, b in enumerate(bs):
            n, m = int(plan.xlen[b]), int(plan.ylen[b])
            dense[i, :n, :m] = plan.pair_view(flat, b)
        xl = torch.tensor([int(plan.xlen[b]) for b in bs], dtype=torch.int32)
        yl = torch.tensor([int(plan.ylen[b]) for b in bs], dtype=torch.int32)
        paths = ops.traceback_batch(dense, xl, yl, variant)
        for i, b in enumerate(bs):
            out[b] = paths[i]
    return out
