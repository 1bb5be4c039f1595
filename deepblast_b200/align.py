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
