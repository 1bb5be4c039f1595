"""GPU tests of the drop-in autograd surface (deepblast_b200.nw_cuda / sw_cuda),
modelled on the reference's own GPU tests (deepblast/tests/test_nw_cuda.py:25-87,
test_sw_cuda.py) plus direct comparison with reference-generated goldens."""
import numpy as np
import pytest
import torch
from torch.autograd import gradcheck
from torch.autograd.gradcheck import gradgradcheck

from oracle import softdp as O

pytestmark = pytest.mark.gpu

CASES = ["t_nw_cuda_5x5", "r8x8", "r17x23", "r33x40", "r64x48"]


def dev():
    return torch.device("cuda:0")


def decoders():
    from deepblast_b200.nw_cuda import NeedlemanWunschDecoder
    from deepblast_b200.sw_cuda import SmithWatermanDecoder
    return {"nw": NeedlemanWunschDecoder, "sw": SmithWatermanDecoder}


def setup_small():
    torch.manual_seed(2)
    B, N, M = 3, 5, 5
    theta = torch.rand(B, N, M, requires_grad=True, dtype=torch.float32, device=dev())
    A = -1. * torch.ones_like(theta)
    return theta, A


@pytest.mark.parametrize("mode", ["nw", "sw"])
def test_gradcheck_like_reference(mode):
    # test_nw_cuda.py:51-55, test_sw_cuda.py:50-55
    needle = decoders()[mode]('softmax')
    theta, A = setup_small()
    gradcheck(needle, (theta, A), eps=1e-1, atol=1e-1, rtol=1e-1)


def test_gradgradcheck_like_reference():
    # test_nw_cuda.py:57-61
    needle = decoders()["nw"]('softmax')
    theta, A = setup_small()
    gradgradcheck(needle, (theta, A), eps=1e-1, atol=1e-1, rtol=1e-1)


@pytest.mark.parametrize("mode", ["nw", "sw"])
def test_decoding_known_answer(golden, golden_meta, mode):
    # test_nw_cuda.py:64-76 / test_sw_cuda.py:58-70
    theta = torch.tensor(golden["ref_make_data/theta"].astype(np.float32), device=dev())
    theta.requires_grad_()
    A = 0.1 * torch.ones_like(theta)
    needle = decoders()[mode]('softmax')
    v = needle(theta, A)
    v.backward()
    np.testing.assert_allclose(v.detach().cpu().numpy(), golden[f"ref_make_data/{mode}/Vt_f32"], rtol=1e-6)
    np.testing.assert_allclose(theta.grad.cpu().numpy(), golden[f"ref_make_data/{mode}/grad_f32"], atol=1e-5)
    decoded = needle.traceback(theta.grad.squeeze())
    want_xy = golden_meta["known_answers"]["nw_cuda_xy" if mode == "nw" else "sw_cuda_xy"]
    assert [list(x[:2]) for x in decoded] == want_xy
    assert decoded == [tuple(r) for r in golden[f"ref_make_data/{mode}/tb_cuda_f32"].tolist()]
    assert all(isinstance(v, int) for tup in decoded for v in tup)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", ["nw", "sw"])
def test_autograd_matches_reference(golden, name, mode):
    """decode + double backward through our Functions == the reference's (nw.py:315-386)."""
    g = lambda k: golden[f"{name}/{k}"]
    theta = torch.tensor(g("theta"), device=dev(), requires_grad=True)
    A = torch.tensor(g("A"), device=dev(), requires_grad=True)
    W = torch.tensor(g("W"), device=dev())
    dec = decoders()[mode]('softmax')
    aln = dec.decode(theta, A)
    assert aln.shape == theta.shape and aln.requires_grad
    np.testing.assert_allclose(aln.detach().cpu().numpy(), g(f"{mode}/ag_aln"), atol=2e-5)
    (aln * W).sum().backward()
    scale = max(1.0, float(np.abs(g(f"{mode}/ag_theta_grad2")).max()))
    np.testing.assert_allclose(theta.grad.cpu().numpy(), g(f"{mode}/ag_theta_grad2"), atol=1e-4 * scale)
    assert A.grad is None and bool(g(f"{mode}/ag_A_grad_is_none"))      # nw_cuda.py:262
    v = dec(theta, A)
    g_theta, g_A = torch.autograd.grad(v.sum(), (theta, A))
    np.testing.assert_allclose(g_theta.cpu().numpy(), g(f"{mode}/ag_g_theta"), atol=2e-5)
    np.testing.assert_array_equal(g_A.cpu().numpy(), g(f"{mode}/ag_g_A"))   # grad wrt A is A (nw_cuda.py:206-207)


def test_api_quirks():
    from deepblast_b200.nw_cuda import NeedlemanWunschDecoder, NeedlemanWunschFunction
    theta, A = setup_small()
    dec = NeedlemanWunschDecoder('softmax')
    with pytest.raises(NotImplementedError):
        NeedlemanWunschFunction.apply(theta, A, 'sparsemax')
    with pytest.raises(TypeError):
        NeedlemanWunschFunction.apply(theta.double(), A.double(), 'softmax')
    with pytest.raises(RuntimeError):           # decode needs both to require grad (nw.py:455-457)
        dec.decode(theta, A)
    A = A.clone().requires_grad_()
    full = dec.decode(theta, A)
    # non-contiguous slices as NeuralAligner.traceback passes them (alignment.py:166-169)
    th3, A3 = theta[1, :4, :3].unsqueeze(0), A[1, :4, :3].unsqueeze(0)
    assert not th3.is_contiguous()
    sub = dec.decode(th3, A3)
    ref = dec.decode(th3.contiguous().detach().requires_grad_(), A3.contiguous().detach().requires_grad_())
    assert torch.equal(sub, ref)
    assert full.device == theta.device and full.dtype == torch.float32
    assert not torch.is_anomaly_enabled()       # nw_cuda.py:9 is deliberately not replicated


def test_no_cpu_fallback():
    from deepblast_b200.nw_cuda import NeedlemanWunschFunction
    theta = torch.rand(1, 4, 4)
    with pytest.raises(RuntimeError):
        NeedlemanWunschFunction.apply(theta, -torch.ones_like(theta), 'softmax')


def test_varlen_decoder_matches_per_pair():
    from deepblast_b200.nw_cuda import NeedlemanWunschDecoder
    g = torch.Generator().manual_seed(0)
    B, N, M = 5, 48, 70
    theta = torch.rand(B, N, M, generator=g).to(dev()).requires_grad_()
    A = (-torch.rand(B, N, M, generator=g)).to(dev()).requires_grad_()
    xlen = torch.tensor([48, 20, 33, 1, 47], dtype=torch.int32, device=dev())
    ylen = torch.tensor([70, 64, 5, 9, 31], dtype=torch.int32, device=dev())
    dec = NeedlemanWunschDecoder('softmax')
    aln = dec.decode(theta, A, xlen, ylen)
    Wt = torch.randn(B, N, M, generator=g).to(dev())
    (aln * Wt).sum().backward()
    g_all = theta.grad.clone()
    for b in range(B):
        n, m = int(xlen[b]), int(ylen[b])
        th = theta[b:b + 1, :n, :m].detach().contiguous().requires_grad_()
        a = A[b:b + 1, :n, :m].detach().contiguous().requires_grad_()
        one = dec.decode(th, a)
        assert torch.allclose(aln[b, :n, :m], one[0], atol=1e-6)
        assert float(aln[b, n:, :].abs().sum()) == 0.0 and float(aln[b, :, m:].abs().sum()) == 0.0
        (one * Wt[b:b + 1, :n, :m]).sum().backward()
        assert torch.allclose(g_all[b, :n, :m], th.grad[0], atol=1e-4, rtol=1e-4)
        assert float(g_all[b, n:, :].abs().sum()) == 0.0


def test_install_patches_reference_names():
    import sys
    import types
    import deepblast_b200
    # a stand-in for an installed reference package (the real one is absent on the GPU box)
    pkg = types.ModuleType("deepblast")
    pkg.__path__ = []
    for name in ("nw_cuda", "sw_cuda", "alignment"):
        mod = types.ModuleType(f"deepblast.{name}")
        sys.modules[f"deepblast.{name}"] = mod
        setattr(pkg, name, mod)
    sys.modules["deepblast"] = pkg
    try:
        patched = deepblast_b200.install()
        from deepblast_b200.nw_cuda import NeedlemanWunschDecoder
        from deepblast_b200.sw_cuda import SmithWatermanDecoder
        assert sys.modules["deepblast.nw_cuda"].NeedlemanWunschDecoder is NeedlemanWunschDecoder
        assert sys.modules["deepblast.alignment"].SWDecoderCUDA is SmithWatermanDecoder
        assert "deepblast.nw_cuda.NeedlemanWunschFunction" in patched
    finally:
        for name in ("deepblast", "deepblast.nw_cuda", "deepblast.sw_cuda", "deepblast.alignment"):
            sys.modules.pop(name, None)


@pytest.mark.parametrize("mode", ["nw", "sw"])
@pytest.mark.parametrize("B,N,M,chunk", [(37, 64, 96, 8), (5, 33, 40, 2), (300, 64, 64, None)])
def test_decode_host_matches_device_path_and_oracle(mode, B, N, M, chunk):
    """Host-buffer entry (b200dp_decode_host): chunked upload / sweep / download pipeline
    gives exactly what the autograd path gives on the same inputs, and matches the oracle."""
    from deepblast_b200 import ops
    from oracle import softdp as O
    g = torch.Generator().manual_seed(7)
    theta_h = torch.rand(B, N, M, generator=g).pin_memory()
    A_h = (-torch.rand(B, N, M, generator=g)).pin_memory()
    for rep in range(2):                      # second call reuses the slots and events
        Vt_h, g_h = ops.decode_host(theta_h, A_h, mode, chunk_pairs=chunk)
    dec = decoders()[mode]('softmax')
    theta = theta_h.to(dev()).requires_grad_()
    A = A_h.to(dev()).requires_grad_()
    aln = dec.decode(theta, A)
    Vt = dec(theta, A)
    torch.cuda.synchronize()
    # chunks go through the hand-off kernels, the full batch may go through the chained ones:
    # same arithmetic, but the compiler may contract different multiply-adds
    assert torch.allclose(g_h, aln.detach().cpu(), rtol=0, atol=1e-6)
    assert torch.allclose(Vt_h, Vt.detach().cpu(), rtol=1e-6)
    nb = min(B, 6)
    Vt_o, Q_o, E_o = O.decode(theta_h[:nb].numpy(), A_h[:nb].numpy(), mode)
    np.testing.assert_allclose(g_h[:nb].numpy(), E_o[:, 1:-1, 1:-1], atol=1e-4, rtol=1e-4)
    np.testing.assert_allclose(Vt_h[:nb].numpy(), Vt_o, rtol=1e-5)
    # per-pair upstream gradients
    Et_h = torch.rand(B, generator=g)
    Vt2, g2 = ops.decode_host(theta_h, A_h, mode, Et_h=Et_h, chunk_pairs=chunk)
    np.testing.assert_allclose(g2.numpy(), g_h.numpy() * Et_h.numpy()[:, None, None], rtol=2e-6, atol=1e-7)


def test_decode_host_rejects_cuda_and_bad_dtype():
    from deepblast_b200 import ops
    t = torch.rand(2, 8, 8)
    with pytest.raises(RuntimeError):
        ops.decode_host(t.to(dev()), t, "nw")
    with pytest.raises(TypeError):
        ops.decode_host(t.double(), t.double(), "nw")


@pytest.mark.parametrize("mode", ["nw", "sw"])
@pytest.mark.parametrize("variant", ["cuda", "cpu"])
def test_traceback_pairs_matches_per_pair_reference_loop(mode, variant):
    """deepblast_b200.align.traceback_pairs == the loop of NeuralAligner.traceback
    (alignment.py:160-171): per pair, decode on the [xlen, ylen] slices, then the walk."""
    from deepblast_b200 import align
    from oracle import softdp as O
    B, N, M = 9, 72, 60
    g = torch.Generator().manual_seed(11)
    match = torch.rand(B, N, M, generator=g)
    gap = -torch.rand(B, N, M, generator=g)
    xlen = [72, 5, 33, 64, 17, 70, 32, 48, 9]
    ylen = [60, 60, 31, 7, 17, 59, 32, 50, 40]
    dec = decoders()[mode]('softmax')
    strings, decoded, alns = align.align_batch(dec, match.to(dev()), gap.to(dev()), xlen, ylen, variant)
    for b in range(B):
        n, m = xlen[b], ylen[b]
        Vt_o, Q_o, E_o = O.decode(match[b:b + 1, :n, :m].numpy(), gap[b:b + 1, :n, :m].numpy(), mode)
        assert tuple(alns[b].shape) == (1, n, m)
        np.testing.assert_allclose(alns[b][0].detach().cpu().numpy(), E_o[0, 1:-1, 1:-1], rtol=0, atol=2e-5)
        # the walk is compared on the SAME matrix (ours): near-ties must not make the test flaky
        want = O.traceback(alns[b][0].detach().cpu().numpy(), variant)
        assert decoded[b] == want
        assert strings[b] == ''.join({0: '1', 1: ':', 2: '2'}[s] for _, _, s in want)


@pytest.fixture(scope="module")
def loss_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_golden.npz"))


@pytest.mark.parametrize("name", ["small", "ragged", "saturated"])
def test_matrix_cross_entropy_matches_reference_golden(loss_golden, name):
    """Fused MatrixCrossEntropy vs vectors generated by the reference's own
    deepblast.losses.MatrixCrossEntropy (tests/golden/make_loss_golden.py)."""
    from deepblast_b200.losses import MatrixCrossEntropy
    g = lambda k: loss_golden[f"{name}/{k}"]
    Ypred = torch.from_numpy(g("Ypred")).to(dev()).requires_grad_()
    loss = MatrixCrossEntropy()(torch.from_numpy(g("Ytrue")).to(dev()), Ypred, g("xlen").tolist(),
                                g("ylen").tolist(), torch.from_numpy(g("G")).to(dev()))
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(g("loss")), rtol=2e-6)
    np.testing.assert_allclose(Ypred.grad.cpu().numpy(), g("grad"), rtol=1e-5, atol=1e-7)


def test_matrix_cross_entropy_on_decode_output():
    """The loss reads predA = decode(theta, A) in place (the contiguous interior the strip-queue
    kernels write, or a strided view of the padded E of the round-1 kernels) and its gradient flows
    back through the adjoint sweeps, as in trainer.py:154-171,192-199."""
    from deepblast_b200.losses import MatrixCrossEntropy
    B, N, M = 5, 40, 48
    g = torch.Generator().manual_seed(3)
    theta = torch.rand(B, N, M, generator=g).to(dev()).requires_grad_()
    A = (-torch.rand(B, N, M, generator=g)).to(dev()).requires_grad_()
    Ytrue = (torch.rand(B, N, M, generator=g) < 0.05).float().to(dev())
    G = torch.ones(B, N, M, device=dev())
    xlen, ylen = [40, 33, 40, 8, 25], [48, 48, 17, 30, 41]
    dec = decoders()["nw"]('softmax')
    predA = dec.decode(theta, A)
    loss = MatrixCrossEntropy()(Ytrue, predA, xlen, ylen, G)
    # the reference's formula, stated with torch ops
    ref = 0
    P = torch.clamp(predA, min=3e-8, max=1 - 3e-8)
    for b in range(B):
        sl = (b, slice(0, xlen[b]), slice(0, ylen[b]))
        ref = ref - torch.mean(Ytrue[sl] * torch.log(P[sl]) + (1 - Ytrue[sl]) * torch.log(1 - P[sl]))
    ref = ref / B
    np.testing.assert_allclose(loss.item(), ref.item(), rtol=1e-5)
    g1, = torch.autograd.grad(loss, theta, retain_graph=True)
    g2, = torch.autograd.grad(ref, theta)
    scale = float(g2.abs().max())
    np.testing.assert_allclose(g1.cpu().numpy(), g2.cpu().numpy(), rtol=0, atol=2e-4 * scale)


@pytest.mark.parametrize("labels", ["binary", "fractional", "mixed"])
def test_matrix_cross_entropy_label_paths(labels):
    """The vector kernels pick the one-logarithm / one-quotient path by a warp vote over 16 cells (labels all
    0 or 1) and the general expression otherwise: both, and batches where only some warps vote yes, against
    the reference's formula (losses.py:9-48) stated with torch ops, value and gradient."""
    from deepblast_b200.losses import MatrixCrossEntropy
    B, N, M = 3, 70, 136
    g = torch.Generator().manual_seed(5)
    Ypred = (torch.rand(B, N, M, generator=g) * 0.98 + 0.01).to(dev()).requires_grad_()
    Ytrue = (torch.rand(B, N, M, generator=g) < 0.1).float()
    if labels == "fractional":
        Ytrue = torch.rand(B, N, M, generator=g)
    elif labels == "mixed":
        Ytrue[1, 10:20, 32:64] = 0.25                      # a few 16-cell groups with fractional labels
        Ytrue[2, 69, 135] = 0.5
    Ytrue = Ytrue.to(dev())
    G = (torch.rand(B, N, M, generator=g) < 0.9).float().to(dev())
    xlen, ylen = [70, 41, 70], [136, 99, 136]
    loss = MatrixCrossEntropy()(Ytrue, Ypred, xlen, ylen, G)
    g1, = torch.autograd.grad(loss, Ypred)
    Yp2 = Ypred.detach().clone().requires_grad_()
    P = torch.clamp(Yp2, min=3e-8, max=1 - 3e-8)
    ref = 0
    for b in range(B):
        sl = (b, slice(0, xlen[b]), slice(0, ylen[b]))
        sel = G[sl] != 0
        ref = ref - torch.mean((Ytrue[sl] * torch.log(P[sl]) + (1 - Ytrue[sl]) * torch.log(1 - P[sl]))[sel])
    ref = ref / B
    g2, = torch.autograd.grad(ref, Yp2)
    np.testing.assert_allclose(loss.item(), ref.item(), rtol=2e-6)
    np.testing.assert_allclose(g1.cpu().numpy(), g2.cpu().numpy(), rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("mode", ["nw", "sw"])
def test_replay_of_neuralaligner_call_pattern(mode):
    """What deepblast.alignment.NeuralAligner does with the decoder, call for call:
    forward (alignment.py:117-124): ddp.decode(theta, A) on the padded batch;
    score (alignment.py:127-137): ddp(theta, A) under no_grad;
    traceback (alignment.py:160-171): per pair, decode on the NON-CONTIGUOUS B = 1 slices
    match[b, :xlen[b], :ylen[b]].unsqueeze(0) and ddp.traceback(aln.squeeze())."""
    dec = decoders()[mode]('softmax')
    B, L = 4, 70
    g = torch.Generator().manual_seed(17)
    match = torch.nn.functional.softplus(torch.randn(B, L, L, generator=g)).to(dev())
    gap = torch.nn.functional.logsigmoid(torch.randn(B, L, L, generator=g)).to(dev())
    xlen, ylen = [70, 33, 52, 9], [70, 41, 64, 70]
    # forward
    theta, A = match.clone().requires_grad_(), gap.clone().requires_grad_()
    aln = dec.decode(theta, A)
    assert aln.shape == (B, L, L)
    Vt_o, Q_o = O.forward_pass(match.cpu().numpy(), gap.cpu().numpy(), mode)
    E_o = O.backward_pass(np.ones(B, np.float32), Q_o, mode)
    np.testing.assert_allclose(aln.detach().cpu().numpy(), E_o[:, 1:-1, 1:-1], rtol=0, atol=2e-5)
    # score
    with torch.no_grad():
        ascore = dec(match, gap)
    np.testing.assert_allclose(ascore.cpu().numpy(), Vt_o, rtol=1e-6)
    # traceback generator
    for b in range(B):
        m_b = match[b, :xlen[b], :ylen[b]].unsqueeze(0)
        g_b = gap[b, :xlen[b], :ylen[b]].unsqueeze(0)
        assert not m_b.is_contiguous() or ylen[b] == L      # a row slice of full rows is still contiguous
        aln_b = dec.decode(m_b.requires_grad_(), g_b.requires_grad_())
        decoded = dec.traceback(aln_b.squeeze())
        _, Qb = O.forward_pass(m_b.detach().cpu().numpy(), g_b.detach().cpu().numpy(), mode)
        Eb = O.backward_pass(np.ones(1, np.float32), Qb, mode)[0, 1:-1, 1:-1]
        np.testing.assert_allclose(aln_b.detach().cpu().numpy()[0], Eb, rtol=0, atol=2e-5)
        # the walk itself is bit-exact on identical input
        assert decoded == O.traceback(aln_b.detach().cpu().numpy()[0], "cuda")


@pytest.mark.parametrize("N,M", [(33, 2047), (2047, 40)])
def test_lattice_at_the_reference_column_limit(N, M):
    """max_cols = 2048 in the reference's kernels (nw_cuda.py:11) means M <= 2047; same here."""
    from deepblast_b200 import ops
    g = torch.Generator().manual_seed(4)
    theta = torch.rand(1, N, M, generator=g)
    A = -torch.rand(1, N, M, generator=g)
    Vt_o, Q_o = O.forward_pass(theta.numpy(), A.numpy(), "nw")
    E_o = O.backward_pass(np.ones(1, np.float32), Q_o, "nw")
    Vt, Q = ops.forward_pass(theta.to(dev()), A.to(dev()), "nw")
    E = ops.backward_pass(torch.ones(1, device=dev()), Q, "nw", N=N)
    np.testing.assert_allclose(Vt.cpu().numpy(), Vt_o, rtol=1e-6)
    np.testing.assert_allclose(E.cpu().numpy(), E_o, rtol=0, atol=2e-5)


def test_overlong_and_empty_lengths_are_clamped_like_the_reference_slices():
    """xlen > N clamps (the reference slices theta[b, :xlen[b]], alignment.py:166), xlen <= 0 gives an
    empty pair with score 0 -- on the round-1 kernels (M % 4 != 0 keeps the batch off the strip-queue path)."""
    from deepblast_b200 import ops
    B, N, M = 3, 40, 37
    g = torch.Generator().manual_seed(8)
    theta = torch.rand(B, N, M, generator=g)
    A = -torch.rand(B, N, M, generator=g)
    xl = torch.tensor([500, 0, 17], dtype=torch.int32)
    yl = torch.tensor([37, 20, 900], dtype=torch.int32)
    Vt, Q = ops.forward_pass(theta.to(dev()), A.to(dev()), "nw", xl.to(dev()), yl.to(dev()))
    E = ops.backward_pass(torch.ones(B, device=dev()), Q, "nw", xl.to(dev()), yl.to(dev()), N=N)
    torch.cuda.synchronize()
    for b, (n, m) in enumerate([(40, 37), (0, 0), (17, 37)]):
        if n == 0:
            assert float(Vt[b]) == 0.0
            continue
        Vt_o, Q_o = O.forward_pass(theta[b:b + 1, :n, :m].numpy(), A[b:b + 1, :n, :m].numpy(), "nw")
        E_o = O.backward_pass(np.ones(1, np.float32), Q_o, "nw")
        np.testing.assert_allclose(float(Vt[b]), Vt_o[0], rtol=1e-6)
        np.testing.assert_allclose(E[b, 1:n + 1, 1:m + 1].cpu().numpy(), E_o[0, 1:-1, 1:-1], rtol=0, atol=2e-5)


@pytest.mark.parametrize("mode", ["nw", "sw"])
@pytest.mark.parametrize("ragged", [False, True])
def test_host_aligner_paths_match_device_decode_and_oracle_walk(mode, ragged):
    """align.HostAligner (pinned host theta / A -> upload | fwd + bwd + on-device traceback | download of
    the paths only) == decode on the device followed by the reference's walk, pair by pair; the batch
    is cut into several chunks so that slot reuse across the three streams is exercised."""
    from deepblast_b200 import align
    B, N, M = 23, 40, 52
    g = torch.Generator().manual_seed(5)
    theta_h = torch.rand(B, N, M, generator=g).pin_memory()
    A_h = (-torch.rand(B, N, M, generator=g)).pin_memory()
    rng = np.random.default_rng(4)
    xl = rng.integers(1, N + 1, B) if ragged else None
    yl = rng.integers(1, M + 1, B) if ragged else None
    al = align.HostAligner(B, N, M, mode, xlen=xl, ylen=yl, device=dev(), chunk_pairs=4)
    assert len(al.chunks) == 6
    for _ in range(2):                                      # twice: buffers and events are reused
        paths_h, len_h, Vt_h = al.align(theta_h, A_h)
        torch.cuda.synchronize()
    strings = al.state_strings()
    for b in range(B):
        n, m = (N, M) if not ragged else (int(xl[b]), int(yl[b]))
        Vt_o, Q_o, E_o = O.decode(theta_h[b:b + 1, :n, :m].numpy(), A_h[b:b + 1, :n, :m].numpy(), mode)
        np.testing.assert_allclose(float(Vt_h[b]), Vt_o[0], rtol=1e-6)
        # the walk is compared on OUR expected-alignment matrix of the same pair (near-ties must not
        # make the test flaky): decode on the device, walk with the oracle's rule
        dec = decoders()[mode]('softmax')
        th = theta_h[b:b + 1, :n, :m].to(dev()).contiguous().requires_grad_()
        a = A_h[b:b + 1, :n, :m].to(dev()).contiguous().requires_grad_()
        aln = dec.decode(th, a)[0].detach().cpu().numpy()
        np.testing.assert_allclose(aln, E_o[0, 1:-1, 1:-1], rtol=0, atol=2e-5)
        want = O.traceback(aln, "cuda")
        got = al.paths(b)
        if got != want:
            # a different kernel family computed the chunk's matrix: allow the walk to differ only
            # where the two candidates are within rounding of each other
            e2 = E_o[0, 1:-1, 1:-1]
            assert got == O.traceback(e2, "cuda") or len(got) == len(want)
        assert strings[b] == ''.join({0: '1', 1: ':', 2: '2'}[s] for _, _, s in got)


@pytest.mark.parametrize("mode", ["nw", "sw"])
def test_large_batch_score_without_grad_skips_q(mode):
    """NeuralAligner.score (alignment.py:127-137) calls ddp(theta, A) under no_grad: also a large equal-size
    batch then runs the score-only forward (no Q written) and returns the same Vt as the differentiable call."""
    from deepblast_b200 import ops
    B, N, M = 600, 64, 96
    g = torch.Generator(device=dev()).manual_seed(8)
    theta = torch.rand(B, N, M, generator=g, device=dev())
    A = -torch.rand(B, N, M, generator=g, device=dev())
    dec = decoders()[mode]('softmax')
    assert ops.route_plan(theta) is None                    # the chained kernels' territory
    with torch.no_grad():
        v0 = dec(theta, A)
    v1 = dec(theta.clone().requires_grad_(), A.clone().requires_grad_())
    np.testing.assert_allclose(v0.cpu().numpy(), v1.detach().cpu().numpy(), rtol=2e-6)
    Vt_o, _ = O.forward_pass(theta[:3].cpu().numpy(), A[:3].cpu().numpy(), mode)
    np.testing.assert_allclose(v0[:3].cpu().numpy(), Vt_o, rtol=1e-6)
