"""GPU parity of the strip-queue kernels (csrc/softdp_sq.cuh, through the C ABI b200dp_sq_* and
the plan builder) against the CPU oracle: equal-size, ragged-dense and PACKED batches, forward,
backward and the adjoint pair, the score-only forward, and BASELINE configs[4] at its stated size
(1024 pairs per GPU, lengths 64..1024 Zipf).  Ragged semantics = the reference's per-pair loop
(deepblast/alignment.py:165-169): pair b is the slice theta[b, :n_b, :m_b] computed on its own.

Tolerances as in test_gpu_parity.py: Vt rtol 1e-6, Q / E atol 1e-5 / 2e-5 (bar: 1e-4).
"""
import numpy as np
import pytest
import torch

from oracle import softdp as O

pytestmark = pytest.mark.gpu
ATOL = 1e-5


def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    from deepblast_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def planmod():
    from deepblast_b200 import plan as _plan
    return _plan


def rand_batch(B, N, M, seed=2):
    g = torch.Generator().manual_seed(seed)
    theta = torch.rand(B, N, M, generator=g)
    A = -torch.rand(B, N, M, generator=g)
    Zt = torch.randn(B, N, M, generator=g)
    ZA = torch.randn(B, N, M, generator=g) * 0.1
    return theta, A, Zt, ZA


def oracle_pair(theta, A, Et, Zt, ZA, n, m, mode):
    """All passes of one pair on its own n x m slice; returns interiors."""
    th, a = theta[None, :n, :m].numpy(), A[None, :n, :m].numpy()
    Vt, Q = O.forward_pass(th, a, mode)
    E = O.backward_pass(np.array([Et], np.float32), Q, mode)
    Ztp = np.zeros((1, n + 2, m + 2), np.float32)
    Ztp[0, 1:-1, 1:-1] = Zt[:n, :m].numpy()
    za = np.zeros((1, n, m), np.float32) if ZA is None else ZA[None, :n, :m].numpy()
    Vtd, Qd = O.adjoint_forward_pass(Q, Ztp, za)
    Ed = O.adjoint_backward_pass(E, Q, Qd)
    return Vt[0], Q[0], E[0, 1:-1, 1:-1], Vtd[0], Qd[0], Ed[0, 1:-1, 1:-1]


def run_and_check(ops, plan, theta, A, Zt, ZA, mode, pairs=None, flags=0, check_q=True, use_za=True,
                  vt_rtol=1e-6, vtd_atol=2e-5):
    """Run the four sweeps on the plan's layout and compare `pairs` with the per-pair oracle."""
    B = plan.B
    d = dev()
    conv = (lambda x: plan.pack(x.to(d))) if plan.packed else (lambda x: x.to(d).contiguous())
    th_d, a_d, zt_d = conv(theta), conv(A), conv(Zt)
    za_d = conv(ZA) if use_za else None
    Et = torch.linspace(0.5, 1.5, B)
    Vt, Q = ops.sq_forward(plan, th_d, a_d, mode, flags=flags)
    E = ops.sq_backward(plan, Et.to(d), Q, mode, flags=flags)
    Vtd, QdE = ops.sq_adjoint_forward(plan, Q, zt_d, za_d, E, flags=flags)
    Ed = ops.sq_adjoint_backward(plan, Q, QdE, flags=flags)
    Vt_s, _ = ops.sq_forward(plan, th_d, a_d, mode, need_q=False, flags=flags)
    torch.cuda.synchronize()
    assert torch.equal(Vt_s, Vt)                               # score-only forward: the same Vt bit for bit
    Vt, Vtd = Vt.cpu().numpy(), Vtd.cpu().numpy()
    for b in (range(B) if pairs is None else pairs):
        n, m = int(plan.xlen[b]), int(plan.ylen[b])
        if n == 0 or m == 0:
            assert Vt[b] == 0.0 and Vtd[b] == 0.0
            continue
        Vt_o, Q_o, E_o, Vtd_o, Qd_o, Ed_o = oracle_pair(theta[b], A[b], float(Et[b]), Zt[b], ZA[b] if use_za else None,
                                                        n, m, mode)
        np.testing.assert_allclose(Vt[b], Vt_o, rtol=vt_rtol, err_msg=f"Vt pair {b} ({n}x{m})")
        if check_q:
            np.testing.assert_allclose(ops.sq_q_to_reference(plan, Q, b).cpu().numpy(), Q_o, rtol=0, atol=ATOL,
                                       err_msg=f"Q pair {b} ({n}x{m})")
        np.testing.assert_allclose(plan.pair_view(E, b).cpu().numpy(), E_o, rtol=0, atol=2 * ATOL,
                                   err_msg=f"E pair {b} ({n}x{m})")
        sc = max(1.0, float(np.abs(Vtd_o)), float(np.abs(Ed_o).max()))
        np.testing.assert_allclose(Vtd[b], Vtd_o, rtol=0, atol=vtd_atol * sc, err_msg=f"Vtd pair {b} ({n}x{m})")
        np.testing.assert_allclose(plan.pair_view(Ed, b).cpu().numpy(), Ed_o, rtol=0, atol=1e-4 * sc,
                                   err_msg=f"Ed pair {b} ({n}x{m})")
    return E, Ed


SHAPES = [(2, 1, 4), (2, 5, 4), (3, 31, 36), (2, 32, 32), (2, 64, 64), (2, 65, 64), (2, 96, 200),
          (1, 300, 76), (2, 256, 256), (1, 130, 520)]


@pytest.mark.parametrize("B,N,M", SHAPES)
@pytest.mark.parametrize("mode", ["nw", "sw"])
def test_sq_equal_size_vs_oracle(ops, planmod, B, N, M, mode):
    theta, A, Zt, ZA = rand_batch(B, N, M)
    plan = planmod.Plan(B, N, M, device=dev())
    run_and_check(ops, plan, theta, A, Zt, ZA, mode)


LENS_X = [200, 1, 33, 64, 0, 199, 32, 150, 97, 200, 7, 500]      # 500: clamped to N like the reference's slice
LENS_Y = [152, 9, 40, 152, 17, 1, 32, 0, 151, 4, 152, 100]


@pytest.mark.parametrize("mode", ["nw", "sw"])
@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("grid", [0, 1, 3])
def test_sq_ragged_vs_per_pair_oracle(ops, planmod, mode, packed, grid):
    """Ragged lengths incl. empty, single-row / single-column and over-long (clamped) pairs, dense
    and packed, also with a grid of 1 and 3 warps (every strip then waits on the queue order)."""
    B, N, M = len(LENS_X), 200, 152
    theta, A, Zt, ZA = rand_batch(B, N, M, seed=7)
    plan = planmod.Plan(B, N, M, LENS_X, LENS_Y, packed=packed, device=dev())
    E, Ed = run_and_check(ops, plan, theta, A, Zt, ZA, mode, flags=grid << ops.CTAS_SHIFT)
    if not packed:
        # outside each pair's corner the dense gradient is exactly zero
        mask = torch.ones(B, N, M, dtype=torch.bool)
        for b in range(B):
            mask[b, :int(plan.xlen[b]), :int(plan.ylen[b])] = False
        assert float(E.cpu()[mask].abs().max()) == 0.0 and float(Ed.cpu()[mask].abs().max()) == 0.0


@pytest.mark.parametrize("mode", ["nw", "sw"])
def test_sq_adjoint_without_za_and_plain_qd(ops, planmod, mode):
    """ZA = None (the usual double backward) and E = None (the stream then holds Qd itself)."""
    B, N, M = 3, 100, 72
    theta, A, Zt, ZA = rand_batch(B, N, M, seed=11)
    plan = planmod.Plan(B, N, M, device=dev())
    run_and_check(ops, plan, theta, A, Zt, ZA, mode, use_za=False)
    d = dev()
    Vt, Q = ops.sq_forward(plan, theta.to(d), A.to(d), mode)
    Vtd, Qd = ops.sq_adjoint_forward(plan, Q, Zt.to(d), ZA.to(d), None)
    for b in range(B):
        _, _, _, Vtd_o, Qd_o, _ = oracle_pair(theta[b], A[b], 1.0, Zt[b], ZA[b], N, M, mode)
        sc = max(1.0, float(np.abs(Qd_o).max()))
        got = ops.sq_q_to_reference(plan, Qd, b, "qd").cpu().numpy()
        np.testing.assert_allclose(got[1:-1, 1:-1], Qd_o[1:-1, 1:-1], rtol=0, atol=2e-5 * sc)


def zipf_lengths(B, rng):
    k = np.arange(1, 17)
    pk = (1.0 / k) / (1.0 / k).sum()
    return 64 * rng.choice(k, size=B, p=pk), 64 * rng.choice(k, size=B, p=pk)


@pytest.mark.parametrize("packed", [True, False])
def test_sq_c5_stated_size(ops, planmod, packed):
    """BASELINE configs[4] as bench.py runs it: 1024 pairs per GPU, lengths 64..1024 (Zipf over
    16 buckets, rng seed 0).  Forward, backward AND the adjoint pair of 18 sampled pairs -- a
    1024 x 1024, a 64 x 1024, a 1024 x 64 and a 64 x 64 among them -- against the per-pair oracle."""
    B = 1024
    xl, yl = zipf_lengths(B, np.random.default_rng(0))
    xl[0], yl[0] = 1024, 1024
    xl[1], yl[1] = 64, 1024
    xl[2], yl[2] = 1024, 64
    xl[3], yl[3] = 64, 64
    N, M = 1024, 1024
    plan = planmod.Plan(B, N, M, xl, yl, packed=packed, device=dev())
    rng = np.random.default_rng(5)
    sample = sorted(set([0, 1, 2, 3] + rng.choice(B, 14, replace=False).tolist()))
    d = dev()
    g = torch.Generator(device=d).manual_seed(3)
    if packed:
        nf = plan.packed_floats
        th_d = torch.rand(nf, generator=g, device=d)
        a_d = -torch.rand(nf, generator=g, device=d)
        zt_d = torch.randn(nf, generator=g, device=d)
    else:
        th_d = torch.rand(B, N, M, generator=g, device=d)
        a_d = -torch.rand(B, N, M, generator=g, device=d)
        zt_d = torch.randn(B, N, M, generator=g, device=d)
    Et = torch.ones(B, device=d)
    Vt, Q = ops.sq_forward(plan, th_d, a_d, "nw")
    E = ops.sq_backward(plan, Et, Q, "nw")
    Vtd, QdE = ops.sq_adjoint_forward(plan, Q, zt_d, None, E)
    Ed = ops.sq_adjoint_backward(plan, Q, QdE)
    torch.cuda.synchronize()
    Vt, Vtd = Vt.cpu().numpy(), Vtd.cpu().numpy()
    assert np.isfinite(Vt).all() and np.isfinite(Vtd).all()
    for b in sample:
        n, m = int(plan.xlen[b]), int(plan.ylen[b])
        th = plan.pair_view(th_d, b).cpu()
        a = plan.pair_view(a_d, b).cpu()
        zt = plan.pair_view(zt_d, b).cpu()
        Vt_o, Q_o, E_o, Vtd_o, Qd_o, Ed_o = oracle_pair(th, a, 1.0, zt, None, n, m, "nw")
        np.testing.assert_allclose(Vt[b], Vt_o, rtol=1e-6, err_msg=f"Vt pair {b} ({n}x{m})")
        np.testing.assert_allclose(plan.pair_view(E, b).cpu().numpy(), E_o, rtol=0, atol=2 * ATOL,
                                   err_msg=f"E pair {b} ({n}x{m})")
        sc = max(1.0, float(np.abs(Vtd_o)), float(np.abs(Ed_o).max()))
        # (a sum over a 2048-step path of fp32 steps: a few 1e-5 relative at 1024 x 1024; bar 1e-4)
        np.testing.assert_allclose(Vtd[b], Vtd_o, rtol=0, atol=1e-4 * sc, err_msg=f"Vtd pair {b} ({n}x{m})")
        np.testing.assert_allclose(plan.pair_view(Ed, b).cpu().numpy(), Ed_o, rtol=0, atol=1e-4 * sc,
                                   err_msg=f"Ed pair {b} ({n}x{m})")


def test_sq_large_equal_batch_matches_chained_kernels(ops, planmod):
    """C2 shape: the strip-queue kernels and the chained kernels agree on every pair (both are
    within 1e-5 of the oracle, so within 2e-5 of each other), and 3 sampled pairs match the oracle."""
    B, N, M = 1024, 256, 256
    d = dev()
    g = torch.Generator(device=d).manual_seed(2)
    theta = torch.rand(B, N, M, generator=g, device=d)
    A = -torch.rand(B, N, M, generator=g, device=d)
    plan = planmod.Plan(B, N, M, device=d)
    Vt, Q = ops.sq_forward(plan, theta, A, "nw")
    E = ops.sq_backward(plan, torch.ones(B, device=d), Q, "nw")
    Vt3, Q3 = ops.forward_pass(theta, A, "nw")
    E3 = ops.backward_pass(torch.ones(B, device=d), Q3, "nw", N=N)
    np.testing.assert_allclose(Vt.cpu().numpy(), Vt3.cpu().numpy(), rtol=1e-6)
    assert float((E - E3[:, 1:-1, 1:-1]).abs().max()) < 2e-5
    for b in (0, 511, 1023):
        Vt_o, Q_o, E_o, *_ = oracle_pair(theta[b].cpu(), A[b].cpu(), 1.0, torch.zeros(N, M), None, N, M, "nw")
        np.testing.assert_allclose(Vt[b].item(), Vt_o, rtol=1e-6)
        np.testing.assert_allclose(E[b].cpu().numpy(), E_o, rtol=0, atol=2 * ATOL)


def test_sq_workspace_is_left_clean_and_reusable(ops, planmod):
    """Back-to-back launches on one workspace (different plans, growing epochs) stay correct."""
    theta, A, Zt, ZA = rand_batch(4, 160, 96, seed=21)
    p1 = planmod.Plan(4, 160, 96, device=dev())
    p2 = planmod.Plan(4, 160, 96, [160, 100, 31, 64], [96, 50, 96, 8], device=dev())
    for _ in range(3):
        run_and_check(ops, p1, theta, A, Zt, ZA, "nw", check_q=False)
        run_and_check(ops, p2, theta, A, Zt, ZA, "nw", check_q=False)


@pytest.mark.parametrize("mode,cls", [("nw", "nw_cuda.NeedlemanWunschDecoder"), ("sw", "sw_cuda.SmithWatermanDecoder")])
@pytest.mark.parametrize("packed", [False, True])
def test_sq_autograd_decode_and_double_backward(planmod, mode, cls, packed):
    """Decoder.decode with per-pair lengths (dense) and with a packed plan: aln and the gradient of
    a weighted sum of aln (the double backward of a training step) match the per-pair oracle."""
    import importlib
    modname, clsname = cls.split(".")
    Dec = getattr(importlib.import_module("deepblast_b200." + modname), clsname)
    B, N, M = 6, 96, 80
    xl, yl = [96, 50, 33, 96, 1, 64], [80, 80, 17, 4, 80, 64]
    theta, A, Zt, _ = rand_batch(B, N, M, seed=13)
    d = dev()
    dec = Dec('softmax')
    plan = planmod.get_plan(B, N, M, xl, yl, packed, d)
    if packed:
        th = plan.pack(theta.to(d)).requires_grad_()
        a = plan.pack(A.to(d)).requires_grad_()
        w = plan.pack(Zt.to(d))
        aln = dec.decode(th, a, plan=plan)
    else:
        th = theta.to(d).requires_grad_()
        a = A.to(d).requires_grad_()
        w = Zt.to(d)
        aln = dec.decode(th, a, torch.tensor(xl), torch.tensor(yl))
    (aln * w).sum().backward()
    for b in range(B):
        n, m = xl[b], yl[b]
        _, _, E_o, _, _, Ed_o = oracle_pair(theta[b], A[b], 1.0, Zt[b], None, n, m, mode)
        np.testing.assert_allclose(plan.pair_view(aln.detach(), b).cpu().numpy(), E_o, rtol=0, atol=2 * ATOL)
        sc = max(1.0, float(np.abs(Ed_o).max()))
        np.testing.assert_allclose(plan.pair_view(th.grad, b).cpu().numpy(), Ed_o, rtol=0, atol=1e-4 * sc)


@pytest.mark.parametrize("mode", ["nw", "sw"])
def test_sq_dense_tma_staging_equals_ldgsts_staging(ops, planmod, mode):
    """Dense plans stage theta / A with TMA boxes (b200dp_sq_fwd_dense), packed plans and the fallback
    with per-lane 16-byte copies (b200dp_sq_fwd): the same shared-memory image, so Q and Vt must agree
    bit for bit -- equal-size lattices, ragged lengths inside a dense tensor, M not a multiple of 16."""
    d = dev()
    for B, N, M, ragged in ((3, 96, 200, False), (4, 130, 260, True), (2, 256, 256, False), (5, 64, 68, True)):
        theta, A, _, _ = rand_batch(B, N, M, seed=7)
        rng = np.random.default_rng(3)
        xl = rng.integers(1, N + 1, B) if ragged else None
        yl = rng.integers(1, M + 1, B) if ragged else None
        plan = planmod.Plan(B, N, M, xl, yl, packed=False, device=d)
        out = {}
        for tma in (True, False):
            ops.SQ_TMA_OPERANDS = tma
            try:
                Vt, Q = ops.sq_forward(plan, theta.to(d), A.to(d), mode)
                Vs, _ = ops.sq_forward(plan, theta.to(d), A.to(d), mode, need_q=False)
            finally:
                ops.SQ_TMA_OPERANDS = True
            torch.cuda.synchronize()
            out[tma] = (Vt.cpu(), Q.cpu(), Vs.cpu())
        assert torch.equal(out[True][0], out[False][0])
        assert torch.equal(out[True][2], out[False][2])
        for b in range(B):                                  # (storage between pairs' streams is never written)
            qa = ops.sq_q_to_reference(plan, out[True][1].to(d), b).cpu()
            qb = ops.sq_q_to_reference(plan, out[False][1].to(d), b).cpu()
            assert torch.equal(qa, qb)


@pytest.mark.parametrize("mode", ["nw", "sw"])
@pytest.mark.parametrize("B,N,M,cs,W", [(3, 96, 128, 0, 0), (2, 130, 260, 0, 0), (100, 64, 128, 8, 1), (37, 70, 72, 8, 2),
                                        (5, 300, 200, 4, 4), (2, 1024, 512, 0, 0), (9, 33, 64, 2, 2)])
def test_cluster_kernels_equal_strip_queue_kernels(ops, planmod, mode, B, N, M, cs, W):
    """Cluster kernels (softdp_cl.cuh: the strips of a pair dealt round-robin to the warps of a thread-block
    cluster, boundary rows handed over through distributed shared memory) against the strip-queue kernels:
    the same cell arithmetic in the same order, so Vt, Q and E must agree bit for bit -- partial strips,
    pairs with fewer strips than the ring has warps (flow control: a producer must not lap its consumer),
    more pairs than resident clusters, forced cluster sizes and warps per CTA, score-only."""
    from deepblast_b200 import _lib
    L = _lib.lib()
    d = dev()
    theta, A, _, _ = rand_batch(B, N, M, seed=9)
    th_d, a_d = theta.to(d), A.to(d)
    Et = torch.linspace(0.5, 1.5, B, device=d)
    plan = planmod.Plan(B, N, M, device=d)
    old = ops.CLUSTER
    ops.CLUSTER = False
    try:
        Vt0, Q0 = ops.sq_forward(plan, th_d, a_d, mode)
        E0 = ops.sq_backward(plan, Et, Q0, mode)
    finally:
        ops.CLUSTER = old
    fl = (cs << 4) | (W << 8)
    st = torch.cuda.current_stream().cuda_stream
    Q = torch.zeros_like(Q0)
    Vt, Vs = torch.empty_like(Vt0), torch.empty_like(Vt0)
    E = torch.full_like(E0, float("nan"))
    for _ in range(2):                                      # twice: the second launch finds warm caches and other timing
        _lib.check(L.b200dp_cl_fwd(th_d.data_ptr(), a_d.data_ptr(), Q.data_ptr(), Vt.data_ptr(), B, N, M, ops.MODES[mode], fl, st),
                   "b200dp_cl_fwd")
        _lib.check(L.b200dp_cl_fwd(th_d.data_ptr(), a_d.data_ptr(), None, Vs.data_ptr(), B, N, M, ops.MODES[mode], fl, st),
                   "b200dp_cl_fwd")
        _lib.check(L.b200dp_cl_bwd(Et.data_ptr(), Et.stride(0), Q.data_ptr(), E.data_ptr(), B, N, M, ops.MODES[mode], fl, st),
                   "b200dp_cl_bwd")
    torch.cuda.synchronize()
    assert torch.equal(Vt, Vt0) and torch.equal(Vs, Vt0)
    assert torch.equal(E, E0)
    for b in range(0, B, max(1, B // 7)):
        assert torch.equal(ops.sq_q_to_reference(plan, Q, b), ops.sq_q_to_reference(plan, Q0, b))


def test_small_batches_are_routed_to_the_cluster_forward(ops, planmod):
    """ops.sq_forward sends small dense equal-size batches to b200dp_cl_fwd (b200dp_cl_applicable) and the
    result still matches the per-pair oracle; ragged and packed plans never take that path."""
    from deepblast_b200 import _lib
    d = dev()
    assert _lib.lib().b200dp_cl_applicable(2, 256, 256) > 0
    assert _lib.lib().b200dp_cl_applicable(1024, 256, 256) == 0
    assert _lib.lib().b200dp_cl_applicable(4, 32, 256) == 0          # a single strip has nothing to hand over
    B, N, M = 2, 96, 200
    theta, A, Zt, ZA = rand_batch(B, N, M, seed=4)
    plan = planmod.Plan(B, N, M, device=d)
    assert ops._use_cluster(plan, 0) > 0
    run_and_check(ops, plan, theta, A, Zt, ZA, "nw")
    rag = planmod.Plan(B, N, M, [90, 40], [200, 64], device=d)
    assert ops._use_cluster(rag, 0) == 0
