"""GPU parity of the fused theta / A producer (csrc/softdp_gemm.cu: tcgen05 GEMM with a bf16 hi/lo split
and the softplus / logsigmoid epilogue, C ABI b200dp_theta_a) against plain fp32 torch
(deepblast/alignment.py:122-123), and end to end through the DP against the oracle.

Tolerance: the bf16 hi/lo split carries 16 bits of mantissa per operand, i.e. a relative error of about
1e-5 of the inner product's scale (measured: 1.8e-5 worst case).  On inner products of the size the
activations are sensitive to (|s| up to ~6; softplus and logsigmoid are linear beyond) that is inside the
north star's 1e-4 absolute bar, which is what these tests assert; larger values are held to 2e-5
relative.  The DP outputs computed FROM our theta / A are compared with the oracle run on torch's
theta / A at 2e-4 (the two error bars added)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import softdp as O

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def embeddings(B, Lx, Ly, D, seed=0, scale=1.0):
    """Entries with standard deviation sqrt(1.2 * scale) / D^(1/4): inner products ~ N(0, (1.2 * scale)^2)."""
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(B, L, D, generator=g) * ((1.2 * scale) ** 0.5 / D ** 0.25)).to(dev()) for L in (Lx, Ly, Lx, Ly)]


def reference(zx, zy, gx, gy):
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        theta = F.softplus(torch.einsum('bid,bjd->bij', zx.double(), zy.double())).float()
        A = F.logsigmoid(torch.einsum('bid,bjd->bij', gx.double(), gy.double())).float()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return theta, A


@pytest.mark.parametrize("B,Lx,Ly,D", [(2, 128, 128, 64), (3, 200, 150, 128), (1, 64, 300, 192), (2, 256, 256, 1024),
                                       (5, 33, 17, 64)])
def test_theta_a_vs_torch(B, Lx, Ly, D):
    from deepblast_b200 import producer
    zx, zy, gx, gy = embeddings(B, Lx, Ly, D)
    theta, A = producer.theta_a(zx, zy, gx, gy)
    th_ref, a_ref = reference(zx, zy, gx, gy)
    np.testing.assert_allclose(theta.cpu().numpy(), th_ref.cpu().numpy(), rtol=0, atol=1e-4)
    np.testing.assert_allclose(A.cpu().numpy(), a_ref.cpu().numpy(), rtol=0, atol=1e-4)


def test_theta_a_prepass_variant_vs_torch(monkeypatch):
    """Version 1 of the kernel (bf16 hi/lo copies written by a pre-pass, operands through TMA into a
    3-stage ring) stays selectable and correct."""
    from deepblast_b200 import producer
    monkeypatch.setattr(producer, "PREPASS", True)
    zx, zy, gx, gy = embeddings(3, 200, 150, 128, seed=2)
    theta, A = producer.theta_a(zx, zy, gx, gy)
    th_ref, a_ref = reference(zx, zy, gx, gy)
    np.testing.assert_allclose(theta.cpu().numpy(), th_ref.cpu().numpy(), rtol=0, atol=1e-4)
    np.testing.assert_allclose(A.cpu().numpy(), a_ref.cpu().numpy(), rtol=0, atol=1e-4)


def test_theta_a_large_values_take_the_linear_branches():
    """softplus above torch's threshold (20) and logsigmoid far in both tails, on inner products of
    standard deviation 14: the error of the hi/lo split is bounded by the products' magnitudes,
    |err(s_ij)| <= 1.5e-5 * sum_d |x_id| |y_jd| (a forward error bound of the usual BLAS form; the
    activations have slope <= 1), not by |s_ij| itself."""
    from deepblast_b200 import producer
    B, L, D = 1, 128, 64
    zx, zy, gx, gy = embeddings(B, L, L, D, seed=3, scale=12.0)
    theta, A = producer.theta_a(zx, zy, gx, gy)
    th_ref, a_ref = reference(zx, zy, gx, gy)
    assert float(th_ref.max()) > 25 and float(a_ref.min()) < -25
    bt = torch.einsum('bid,bjd->bij', zx.abs().double(), zy.abs().double()).float()
    ba = torch.einsum('bid,bjd->bij', gx.abs().double(), gy.abs().double()).float()
    assert bool(((theta - th_ref).abs() <= 1.5e-5 * bt + 1e-6).all())
    assert bool(((A - a_ref).abs() <= 1.5e-5 * ba + 1e-6).all())
    # the linear branches themselves: where s > 20 softplus(s) = s, where s < -20 logsigmoid(s) = s
    s_t = torch.einsum('bid,bjd->bij', zx.double(), zy.double()).float()
    big = s_t > 21
    assert bool(big.any()) and bool(((theta - s_t).abs()[big] <= 1.5e-5 * bt[big] + 1e-6).all())


def test_theta_a_packed_feeds_the_dp_and_matches_the_oracle():
    """Ragged batch: the producer writes theta / A straight into the packed layout the strip-queue
    kernels read; decode on it matches the per-pair oracle run on torch's theta / A."""
    from deepblast_b200 import producer, plan as P
    from deepblast_b200.nw_cuda import NeedlemanWunschDecoder
    B, Lx, Ly, D = 6, 160, 200, 128
    xl, yl = [160, 100, 33, 128, 7, 64], [200, 64, 150, 129, 200, 32]
    zx, zy, gx, gy = embeddings(B, Lx, Ly, D, seed=5)
    plan = P.get_plan(B, Lx, Ly, xl, yl, True, dev())
    theta, A = producer.theta_a(zx, zy, gx, gy, plan=plan)
    th_ref, a_ref = reference(zx, zy, gx, gy)
    for b in range(B):
        np.testing.assert_allclose(plan.pair_view(theta, b).cpu().numpy(), th_ref[b, :xl[b], :yl[b]].cpu().numpy(),
                                   rtol=0, atol=1e-4)
        np.testing.assert_allclose(plan.pair_view(A, b).cpu().numpy(), a_ref[b, :xl[b], :yl[b]].cpu().numpy(),
                                   rtol=0, atol=1e-4)
    aln = NeedlemanWunschDecoder('softmax').decode(theta.requires_grad_(), A.requires_grad_(), plan=plan)
    for b in range(B):
        n, m = xl[b], yl[b]
        Vt_o, Q_o = O.forward_pass(th_ref[b:b + 1, :n, :m].cpu().numpy(), a_ref[b:b + 1, :n, :m].cpu().numpy(), "nw")
        E_o = O.backward_pass(np.ones(1, np.float32), Q_o, "nw")
        np.testing.assert_allclose(plan.pair_view(aln.detach(), b).cpu().numpy(), E_o[0, 1:-1, 1:-1], rtol=0, atol=2e-4)


def test_theta_a_autograd_matches_torch():
    """Gradients of a scalar of (theta, A) with respect to the four embeddings against torch's own
    autograd of the reference expression."""
    from deepblast_b200.producer import ThetaA
    B, Lx, Ly, D = 2, 96, 80, 64
    zs = [t.requires_grad_() for t in embeddings(B, Lx, Ly, D, seed=9)]
    g = torch.Generator().manual_seed(1)
    w1, w2 = torch.randn(B, Lx, Ly, generator=g).to(dev()), torch.randn(B, Lx, Ly, generator=g).to(dev())
    theta, A = ThetaA.apply(*zs)
    ((theta * w1).sum() + (A * w2).sum()).backward()
    got = [z.grad.clone() for z in zs]
    for z in zs:
        z.grad = None
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        th = F.softplus(torch.einsum('bid,bjd->bij', zs[0], zs[1]))
        a = F.logsigmoid(torch.einsum('bid,bjd->bij', zs[2], zs[3]))
        ((th * w1).sum() + (a * w2).sum()).backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    for z, gg in zip(zs, got):
        sc = max(1.0, float(z.grad.abs().max()))
        np.testing.assert_allclose(gg.cpu().numpy(), z.grad.cpu().numpy(), rtol=0, atol=2e-4 * sc)
