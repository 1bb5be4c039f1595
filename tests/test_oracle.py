"""Pin the CPU oracle (oracle/softdp_oracle.c) against vectors generated from
the reference itself (tests/golden/make_golden.py) and against the reference
tests' own known answers (deepblast/tests/test_nw.py:43-54, test_sw.py:42-52,
test_nw_cuda.py:64-76, test_sw_cuda.py:58-70)."""
import json
import os

import numpy as np
import pytest

from oracle import softdp as O

CASES = ["t_nw_cuda_5x5", "t_nw_4x4", "r8x8", "r17x23", "r33x40", "r64x48",
         "r40x70", "r12x9_f64"]


def tol(dtype):
    # the oracle repeats the reference's fp64 arithmetic; after the cast to the
    # tensor dtype the two agree to the last place or one ulp (libm exp/log)
    return dict(rtol=0, atol=2e-7) if dtype == np.float32 else dict(rtol=0, atol=1e-13)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", ["nw", "sw"])
def test_passes_match_reference(golden, name, mode):
    g = lambda k: golden[f"{name}/{k}"]
    theta, A, Et, Zt, ZA = g("theta"), g("A"), g("Et"), g("Zt"), g("ZA")
    t = tol(theta.dtype)
    Vt, Q = O.forward_pass(theta, A, mode)
    np.testing.assert_allclose(Vt, g(f"{mode}/Vt"), rtol=1e-6 if theta.dtype == np.float32 else 1e-14)
    np.testing.assert_allclose(Q, g(f"{mode}/Q"), **t)
    E = O.backward_pass(Et, g(f"{mode}/Q"), mode)
    np.testing.assert_allclose(E, g(f"{mode}/E"), **t)
    Vtd, Qd = O.adjoint_forward_pass(g(f"{mode}/Q"), Zt, ZA)
    scale = max(1.0, float(np.abs(g(f"{mode}/Vtd")).max()))
    np.testing.assert_allclose(Vtd, g(f"{mode}/Vtd"), rtol=0, atol=t["atol"] * 10 * scale)
    np.testing.assert_allclose(Qd, g(f"{mode}/Qd"), rtol=0, atol=t["atol"] * 10 * scale)
    Ed = O.adjoint_backward_pass(g(f"{mode}/E"), g(f"{mode}/Q"), g(f"{mode}/Qd"))
    np.testing.assert_allclose(Ed, g(f"{mode}/Ed"), rtol=0, atol=t["atol"] * 10 * scale)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", ["nw", "sw"])
def test_structure(golden, name, mode):
    """Borders: zeros except Q[N+1,M+1,:]=1 and E[N+1,M+1]=Et; SW freezes the
    first row/column (sw.py:54-55,107-109)."""
    g = lambda k: golden[f"{name}/{k}"]
    theta, A, Et = g("theta"), g("A"), g("Et")
    Vt, Q = O.forward_pass(theta, A, mode)
    E = O.backward_pass(Et, Q, mode)
    B, N, M = theta.shape
    assert (Q[:, N + 1, M + 1, :] == 1).all()
    Qz = Q.copy(); Qz[:, N + 1, M + 1, :] = 0
    assert (Qz[:, 0] == 0).all() and (Qz[:, :, 0] == 0).all()
    assert (Qz[:, N + 1] == 0).all() and (Qz[:, :, M + 1] == 0).all()
    np.testing.assert_array_equal(E[:, N + 1, M + 1], Et)
    np.testing.assert_allclose(E[:, N, M], Et)
    if mode == "sw":
        assert (Q[:, 1] == 0).all() and (Q[:, :, 1] == 0).all()
        assert (E[:, 1] == 0).all() and (E[:, :, 1] == 0).all()
    s = Q[:, 2:N + 1, 2:M + 1].sum(-1)
    np.testing.assert_allclose(s, 1.0, atol=1e-6)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", ["nw", "sw"])
def test_traceback_matches_reference(golden, name, mode):
    grad = golden[f"{name}/{mode}/tb_grad"]
    for variant in ("cpu", "cuda"):
        want = [tuple(r) for r in golden[f"{name}/{mode}/tb_{variant}"].tolist()]
        assert O.traceback(grad, variant) == want


def test_traceback_random(golden, golden_meta):
    for idx in range(golden_meta["n_tb_rand"]):
        grad = golden[f"tb_rand{idx}/grad"]
        for variant in ("cpu", "cuda"):
            want = golden[f"tb_rand{idx}/tb_{variant}"].tolist()
            if want == [[-999, -999, -999]]:
                with pytest.raises(IndexError):
                    O.traceback(grad, variant)
            else:
                assert O.traceback(grad, variant) == [tuple(r) for r in want]


def test_traceback_small_matrices(golden, golden_meta):
    """300 quantised (tie-rich) matrices traced by the reference: Python's negative-index
    wrap-around on rows AND columns and its IndexError cases, both stop rules."""
    shapes = golden_meta["tb_small_shapes"]
    n_err = 0
    for idx, (N, M) in enumerate(shapes):
        grad = golden["tb_small/grad"][idx, :N, :M]
        for variant in ("cpu", "cuda"):
            want = golden[f"tb_small/tb_{variant}"][idx]
            want = want[want[:, 0] != -12345]
            if want.tolist() == [[-999, -999, -999]]:
                n_err += 1
                with pytest.raises(IndexError):
                    O.traceback(grad, variant)
            else:
                assert O.traceback(grad, variant) == [tuple(r) for r in want.tolist()], (idx, variant)
    assert n_err > 10


def test_reference_known_answers(golden, golden_meta):
    """The literal expectations of the reference's own unit tests."""
    ka = golden_meta["known_answers"]
    theta = golden["ref_make_data/theta"]            # fp64, 1x5x4
    A = np.full_like(theta, 0.1)
    for mode, key_cpu, key_cuda in (("nw", "nw_cpu", "nw_cuda_xy"),
                                    ("sw", "sw_cpu", "sw_cuda_xy")):
        Vt, Q, E = O.decode(theta, A, mode)
        np.testing.assert_allclose(Vt, golden[f"ref_make_data/{mode}/Vt"], rtol=1e-14)
        np.testing.assert_allclose(Vt[0], ka[f"survey_anchor_{mode}_Vt"], rtol=1e-12)
        grad = E[0, 1:-1, 1:-1]
        np.testing.assert_allclose(grad, golden[f"ref_make_data/{mode}/grad"][0], atol=1e-14)
        assert O.traceback(grad, "cpu") == [tuple(r) for r in ka[key_cpu]]
        assert O.traceback(grad, "cpu") == [
            tuple(r) for r in golden[f"ref_make_data/{mode}/tb_cpu"].tolist()]
        # fp32 flavour used by test_nw_cuda.py / test_sw_cuda.py
        Vt32, Q32, E32 = O.decode(theta.astype(np.float32), A.astype(np.float32), mode)
        got = O.traceback(E32[0, 1:-1, 1:-1], "cuda")
        assert [list(r[:2]) for r in got] == ka[key_cuda]
        assert got == [tuple(r) for r in golden[f"ref_make_data/{mode}/tb_cuda_f32"].tolist()]


def test_batch_driver_matches_single(golden):
    g = lambda k: golden[f"r33x40/{k}"]
    theta, A = g("theta"), g("A")
    for mode in ("nw", "sw"):
        Vt, Q, E = O.decode(theta, A, mode)
        Vt2, E2 = O.fwd_bwd_batch_f32(theta, A, mode, nthreads=2)
        np.testing.assert_array_equal(Vt, Vt2)
        np.testing.assert_array_equal(E, E2)
    # ragged: each pair on its own (n, m) slice == per-pair oracle (alignment.py:165-169)
    xlen, ylen = np.array([20, 33], np.int32), np.array([40, 7], np.int32)
    Vt3, E3 = O.fwd_bwd_batch_f32(theta, A, "nw", xlen, ylen, nthreads=2)
    for b in range(2):
        n, m = xlen[b], ylen[b]
        v, q, e = O.decode(theta[b:b + 1, :n, :m], A[b:b + 1, :n, :m], "nw")
        assert v[0] == Vt3[b]
        np.testing.assert_array_equal(E3[b, 1:n + 1, 1:m + 1], e[0, 1:-1, 1:-1])
        assert E3[b, n + 1:, :].sum() == 1.0 and E3[b, -1, -1] == 1.0
