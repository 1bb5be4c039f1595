"""GPU-box stress: random small batches (shapes, per-pair lengths, dense / packed, NW / SW, with and without ZA)
through the strip-queue / cluster kernels and the decoder API, every pass checked against the per-pair oracle.
usage: python scripts/gpu_stress.py [n_cases] [seed]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from deepblast_b200 import ops, plan as P
import test_gpu_sq as T

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
d = torch.device("cuda:0")
fails = 0
for it in range(n_cases):
    B = int(rng.integers(1, 7))
    big = rng.random() < 0.25                         # some cases long enough for many strips / the cluster forward
    N = int(rng.integers(1, 700 if big else 200))
    M = int(rng.integers(1, 160 if big else 50)) * 4                 # dense layout: M % 4 == 0
    mode = "sw" if rng.random() < 0.4 else "nw"
    packed = bool(rng.random() < 0.4)
    ragged = packed or rng.random() < 0.5
    xl = rng.integers(0 if rng.random() < 0.1 else 1, N + 1, B) if ragged else None
    yl = rng.integers(1, M + 1, B) if ragged else None
    theta, A, Zt, ZA = T.rand_batch(B, N, M, seed=int(rng.integers(1 << 30)))
    try:
        plan = P.Plan(B, N, M, xl, yl, packed=packed, device=d)
        T.run_and_check(ops, plan, theta, A, Zt, ZA, mode, use_za=bool(rng.random() < 0.5),
                        vt_rtol=1e-5, vtd_atol=1e-4)     # the two scalars (sums over up to 1e5 cells) at the 1e-4 bar;
        #                                                   per-cell Q / E / Ed keep the tests' 1e-5 / 2e-5 / 1e-4
    except Exception as e:                           # noqa: BLE001
        fails += 1
        print("FAIL case", it, dict(B=B, N=N, M=M, mode=mode, packed=packed, xl=None if xl is None else xl.tolist(),
                                    yl=None if yl is None else yl.tolist()), repr(e)[:300], flush=True)

# ---- the public decoder API (whatever kernels the dispatch picks for the shape): decode + a training-shaped
# double backward, a sample of pairs against the oracle
from deepblast_b200.nw_cuda import NeedlemanWunschDecoder
from deepblast_b200.sw_cuda import SmithWatermanDecoder
n_api = max(10, n_cases // 5)
for it in range(n_api):
    B = int(rng.choice([1, 2, 3, 17, 64, 300, 700, 1500]))
    N = int(rng.integers(1, 70 if B > 64 else 300))
    M = int(rng.integers(1, 70 if B > 64 else 300))
    mode = "sw" if rng.random() < 0.4 else "nw"
    ragged = rng.random() < 0.5
    xl = rng.integers(1, N + 1, B) if ragged else np.full(B, N)
    yl = rng.integers(1, M + 1, B) if ragged else np.full(B, M)
    theta, A, Zt, _ = T.rand_batch(B, N, M, seed=int(rng.integers(1 << 30)))
    try:
        dec = (SmithWatermanDecoder if mode == "sw" else NeedlemanWunschDecoder)('softmax')
        th = theta.to(d).requires_grad_()
        a = A.to(d).requires_grad_()
        aln = dec.decode(th, a, torch.tensor(xl), torch.tensor(yl)) if ragged else dec.decode(th, a)
        (aln * Zt.to(d)).sum().backward()
        for b in rng.choice(B, min(B, 4), replace=False):
            n, m = int(xl[b]), int(yl[b])
            _, _, E_o, _, _, Ed_o = T.oracle_pair(theta[b], A[b], 1.0, Zt[b], None, n, m, mode)
            np.testing.assert_allclose(aln.detach()[b, :n, :m].cpu().numpy(), E_o, rtol=0, atol=2e-5, err_msg="aln")
            sc = max(1.0, float(np.abs(Ed_o).max()))
            np.testing.assert_allclose(th.grad[b, :n, :m].cpu().numpy(), Ed_o, rtol=0, atol=1e-4 * sc, err_msg="grad")
    except Exception as e:                           # noqa: BLE001
        fails += 1
        print("FAIL api case", it, dict(B=B, N=N, M=M, mode=mode, ragged=ragged), repr(e)[:300], flush=True)
print("stress: %d kernel cases + %d API cases, %d failures" % (n_cases, n_api, fails))
sys.exit(1 if fails else 0)
