"""Generate golden vectors for the soft-DP path FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference and numba):

    python tests/golden/make_golden.py

Imports deepblast.nw / deepblast.sw / deepblast.nw_cuda / deepblast.sw_cuda
unmodified from /root/reference and records, for a handful of seeded inputs,
every intermediate of the path: Vt, Q (nw.py:65-117), E (nw.py:138-175), the
adjoint pair Vtd/Qd/Ed (nw.py:202-312), the autograd-level results
(decode, double backward; nw.py:315-386) and both traceback variants
(nw.py:401-444, nw_cuda.py:273-317).  The fixtures travel to the GPU box as
tests/golden/softdp_golden.npz; /root/reference does not.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
import deepblast.nw as rnw          # noqa: E402
import deepblast.sw as rsw          # noqa: E402
import deepblast.nw_cuda as rnwc    # noqa: E402
import deepblast.sw_cuda as rswc    # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make_data():
    """theta of deepblast/tests/test_nw.py:10-19 restated without sklearn's
    helper: 1 / (pairwise euclidean distance + 0.1), RandomState(0)."""
    rng = np.random.RandomState(0)
    m, n, k = 2, 1, 3
    Mx = rng.randn(k, 3)
    X = rng.randn(m, 3)
    Y = rng.randn(n, 3)
    X = np.concatenate((X, Mx), axis=0)
    Y = np.concatenate((Mx, Y), axis=0)
    from sklearn.metrics.pairwise import pairwise_distances
    return 1 / (pairwise_distances(X, Y) + 0.1)


def passes(mod, theta, A, Et, Zt, ZA):
    """All four private passes of the reference module `mod` (nw or sw)."""
    B = theta.shape[0]
    Vt, Q = mod._forward_pass(theta, A, 'softmax')
    E = torch.stack([mod._backward_pass(Et[b], Q[b]) for b in range(B)])
    E = E.to(theta.dtype)   # `E[b] = ...` into a theta.dtype tensor, nw.py:347-352
    outs = [mod._adjoint_forward_pass(Q[b], Zt[b], ZA[b], 'softmax')
            for b in range(B)]
    Vtd = torch.stack([o[0] for o in outs]).to(Zt.dtype)
    Qd = torch.stack([o[1] for o in outs]).to(Zt.dtype)   # nw.py:371-382
    Ed = torch.stack([mod._adjoint_backward_pass(E[b], Q[b], Qd[b])
                      for b in range(B)]).to(Zt.dtype)
    return dict(Vt=Vt, Q=Q, E=E, Vtd=Vtd, Qd=Qd, Ed=Ed)


def autograd_level(decoder_cls, theta, A, W):
    """decode + double backward through the reference autograd Functions."""
    theta = theta.clone().requires_grad_()
    A = A.clone().requires_grad_()
    dec = decoder_cls('softmax')
    aln = dec.decode(theta, A)
    loss = (aln * W).sum()
    loss.backward()
    v = dec(theta, A)
    g_theta, g_A = torch.autograd.grad(v.sum(), (theta, A))
    return dict(aln=aln.detach(), theta_grad2=theta.grad.detach(),
                A_grad_is_none=np.array(A.grad is None),
                g_theta=g_theta, g_A=g_A)


def main():
    out = {}
    meta = {"cases": []}
    cases = [  # name, B, N, M, A kind, dtype
        ("t_nw_cuda_5x5", 3, 5, 5, "const-1", torch.float32),
        ("t_nw_4x4", 1, 4, 4, "const-1", torch.float32),
        ("r8x8", 2, 8, 8, "rand", torch.float32),
        ("r17x23", 2, 17, 23, "rand", torch.float32),
        ("r33x40", 2, 33, 40, "rand", torch.float32),
        ("r64x48", 1, 64, 48, "rand", torch.float32),
        ("r40x70", 1, 40, 70, "rand", torch.float32),
        ("r12x9_f64", 2, 12, 9, "rand", torch.float64),
    ]
    for name, B, N, M, akind, dt in cases:
        g = torch.Generator().manual_seed(2)
        theta = torch.rand(B, N, M, generator=g, dtype=dt)
        if akind == "rand":
            A = -torch.rand(B, N, M, generator=g, dtype=dt)
        else:
            A = -torch.ones(B, N, M, dtype=dt)
        Et = torch.rand(B, generator=g, dtype=dt) + 0.5
        Zt = torch.randn(B, N + 2, M + 2, generator=g, dtype=dt)
        ZA = torch.randn(B, N, M, generator=g, dtype=dt) * 0.1
        W = torch.randn(B, N, M, generator=g, dtype=dt)
        for k, v in dict(theta=theta, A=A, Et=Et, Zt=Zt, ZA=ZA, W=W).items():
            out[f"{name}/{k}"] = v.numpy()
        for mode, mod, deccls, cuda_mod in (("nw", rnw, rnw.NeedlemanWunschDecoder, rnwc),
                                            ("sw", rsw, rsw.SmithWatermanDecoder, rswc)):
            r = passes(mod, theta, A, Et, Zt, ZA)
            for k, v in r.items():
                out[f"{name}/{mode}/{k}"] = v.numpy()
            ag = autograd_level(deccls, theta, A, W)
            for k, v in ag.items():
                out[f"{name}/{mode}/ag_{k}"] = v.numpy() if hasattr(v, "numpy") else v
            # tracebacks of the Et = 1 expected-alignment matrix, pair 0
            E1 = torch.stack([mod._backward_pass(torch.tensor(1.0), r["Q"][b])
                              for b in range(B)])
            grad = E1[0, 1:-1, 1:-1].to(dt)
            out[f"{name}/{mode}/tb_grad"] = grad.numpy()
            tb_cpu = deccls('softmax').traceback(grad)
            cdec = (cuda_mod.NeedlemanWunschDecoder if mode == "nw"
                    else cuda_mod.SmithWatermanDecoder)('softmax')
            tb_cuda = cdec.traceback(grad)
            out[f"{name}/{mode}/tb_cpu"] = np.array(tb_cpu, dtype=np.int32)
            out[f"{name}/{mode}/tb_cuda"] = np.array(tb_cuda, dtype=np.int32)
        meta["cases"].append(dict(name=name, B=B, N=N, M=M, A=akind,
                                  dtype=str(dt).replace("torch.", "")))

    # --- the reference tests' own known-answer vectors -------------------
    th = torch.from_numpy(make_data()).unsqueeze(0)          # fp64 5x4
    out["ref_make_data/theta"] = th.numpy()
    for mode, deccls, cuda_mod in (("nw", rnw.NeedlemanWunschDecoder, rnwc),
                                   ("sw", rsw.SmithWatermanDecoder, rswc)):
        theta = th.clone().requires_grad_()
        A = (torch.ones_like(theta) * 0.1).requires_grad_()
        dec = deccls('softmax')
        v = dec(theta, A)
        v.backward()
        out[f"ref_make_data/{mode}/Vt"] = v.detach().numpy()
        out[f"ref_make_data/{mode}/grad"] = theta.grad.numpy()
        out[f"ref_make_data/{mode}/tb_cpu"] = np.array(
            dec.traceback(theta.grad.squeeze()), dtype=np.int32)
        # fp32 variant used by the reference's *_cuda tests (test_nw_cuda.py:64-76)
        theta32 = th.float().clone().requires_grad_()
        A32 = (torch.ones_like(theta32) * 0.1).requires_grad_()
        v32 = dec(theta32, A32)
        v32.backward()
        cdec = (cuda_mod.NeedlemanWunschDecoder if mode == "nw"
                else cuda_mod.SmithWatermanDecoder)('softmax')
        out[f"ref_make_data/{mode}/grad_f32"] = theta32.grad.numpy()
        out[f"ref_make_data/{mode}/Vt_f32"] = v32.detach().numpy()
        out[f"ref_make_data/{mode}/tb_cuda_f32"] = np.array(
            cdec.traceback(theta32.grad.squeeze()), dtype=np.int32)
    # literals copied from the reference's assertions (test_nw.py:51-52,
    # test_sw.py:49-50, test_nw_cuda.py:75, test_sw_cuda.py:69)
    meta["known_answers"] = {
        "nw_cpu": [[0, 0, 0], [1, 0, 0], [2, 0, 1], [3, 1, 1], [4, 2, 2], [4, 3, 1]],
        "sw_cpu": [[-1, 0, 1], [0, 1, 0], [1, 1, 0], [2, 1, 0], [3, 1, 1], [4, 2, 2], [4, 3, 1]],
        "nw_cuda_xy": [[0, 0], [1, 0], [2, 0], [3, 1], [4, 2], [4, 3]],
        "sw_cuda_xy": [[0, 0], [0, 1], [1, 1], [2, 1], [3, 1], [4, 2], [4, 3]],
        "survey_anchor_nw_Vt": 36.84106105685503,
        "survey_anchor_sw_Vt": 24.778630471318223,
    }
    # random-matrix tracebacks (exercise ties, wrap-around, non-square)
    rng = np.random.default_rng(7)
    for idx, (N, M) in enumerate([(6, 9), (9, 6), (25, 23), (1, 7), (7, 1), (1, 1)]):
        gm = rng.random((N, M)).astype(np.float32)
        gm[rng.random((N, M)) < 0.3] = 0.0     # exact ties at 0 like underflowed E
        out[f"tb_rand{idx}/grad"] = gm
        for variant, dec in (("cpu", rnw.NeedlemanWunschDecoder('softmax')),
                             ("cuda", rnwc.NeedlemanWunschDecoder('softmax'))):
            try:
                tb = np.array(dec.traceback(torch.from_numpy(gm)), dtype=np.int32)
            except IndexError:
                tb = np.array([[-999, -999, -999]], dtype=np.int32)
            out[f"tb_rand{idx}/tb_{variant}"] = tb
    meta["n_tb_rand"] = 6
    # many small random matrices with ties and zeros: pins the Python index semantics
    # (negative wrap-around of rows AND columns, IndexError) of both traceback rules
    rng = np.random.default_rng(11)
    grads, tbs = [], {"cpu": [], "cuda": []}
    n_small = 300
    for idx in range(n_small):
        N, M = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        gm = np.round(rng.random((N, M)) * 4).astype(np.float32) / 4     # quantised: many exact ties
        gm[rng.random((N, M)) < 0.35] = 0.0
        pad = np.full((8, 8), np.nan, dtype=np.float32)
        pad[:N, :M] = gm
        grads.append(pad)
        for variant, dec in (("cpu", rnw.NeedlemanWunschDecoder('softmax')),
                             ("cuda", rnwc.NeedlemanWunschDecoder('softmax'))):
            try:
                tb = np.array(dec.traceback(torch.from_numpy(gm)), dtype=np.int32)
            except IndexError:
                tb = np.array([[-999, -999, -999]], dtype=np.int32)
            row = np.full((40, 3), -12345, dtype=np.int32)
            row[:len(tb)] = tb
            tbs[variant].append(row)
        meta.setdefault("tb_small_shapes", []).append([N, M])
    out["tb_small/grad"] = np.stack(grads)
    out["tb_small/tb_cpu"] = np.stack(tbs["cpu"])
    out["tb_small/tb_cuda"] = np.stack(tbs["cuda"])

    np.savez_compressed(os.path.join(HERE, "softdp_golden.npz"), **out)
    with open(os.path.join(HERE, "softdp_golden.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
