"""GPU-box helper: strip-queue kernels against the round-1 kernels (chained / hand-off), kernel-only
timing with CUDA events: C2, C3, C4 slice, C5 (dense-ragged and packed), small batches of long pairs."""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops, plan as P

d = torch.device("cuda:0")


def timeit(fn, it=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def zipf_lengths(B, rng):
    k = np.arange(1, 17)
    pk = (1.0 / k) / (1.0 / k).sum()
    return 64 * rng.choice(k, size=B, p=pk), 64 * rng.choice(k, size=B, p=pk)


def bench(name, B, N, M, mode="nw", xl=None, yl=None, packed=False, legacy=True, flags=0, adj=True):
    g = torch.Generator(device=d).manual_seed(2)
    pl = P.Plan(B, N, M, xl, yl, packed=packed, device=d)
    shape = (pl.packed_floats,) if packed else (B, N, M)
    theta = torch.rand(shape, generator=g, device=d)
    A = -torch.rand(shape, generator=g, device=d)
    Zt = torch.randn(shape, generator=g, device=d)
    Et = torch.ones(B, device=d)
    cells = pl.cells
    out = {"name": name, "cells": cells, "nstrips": pl.nstrips}
    st = {}

    def f():
        st["Vt"], st["Q"] = ops.sq_forward(pl, theta, A, mode, flags=flags)
    def fs():
        ops.sq_forward(pl, theta, A, mode, need_q=False, flags=flags)
    def b():
        st["E"] = ops.sq_backward(pl, Et, st["Q"], mode, flags=flags)
    def af():
        st["Vtd"], st["QdE"] = ops.sq_adjoint_forward(pl, st["Q"], Zt, None, st["E"], flags=flags)
    def ab():
        st["Ed"] = ops.sq_adjoint_backward(pl, st["Q"], st["QdE"], flags=flags)
    tf, tfs, tb = timeit(f), timeit(fs), timeit(b)
    out["sq"] = {"fwd_ms": tf, "score_ms": tfs, "bwd_ms": tb, "G": cells / (tf + tb) / 1e6,
                 "fwd_GBps_moved": cells * 16 / tf / 1e6, "bwd_GBps_moved": cells * 12 / tb / 1e6}
    if adj:
        taf, tab_ = timeit(af), timeit(ab)
        out["sq"].update({"adj_fwd_ms": taf, "adj_bwd_ms": tab_})
    if legacy and not packed:
        xld = None if xl is None else torch.tensor(xl, dtype=torch.int32, device=d)
        yld = None if yl is None else torch.tensor(yl, dtype=torch.int32, device=d)
        def lf():
            st["lVt"], st["lQ"] = ops.forward_pass(theta, A, mode, xld, yld)
        def lb():
            st["lE"] = ops.backward_pass(Et, st["lQ"], mode, xld, yld, N=N)
        tlf, tlb = timeit(lf), timeit(lb)
        out["legacy"] = {"fwd_ms": tlf, "bwd_ms": tlb, "G": cells / (tlf + tlb) / 1e6}
        if xl is None:
            err = float((st["E"] - st["lE"][:, 1:-1, 1:-1]).abs().max())
            out["max_abs_E_diff"] = err
    print(json.dumps(out), flush=True)
    del st, theta, A, Zt
    torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c3", "c4", "c5", "c5p", "small"]
    fl = int(os.environ.get("SQ_FLAGS", "0"), 0)
    if "c2" in which:
        bench("c2 1024x256x256 nw", 1024, 256, 256, flags=fl)
    if "c3" in which:
        bench("c3 1024x256x256 sw", 1024, 256, 256, "sw", flags=fl)
    if "c4" in which:
        bench("c4 1024x512x512 nw", 1024, 512, 512, flags=fl)
    xl, yl = zipf_lengths(1024, np.random.default_rng(0))
    if "c5" in which:
        bench("c5 dense-ragged 1024 zipf", 1024, 1024, 1024, xl=xl, yl=yl, flags=fl)
    if "c5p" in which:
        bench("c5 packed 1024 zipf", 1024, 1024, 1024, xl=xl, yl=yl, packed=True, flags=fl)
    if "small" in which:
        bench("32x1024x1024", 32, 1024, 1024, flags=fl)
        bench("32x512x512", 32, 512, 512, flags=fl)
        bench("64x256x256", 64, 256, 256, flags=fl)
        bench("256x256x256", 256, 256, 256, flags=fl)
        bench("512x1024x1024", 512, 1024, 1024, flags=fl, adj=False)
