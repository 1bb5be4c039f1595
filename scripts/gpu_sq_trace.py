"""GPU-box helper: per-ticket start/end times of one strip-queue launch (b200dp_sq_set_trace), summarised:
how long strips of each ordinal k take, how much of that is waiting, the makespan against the work."""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import ops, plan as P, _lib
from gpu_sq_perf import zipf_lengths

d = torch.device("cuda:0")
REC = np.dtype([('t_off', '<i8'), ('q_off', '<i8'), ('b_in', '<i8'), ('b_out', '<i8'), ('rows', '<i4'), ('m', '<i4'),
                ('pitch', '<i4'), ('pair', '<i4'), ('flags', '<i4'), ('k', '<i4'), ('pad', '<i4', 2)])


def trace(name, pl, which="fwd", flags=0, mode="nw"):
    g = torch.Generator(device=d).manual_seed(2)
    shape = (pl.packed_floats,) if pl.packed else (pl.B, pl.N, pl.M)
    theta = torch.rand(shape, generator=g, device=d)
    A = -torch.rand(shape, generator=g, device=d)
    Et = torch.ones(pl.B, device=d)
    Vt, Q = ops.sq_forward(pl, theta, A, mode, flags=flags)
    ops.sq_backward(pl, Et, Q, mode, flags=flags)
    tr = torch.zeros(2 * pl.nstrips, dtype=torch.int64, device=d)
    torch.cuda.synchronize()
    _lib.lib().b200dp_sq_set_trace(tr.data_ptr())
    if which == "fwd":
        ops.sq_forward(pl, theta, A, mode, flags=flags)
    else:
        ops.sq_backward(pl, Et, Q, mode, flags=flags)
    torch.cuda.synchronize()
    _lib.lib().b200dp_sq_set_trace(None)
    t = tr.cpu().numpy().reshape(-1, 2).astype(np.float64)
    tab = pl.tabs_host[0 if which == "fwd" else 1].view(REC).reshape(-1)[:pl.nstrips]
    t0 = t[:, 0].min()
    start, end = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3       # us
    dur = end - start
    steps = tab['m'] + 31
    print(f"== {name} {which} flags {flags:#x}: makespan {end.max():.1f} us, strips {pl.nstrips}, "
          f"sum(dur) {dur.sum() / 1e3:.2f} ms, ns/step mean {1e3 * dur.sum() / steps.sum():.1f}")
    # order of a strip inside its pair in processing direction
    kk = tab['k'] if which == "fwd" else None
    if which != "fwd":
        Kp = {}
        for r in tab:
            Kp[r['pair']] = max(Kp.get(r['pair'], 0), r['k'] + 1)
        kk = np.array([Kp[r['pair']] - 1 - r['k'] for r in tab])
    for k in sorted(set(kk.tolist()))[:40]:
        sel = kk == k
        print(f"   k {k:2d}: n {sel.sum():5d}  start {start[sel].min():8.1f}..{start[sel].max():8.1f}  end max {end[sel].max():8.1f} "
              f" ns/step {1e3 * dur[sel].sum() / steps[sel].sum():7.1f}")
    # the longest pair: timeline of its strips
    big = tab['pair'][np.argmax(tab['m'].astype(np.int64) * 1000 + kk)]
    sel = np.where(tab['pair'] == big)[0]
    sel = sel[np.argsort(kk[sel])]
    print("   longest pair", int(big), "m", int(tab['m'][sel[0]]), "strips", len(sel))
    print("   start:", " ".join(f"{start[i]:.0f}" for i in sel[:33]))
    print("   end:  ", " ".join(f"{end[i]:.0f}" for i in sel[:33]))
    np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"trace_{name}_{which}.npy"),
            np.stack([start, end, kk, tab['m'], tab['pair']], 1))


if __name__ == "__main__":
    xl, yl = zipf_lengths(1024, np.random.default_rng(0))
    for fl in [int(x, 0) for x in os.environ.get("SQ_FLAGS", "0").split(",")]:
        for w in sys.argv[1:] or ["c2", "c5p", "b32"]:
            if w == "c2":
                trace("c2", P.Plan(1024, 256, 256, device=d), "fwd", fl)
                trace("c2", P.Plan(1024, 256, 256, device=d), "bwd", fl)
            if w == "c5p":
                trace("c5p", P.Plan(1024, 1024, 1024, xl, yl, packed=True, device=d), "fwd", fl)
                trace("c5p", P.Plan(1024, 1024, 1024, xl, yl, packed=True, device=d), "bwd", fl)
            if w == "b32":
                trace("b32", P.Plan(32, 1024, 1024, device=d), "fwd", fl)
                trace("b32", P.Plan(32, 1024, 1024, device=d), "bwd", fl)
