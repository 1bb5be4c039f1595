import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B200DP_V3MIN"] = "1"
from deepblast_b200 import ops
from oracle import softdp as O
dev = torch.device("cuda:0")
for (B, N, M, G) in [(1, 64, 64, 0), (2, 64, 128, 1), (3, 96, 64, 0)]:
    os.environ["B200DP_CTAS"] = str(G) if G else "0"
    g = torch.Generator().manual_seed(2)
    theta = torch.rand(B, N, M, generator=g); A = -torch.rand(B, N, M, generator=g)
    Et = torch.linspace(0.5, 1.5, B)
    Vt_o, Q_o = O.forward_pass(theta.numpy(), A.numpy(), "nw")
    E_o = O.backward_pass(Et.numpy(), Q_o, "nw")
    Q = ops.q_from_reference(torch.from_numpy(Q_o).to(dev))
    E = ops.backward_pass(Et.to(dev), Q, "nw", N=N).cpu().numpy()
    d = np.abs(E - E_o)
    print("== B,N,M,G", B, N, M, G, "max", d.max(), "nbad", int((d > 1e-4).sum()), "of", d.size)
    for b in range(B):
        bad = np.argwhere(d[b] > 1e-4)
        if len(bad) == 0:
            print(" pair", b, "ok"); continue
        rows = sorted(set(bad[:, 0].tolist())); cols = sorted(set(bad[:, 1].tolist()))
        print(" pair", b, "bad rows", rows[:12], "..", rows[-3:], "n", len(rows), "| bad cols", cols[:12], "..", cols[-3:], "n", len(cols))
        for (i, j) in bad[:6]:
            print("   E[%d,%d] got %.6f want %.6f" % (i, j, E[b, i, j], E_o[b, i, j]))
        # does a shifted version match?
        for di in (-1, 0, 1):
            for dj in (-1, 0, 1):
                if di == 0 and dj == 0: continue
                sh = np.roll(np.roll(E[b], di, 0), dj, 1)
                print("   shift", di, dj, "err", float(np.abs(sh - E_o[b])[2:-2, 2:-2].max()))
