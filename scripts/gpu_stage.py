#!/usr/bin/env python
"""Staged GPU bring-up: each micro-check runs in its own subprocess with a short
timeout, so one hung or trapped kernel costs seconds, not the whole box visit.
usage: python scripts/gpu_stage.py [check ...]   (no args = all)"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHECK = r'''
import sys, numpy as np, torch
sys.path.insert(0, %(root)r)
from deepblast_b200 import ops
from oracle import softdp as O
name, mode, B, N, M, W, flags, grid = %(spec)r
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(2)
theta = torch.rand(B, N, M, generator=g); A = -torch.rand(B, N, M, generator=g)
Et = torch.linspace(0.5, 1.5, B)
Zt = torch.randn(B, N + 2, M + 2, generator=g); ZA = torch.randn(B, N, M, generator=g) * 0.1
fl = flags | (W << 4) | (grid << 8)
Vt_o, Q_o = O.forward_pass(theta.numpy(), A.numpy(), mode)
E_o = O.backward_pass(Et.numpy(), Q_o, mode)
res = {}
if name.startswith("fwd") or name.startswith("all"):
    Vt, Q = ops.forward_pass(theta.to(dev), A.to(dev), mode, flags=fl)
    torch.cuda.synchronize()
    res["dVt_rel"] = float(np.abs(Vt.cpu().numpy() - Vt_o).max() / max(1.0, np.abs(Vt_o).max()))
    res["dQ"] = float(np.abs(ops.q_to_reference(Q, N).cpu().numpy() - Q_o).max())
else:
    Q = ops.q_from_reference(torch.from_numpy(Q_o).to(dev))
if name.startswith("bwd") or name.startswith("all"):
    E = ops.backward_pass(Et.to(dev), Q, mode, flags=fl, N=N)
    torch.cuda.synchronize()
    res["dE"] = float(np.abs(E.cpu().numpy() - E_o).max())
if name.startswith("adj") or name.startswith("all"):
    Qr = ops.q_from_reference(torch.from_numpy(Q_o).to(dev))
    Vtd_o, Qd_o = O.adjoint_forward_pass(Q_o, Zt.numpy(), ZA.numpy())
    Ed_o = O.adjoint_backward_pass(E_o, Q_o, Qd_o)
    Vtd, Qd = ops.adjoint_forward_pass(Qr, Zt.to(dev), ZA.to(dev), flags=fl)
    torch.cuda.synchronize()
    sc = max(1.0, float(np.abs(Vtd_o).max()))
    res["dVtd_rel"] = float(np.abs(Vtd.cpu().numpy() - Vtd_o).max() / sc)
    res["dQd_rel"] = float(np.abs(ops.q_to_reference(Qd, N)[:, 1:-1, 1:-1].cpu().numpy() - Qd_o[:, 1:-1, 1:-1]).max() / sc)
    Ed = ops.adjoint_backward_pass(torch.from_numpy(E_o).to(dev), Qr, Qd, flags=fl)
    torch.cuda.synchronize()
    res["dEd_rel"] = float(np.abs(Ed.cpu().numpy() - Ed_o).max() / max(1.0, float(np.abs(Ed_o).max())))
bad = [k for k, v in res.items() if not (v < 1e-4)]
print(("FAIL " if bad else "PASS ") + " ".join("%%s=%%.2e" %% kv for kv in res.items()))
'''

NO_TMA = 2
V1 = 4
SPECS = [
    ("fwd_v2_64", "nw", 2, 64, 64, 1, 0, 0),
    ("bwd_v2_64", "nw", 2, 64, 64, 1, 0, 0),
    ("all_v2_256", "nw", 4, 256, 256, 0, 0, 0),
    ("all_v2_sw_256", "sw", 4, 256, 256, 0, 0, 0),
    ("all_v2_w2_200", "nw", 3, 200, 152, 2, 0, 0),
    ("all_v2_w4_300", "sw", 2, 300, 100, 4, 0, 0),
    ("all_v2_grid", "nw", 13, 70, 92, 2, 0, 3),
    ("all_v2_w1_grid", "sw", 13, 96, 128, 1, 0, 3),
    ("all_v2_1024", "nw", 1, 1024, 1024, 0, 0, 0),
    ("all_v2_520", "nw", 2, 130, 520, 0, 0, 0),
]
SPECS_V1 = [
    # name, mode, B, N, M, W, flags, grid
    ("fwd_tma_64", "nw", 2, 64, 64, 1, V1, 0),
    ("fwd_gen_64", "nw", 2, 64, 64, 1, NO_TMA, 0),
    ("bwd_tma_64", "nw", 2, 64, 64, 1, 0, 0),
    ("bwd_gen_64", "nw", 2, 64, 64, 1, NO_TMA, 0),
    ("adj_tma_64", "nw", 2, 64, 64, 1, 0, 0),
    ("adj_gen_64", "nw", 2, 64, 64, 1, NO_TMA, 0),
    ("all_tma_5x4", "nw", 3, 5, 4, 1, 0, 0),
    ("all_tma_w2_200", "nw", 3, 200, 152, 2, 0, 0),
    ("all_tma_w4_200", "sw", 3, 200, 152, 4, 0, 0),
    ("all_tma_w8_300", "nw", 2, 300, 77, 8, 0, 0),
    ("all_gen_w2_grid", "nw", 13, 70, 90, 2, NO_TMA, 3),
    ("all_tma_w1_grid", "sw", 13, 96, 128, 1, 0, 3),
    ("all_tma_256", "nw", 4, 256, 256, 0, 0, 0),
    ("all_tma_1024", "nw", 1, 1024, 1024, 0, 0, 0),
]


# chained kernels (softdp_fwd3 / softdp_bwd3): forced on small batches through the env
SPECS_V3 = [
    ("fwd_v3_n1_64", "nw", 3, 64, 64, 0, 0, 0),
    ("all_v3_n1_grid", "nw", 13, 96, 128, 0, 0, 3),
    ("all_v3_n2_256", "nw", 5, 256, 256, 0, 0, 0),
    ("all_v3_n2_odd", "nw", 7, 160, 112, 0, 0, 2),
    ("all_v3_n2_ragN", "nw", 4, 100, 96, 0, 0, 0),
    ("all_v3_n1_ragN", "sw", 4, 77, 80, 0, 0, 3),
    ("all_v3_n1_sw", "sw", 6, 128, 64, 0, 0, 4),
    ("all_v3_n2_sw", "sw", 6, 192, 160, 0, 0, 4),
    ("all_v3_n2_r6", "nw", 9, 128, 256, 0, 0, 2),
    ("all_v3_dflt_300", "nw", 300, 64, 64, 0, 0, 0),
    ("all_v3_n2_1024", "nw", 2, 1024, 1024, 0, 0, 0),
]
ENVS = {
    "fwd_v3_n1_64": {"B200DP_NCH": "1"},
    "all_v3_n1_grid": {"B200DP_NCH": "1", "B200DP_RING": "4", "B200DP_BRING": "2", "B200DP_PFW": "64"},
    "all_v3_n2_256": {"B200DP_NCH": "2"},
    "all_v3_n2_odd": {"B200DP_NCH": "2", "B200DP_RING": "3"},
    "all_v3_n2_ragN": {"B200DP_NCH": "2"},
    "all_v3_n1_ragN": {"B200DP_NCH": "1"},
    "all_v3_n1_sw": {"B200DP_NCH": "1"},
    "all_v3_n2_sw": {"B200DP_NCH": "2"},
    "all_v3_n2_r6": {"B200DP_NCH": "2", "B200DP_RING": "6", "B200DP_BRING": "6", "B200DP_PFW": "128", "B200DP_PFD": "12"},
    "all_v3_dflt_300": {},
    "all_v3_n2_1024": {"B200DP_NCH": "2"},
}


def main():
    want = sys.argv[1:]
    for spec in SPECS_V3 + SPECS + SPECS_V1:
        if want and not any(w in spec[0] for w in want):
            continue
        code = CHECK % dict(root=ROOT, spec=spec)
        t0 = time.time()
        try:
            env = dict(os.environ)
            if spec[0] in ENVS:
                env["B200DP_V3MIN"] = "1"
                env.update(ENVS[spec[0]])
                if spec[0] == "all_v3_dflt_300":
                    env.pop("B200DP_V3MIN")
            else:
                env["B200DP_V3"] = "0"
            p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=90, env=env)
            out = (p.stdout.strip().splitlines() or ["(no stdout)"])[-1]
            if p.returncode != 0:
                out = "ERROR rc=%d %s | %s" % (p.returncode, out, p.stderr.strip().splitlines()[-1:] )
        except subprocess.TimeoutExpired:
            out = "TIMEOUT (hang)"
        print("%-18s %5.1fs  %s" % (spec[0], time.time() - t0, out), flush=True)


if __name__ == "__main__":
    main()
