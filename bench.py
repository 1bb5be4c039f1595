#!/usr/bin/env python
"""bench.py -- DP cell-updates/s (forward + backward) of the batched soft-DP alignment
path on B200, next to the CPU oracle port timed on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--no-extras]
    python bench.py --impl reference ...      # CPU arm (oracle port, all host threads)
    torchrun --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch: NeedlemanWunschFunction.apply
(theta, A -> Vt) then Vt.sum().backward() (-> dVt/dtheta), i.e. the reference's
`_forward_pass_kernel` + `_backward_pass_kernel` (deepblast/nw_cuda.py:46-102) behind
the reference's autograd API.  Metric and byte accounting: SURVEY.md section 8(d):
36 algorithmic bytes per cell-update (fwd: theta 4 + A 4 read, Q 12 written;
bwd: Q 12 read, E 4 written).

The JSON line's headline (`value`, `roofline`, `e2e`, `cpu_baseline`) is the workload named
by --workload (default: BASELINE configs[1], C2); unless --no-extras the same line carries
`workloads`: every other BASELINE config (c3, c4 slice, c5 packed ragged), a small batch of
long pairs and two training-shaped steps, each timed the same way under the same clock.

The forward and the backward of a step are captured once in two CUDA graphs (through the public
autograd API, static input buffers) and replayed: the timed region holds K x {fwd graph, bwd
graph} with CUDA events between them, so the per-kernel times are device times and the host
only enqueues two graph launches per step.  `eager` reports the same K steps issued call by call.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_FWD, BYTES_BWD = 20, 16
METRIC = "DP cell-updates/sec (fwd+bwd), batched NW soft-align"

WORKLOADS = {
    # name: (mode, pairs per GPU, N, M, description)
    "c2": ("nw", 1024, 256, 256, "BASELINE configs[1]: batch=1024 pairs 256x256, NW forward+backward fp32"),
    "c3": ("sw", 1024, 256, 256, "BASELINE configs[2]: batch=1024 pairs 256x256, Smith-Waterman soft-DP"),
    "c4": ("nw", 1024, 512, 512, "BASELINE configs[3]: 8192 pairs 512x512 batch-sharded 8 ways = 1024 pairs/GPU"),
    "c5": ("nw", 1024, 1024, 1024, "BASELINE configs[4]: variable lengths 64..1024 (Zipf over 16 buckets), 1024 pairs/GPU, "
                                   "PACKED layout (no padding to the longest pair)"),
    "b32": ("nw", 32, 1024, 1024, "small batch of long pairs: 32 pairs 1024x1024 (trainer.py:375 batch size, max_len 1024)"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def zipf_lengths(B, rng):
    k = np.arange(1, 17)
    pk = (1.0 / k) / (1.0 / k).sum()
    return 64 * rng.choice(k, size=B, p=pk), 64 * rng.choice(k, size=B, p=pk)


_ALL_CPUS = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None


def bind_to_gpu_numa_node(index):
    """Pin this process (and so the pinned host buffers it first-touches) to the CPU cores
    NVML reports as local to GPU `index`: host<->device copies then stay on the GPU's own
    socket instead of crossing the inter-socket link.  Returns a short description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} GPU-local cores"
        return "no GPU-local cores reported"
    except Exception as e:      # no NVML / not permitted: run unbound
        return f"unbound ({type(e).__name__})"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port_throughput(mode, N, M, xlen=None, ylen=None, target_s=12.0, seed=2):
    """Oracle port (oracle/softdp_oracle.c, fp64, OpenMP over pairs) on a bounded sample
    of the same workload.  Returns (cells/s, cores, sample description, seconds)."""
    from oracle import softdp as O
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    rng = np.random.default_rng(seed)

    def make(Bs):
        theta = rng.random((Bs, N, M), dtype=np.float32)
        A = -rng.random((Bs, N, M), dtype=np.float32)
        xl = None if xlen is None else np.resize(xlen, Bs).astype(np.int32)
        yl = None if ylen is None else np.resize(ylen, Bs).astype(np.int32)
        cells = Bs * N * M if xl is None else int((xl.astype(np.int64) * yl).sum())
        if mode == "sw" and xl is None:
            cells = Bs * (N - 1) * (M - 1)
        return theta, A, xl, yl, cells

    def run(batch):
        theta, A, xl, yl, cells = batch
        t0 = time.perf_counter()
        O.fwd_bwd_batch_f32(theta, A, mode, xl, yl, nthreads=cores, want_E=True)
        return cells, time.perf_counter() - t0

    run(make(max(cores, 8)))                             # warm the .so and the thread pool
    cells, dt = run(make(4 * max(cores, 8)))             # calibration
    rate = cells / dt
    per_pair = (N * M) if xlen is None else float(np.mean(np.asarray(xlen, np.int64) * np.asarray(ylen)))
    # a sample of about target_s seconds: one batch of at most 1024 pairs (generated once,
    # outside the timed part), swept repeatedly
    want = max(cores, target_s * rate / per_pair)
    Bs = int(min(1024, want))
    Bs = max(cores, (Bs // cores) * cores)
    reps = max(1, int(round(want / Bs)))
    batch = make(Bs)
    cells = 0
    dt = 0.0
    for _ in range(reps):
        c1, d1 = run(batch)
        cells += c1
        dt += d1
    return cells / dt, cores, f"{reps} x {Bs} pairs of the workload ({cells} cells), fwd+bwd, {dt:.1f} s", dt


def reference_arm(args):
    """--impl reference: the CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    mode, Bg, N, M, desc = WORKLOADS[args.workload]
    xlen = ylen = None
    if args.workload == "c5":
        xlen, ylen = zipf_lengths(Bg, np.random.default_rng(0))
    vals, times = [], []
    cores, sample = 1, ""
    # every step is a bounded sample of the workload; the whole run stays within a few minutes
    per_step_s = min(args.cpu_seconds, max(2.0, 150.0 / max(1, args.warmup + args.steps)))
    for it in range(args.warmup + args.steps):
        v, cores, sample, dt = cpu_port_throughput(mode, N, M, xlen, ylen, target_s=per_step_s, seed=2 + it)
        if it >= args.warmup:
            vals.append(v); times.append(dt)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "cell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "mode": mode, "pairs_per_gpu": Bg, "N": N, "M": M,
                   "global_pairs": Bg * args.gpus,
                   "parallelism": f"batch-sliced x{args.gpus}, no DP collective",
                   "note": "the reference is Python+Numba (not compilable to oracle/_ref); this arm times the C oracle "
                           "port of deepblast/nw.py with OpenMP over pairs on a bounded sample per step"},
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
class Ctx:
    pass


def make_workload(cx, name, pairs_override=0):
    """Synthetic inputs of one workload on this rank: theta, A (dense or packed), plan, cells."""
    import torch
    from deepblast_b200 import plan as P
    mode, Bg, N, M, desc = WORKLOADS[name]
    if pairs_override:
        Bg = pairs_override
    w = Ctx()
    w.name, w.mode, w.N, w.M, w.desc = name, mode, N, M, desc
    gen = torch.Generator(device=cx.dev).manual_seed(2 + cx.rank)
    w.plan = None
    w.stats = {}
    if name == "c5":
        from deepblast_b200.sharding import lpt_assign, packing_stats
        xl_all, yl_all = zipf_lengths(Bg * cx.world, np.random.default_rng(0))
        asg = lpt_assign(xl_all * yl_all, cx.world)
        w.stats = packing_stats(xl_all, yl_all, asg)
        mine = asg[cx.rank]
        w.xl, w.yl = xl_all[mine].astype(np.int32), yl_all[mine].astype(np.int32)
        Bg = len(mine)
        w.plan = P.get_plan(Bg, int(w.xl.max()), int(w.yl.max()), w.xl, w.yl, True, cx.dev)
        w.stats["padded_cells_if_dense"] = int(Bg * 1024 * 1024)
        w.stats["packed_floats"] = int(w.plan.packed_floats)
        shape = (w.plan.packed_floats,)
        w.cells = int(w.plan.cells)
    else:
        w.xl = w.yl = None
        shape = (Bg, N, M)
        w.cells = Bg * (N - 1) * (M - 1) if mode == "sw" else Bg * N * M
    w.B = Bg
    w.theta = torch.rand(shape, generator=gen, device=cx.dev).requires_grad_()
    w.A = -torch.rand(shape, generator=gen, device=cx.dev)
    return w


def step_fns(cx, w):
    """(forward, backward) closures through the public autograd API."""
    import torch
    from deepblast_b200.nw_cuda import NeedlemanWunschFunction
    from deepblast_b200.sw_cuda import SmithWatermanFunction
    Fn = NeedlemanWunschFunction if w.mode == "nw" else SmithWatermanFunction
    st = {}

    def fwd():
        if w.plan is None:
            st["Vt"] = Fn.apply(w.theta, w.A, 'softmax')
        else:
            st["Vt"] = Fn.apply(w.theta, w.A, 'softmax', None, None, w.plan)

    def bwd():
        st["g"], = torch.autograd.grad(st["Vt"].sum(), w.theta)
    return fwd, bwd, st


def time_workload(cx, w, steps, warmup, use_graph=True):
    """W warm-up steps, then exactly `steps` timed steps bracketed by barrier + synchronize; CUDA
    events between the forward and the backward of every step.  Returns a dict of timings."""
    import torch
    import torch.distributed as dist
    fwd, bwd, st = step_fns(cx, w)

    def sync_all():
        if cx.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    res = {}
    # ---- eager: call by call (also the warm-up of every lazy initialisation) ------------------
    for it in range(max(warmup, 3)):
        fwd(); bwd()
        if it == 0:
            sync_all()
    sync_all()
    n_e = max(3, min(steps, 10))
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_e)]
    t0 = time.perf_counter()
    for it in range(n_e):
        ev[it][0].record(); fwd(); ev[it][1].record(); bwd(); ev[it][2].record()
    host_eager = (time.perf_counter() - t0) * 1e3 / n_e
    torch.cuda.synchronize()
    res["eager"] = {"ms_per_step": ev[0][0].elapsed_time(ev[-1][2]) / n_e,
                    "host_enqueue_ms_per_step": host_eager, "steps": n_e}
    # ---- graphs: forward and backward captured once, replayed ------------------------------------
    run_f, run_b, graphed = fwd, bwd, False
    for attempt in range(2 if use_graph else 0):
        try:
            # (the whole benchmark runs on cx.side, a non-default stream: a leaf first used on the
            # legacy default stream would make its backward synchronise with that stream, which
            # a capture cannot express)
            side = torch.cuda.current_stream(cx.dev)
            st.clear()
            for _ in range(3):
                fwd(); bwd()
            torch.cuda.synchronize()
            gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(gf, stream=side):
                fwd()
            with torch.cuda.graph(gb, pool=gf.pool(), stream=side):
                bwd()
            run_f, run_b, graphed = gf.replay, gb.replay, True
            res.pop("graph_error", None)
            break
        except Exception as e:                           # capture not possible here: stay eager, say so
            import traceback
            traceback.print_exc(file=sys.stderr)
            res["graph_error"] = f"attempt {attempt}: {type(e).__name__}: {str(e)[:160]}"
            try:
                torch.cuda.synchronize()
            except Exception:
                pass
    for _ in range(max(warmup, 3)):
        run_f(); run_b()
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    sync_all()
    host_t0 = time.perf_counter()
    ev0.record()
    for it in range(steps):
        kev[it][0].record(); run_f(); kev[it][1].record(); run_b(); kev[it][2].record()
    ev1.record()
    host_ms = (time.perf_counter() - host_t0) * 1e3 / steps
    torch.cuda.synchronize()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=cx.dev)
    cells = torch.tensor([float(w.cells)], dtype=torch.float64, device=cx.dev)
    if cx.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cells, op=dist.ReduceOp.SUM)
    res.update({
        "ms_total": float(t.item()), "total_cells": float(cells.item()), "graphed": graphed,
        "fwd_ms": float(np.mean([k[0].elapsed_time(k[1]) for k in kev])),
        "bwd_ms": float(np.mean([k[1].elapsed_time(k[2]) for k in kev])),
        "host_enqueue_ms_per_step": host_ms,
        "step_ms_median": float(np.median([k[0].elapsed_time(k[2]) for k in kev])),
        "step_ms_max": float(np.max([k[0].elapsed_time(k[2]) for k in kev])),
        "step_ms_first3": [round(k[0].elapsed_time(k[2]), 4) for k in kev[:3]],
    })
    # keep the captured graphs (and their pool) alive until the numbers are out
    res["_keep"] = (run_f, run_b, st)
    return res


def kernel_names(cx, w):
    """Which kernels the dispatch picks for this workload (deepblast_b200/ops.py route_plan,
    _functions.py): forward, backward."""
    sms = cx.sms
    if w.plan is not None:
        return "softdp_sq_fwd_kernel", "softdp_sq_bwd_kernel"
    chained = w.B >= 4 * sms and w.N >= 32 and w.M >= 64 and w.M % 32 == 0
    if chained:
        return "softdp_fwd3_kernel", "softdp_sq_bwd_kernel"
    if w.B <= 64 or w.N >= 512:
        return "softdp_sq_fwd_kernel", "softdp_sq_bwd_kernel"
    return "softdp_fwd2_kernel", "softdp_bwd2_kernel"


def roofline(cx, w, r, traffic_tab):
    peak, peak_src = peaks()
    fwd_gbs = w.cells * BYTES_FWD / (r["fwd_ms"] * 1e-3) / 1e9
    bwd_gbs = w.cells * BYTES_BWD / (r["bwd_ms"] * 1e-3) / 1e9
    dom = "fwd" if r["fwd_ms"] >= r["bwd_ms"] else "bwd"
    ach = fwd_gbs if dom == "fwd" else bwd_gbs
    kf, kb = kernel_names(cx, w)
    traffic = traffic_tab.get(f"{w.name}_{dom}")
    dom_ms = r["fwd_ms"] if dom == "fwd" else r["bwd_ms"]
    return {"bound": "hbm", "kernel": kf if dom == "fwd" else kb, "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
            # measured DRAM bytes of that launch (ncu, profiles/traffic.json) over the live duration: the
            # kernels store two of the three Q states, so they move fewer bytes than the 36 B/cell contract
            "traffic_GBps": (traffic / (dom_ms * 1e-3) / 1e9) if traffic else None,
            "traffic_frac": (traffic / (dom_ms * 1e-3) / 1e9 / peak) if traffic else None,
            "algorithmic_bytes_per_cell": BYTES_FWD if dom == "fwd" else BYTES_BWD,
            "fwd": {"kernel": kf, "ms": r["fwd_ms"], "GBps": fwd_gbs, "frac": fwd_gbs / peak},
            "bwd": {"kernel": kb, "ms": r["bwd_ms"], "GBps": bwd_gbs, "frac": bwd_gbs / peak}}


def host_ceiling(cx, h2d_bytes, d2h_bytes, reps=5):
    """What the box's host link allows for one step's copies alone: plain pinned cudaMemcpyAsync of
    the same byte counts, uploads and downloads on two streams at once, every rank at the same
    time (barrier first).  ms per step on this rank (max over ranks)."""
    import torch
    import torch.distributed as dist
    hb = torch.empty(max(h2d_bytes, 1), dtype=torch.uint8, pin_memory=True)
    hd = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, pin_memory=True)
    db = torch.empty(max(h2d_bytes, 1), dtype=torch.uint8, device=cx.dev)
    dd = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, device=cx.dev)
    s1, s2 = torch.cuda.Stream(device=cx.dev), torch.cuda.Stream(device=cx.dev)

    def once():
        with torch.cuda.stream(s1):
            db.copy_(hb, non_blocking=True)
        with torch.cuda.stream(s2):
            hd.copy_(dd, non_blocking=True)
    once()
    torch.cuda.synchronize()
    if cx.world > 1:
        dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(reps):
        once()
    e1.record(s1); e2.record(s2)
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(e1), e0.elapsed_time(e2)) / reps
    t = torch.tensor([ms], dtype=torch.float64, device=cx.dev)
    if cx.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def e2e_dense(cx, w, nst):
    """Host buffers in, host results out, every step, through Decoder.decode_host -> C ABI
    b200dp_decode_host (chunked upload | fwd+bwd | download on three streams)."""
    import torch
    import torch.distributed as dist
    from deepblast_b200 import ops
    B, N, M = w.B, w.N, w.M
    h_theta = torch.empty((B, N, M), dtype=torch.float32, pin_memory=True).copy_(w.theta.detach().cpu())
    h_A = torch.empty((B, N, M), dtype=torch.float32, pin_memory=True).copy_(w.A.cpu())
    h_Vt = torch.empty(B, dtype=torch.float32, pin_memory=True)
    h_E = torch.empty((B, N + 2, M + 2), dtype=torch.float32, pin_memory=True)

    def step():
        ops.decode_host_async(h_theta, h_A, w.mode, out=(h_Vt, h_E), device=cx.dev)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if cx.world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(nst):
        step()
    e1.record()
    torch.cuda.synchronize()
    te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=cx.dev)
    if cx.world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    h2d, d2h = int(2 * B * N * M * 4), int(B * (N + 2) * (M + 2) * 4 + B * 4)
    return float(te.item()) / nst, h2d, d2h, (
        "pinned host theta/A -> b200dp_decode_host (chunked H2D | fwd+bwd | D2H pipeline, 3 streams) -> "
        "pinned host Vt and padded E (dVt/dtheta = E[:,1:-1,1:-1]), per rank")


def e2e_packed(cx, w, nst):
    """Ragged batch, host buffers in the PACKED layout: only the useful cells cross PCIe.  The
    batch is cut into chunks of pairs; upload, sweeps and download of consecutive chunks overlap
    (deepblast_b200.packed.PackedHostDecoder)."""
    import torch
    import torch.distributed as dist
    from deepblast_b200.packed import PackedHostDecoder
    dec = PackedHostDecoder(w.xl, w.yl, w.mode, device=cx.dev)
    h_theta = torch.empty(dec.packed_floats, dtype=torch.float32, pin_memory=True).uniform_()
    h_A = torch.empty(dec.packed_floats, dtype=torch.float32, pin_memory=True).uniform_().neg_()
    for _ in range(2):
        dec.decode(h_theta, h_A)
    torch.cuda.synchronize()
    if cx.world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(nst):
        dec.decode(h_theta, h_A)
    e1.record()
    torch.cuda.synchronize()
    te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=cx.dev)
    if cx.world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    h2d, d2h = int(2 * dec.packed_floats * 4), int(dec.packed_floats * 4 + w.B * 4)
    return float(te.item()) / nst, h2d, d2h, (
        f"pinned host PACKED theta/A -> {dec.nchunks} chunks: H2D | strip-queue fwd+bwd | D2H on three streams -> "
        "pinned host Vt and packed dVt/dtheta, per rank")


def e2e_infer(cx, w, nst):
    """Inference from host buffers: what DeepBLAST.align needs per pair is the alignment PATH
    (trainer.py:80-88), so the walk runs on the device and only the paths and scores come back
    (deepblast_b200.align.HostAligner)."""
    import torch
    import torch.distributed as dist
    from deepblast_b200.align import HostAligner
    B, N, M = w.B, w.N, w.M
    h_theta = torch.empty((B, N, M), dtype=torch.float32, pin_memory=True).copy_(w.theta.detach().cpu())
    h_A = torch.empty((B, N, M), dtype=torch.float32, pin_memory=True).copy_(w.A.cpu())
    al = HostAligner(B, N, M, w.mode, device=cx.dev)
    for _ in range(2):
        al.align(h_theta, h_A)
    torch.cuda.synchronize()
    if cx.world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(nst):
        al.align(h_theta, h_A)
    e1.record()
    torch.cuda.synchronize()
    te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=cx.dev)
    if cx.world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    h2d = int(2 * B * N * M * 4)
    d2h = int(al.paths_h.numel() * 4 + B * 8)
    return float(te.item()) / nst, h2d, d2h


def train_step_bench(cx, B, N, M, steps):
    """Training-shaped step (trainer.py:154-188): decode (forward + backward, create_graph) ->
    MatrixCrossEntropy -> loss.backward() (the adjoint pair).  Eager, and replayed from one CUDA graph."""
    import torch
    from deepblast_b200.nw_cuda import NeedlemanWunschDecoder
    from deepblast_b200.losses import MatrixCrossEntropy
    g = torch.Generator(device=cx.dev).manual_seed(5)
    theta = torch.rand(B, N, M, generator=g, device=cx.dev).requires_grad_()
    A = (-torch.rand(B, N, M, generator=g, device=cx.dev)).requires_grad_()
    Ytrue = (torch.rand(B, N, M, generator=g, device=cx.dev) < 0.01).float()
    G = torch.ones(B, N, M, device=cx.dev)
    xl = torch.full((B,), N, dtype=torch.int32, device=cx.dev)
    yl = torch.full((B,), M, dtype=torch.int32, device=cx.dev)
    dec, crit = NeedlemanWunschDecoder('softmax'), MatrixCrossEntropy()
    st = {}

    def step():
        theta.grad = None
        aln = dec.decode(theta, A)
        loss = crit(Ytrue, aln, xl, yl, G)
        loss.backward()
        st["loss"] = loss

    def timeit(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        host = (time.perf_counter() - t0) * 1e3 / n
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, host
    out = {"B": B, "N": N, "M": M, "cells": B * N * M}
    ms, host = timeit(step, steps)
    out["eager_ms"], out["eager_host_ms"] = ms, host
    try:
        side = torch.cuda.current_stream(cx.dev)
        st.clear()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        theta.grad = None
        with torch.cuda.graph(gr, stream=side):
            aln = dec.decode(theta, A)
            loss = crit(Ytrue, aln, xl, yl, G)
            gth, = torch.autograd.grad(loss, theta)
        ms, host = timeit(gr.replay, steps)
        out["graph_ms"], out["graph_host_ms"] = ms, host
        out["_keep"] = (gr, gth)
    except Exception as e:
        out["graph_error"] = f"{type(e).__name__}: {str(e)[:160]}"
        torch.cuda.synchronize()
    best = out.get("graph_ms", out["eager_ms"])
    out["cell_updates_per_s_4_sweeps"] = B * N * M / (best * 1e-3)
    return out


def producer_bench(cx, B, L, D, steps):
    """The step before the DP (alignment.py:122-123): theta / A from the embeddings, fused tcgen05
    GEMM + activation (deepblast_b200.producer) next to torch's fp32 einsum + softplus / logsigmoid."""
    import torch
    import torch.nn.functional as F
    from deepblast_b200 import producer
    g = torch.Generator(device=cx.dev).manual_seed(7)
    zs = [torch.randn(B, L, D, generator=g, device=cx.dev) * (1.1 / D ** 0.25) for _ in range(4)]

    def timeit(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def ref():
        return (F.softplus(torch.einsum('bid,bjd->bij', zs[0], zs[1])),
                F.logsigmoid(torch.einsum('bid,bjd->bij', zs[2], zs[3])))
    ours = timeit(lambda: producer.theta_a(*zs), steps)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        t_ref = timeit(ref, max(2, steps // 2))
        th_r, a_r = ref()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    th, a = producer.theta_a(*zs)
    flop = 2 * 2.0 * B * L * L * D                      # both products, useful
    peak = 1656.8
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            peak = float(json.load(f).get("bf16_tflops", peak))
    # algorithmic bytes: the four fp32 embedding tensors read once, theta and A written once.  At
    # D = 1024, L = 256 that is 170 issued bf16 FLOP per byte against a machine balance of ~250: the
    # producer is bound by reading its inputs, not by the tensor cores
    nbytes = 4.0 * B * L * D * 4 + 2.0 * B * L * L * 4
    hbm = peaks()[0]
    return {"B": B, "L": L, "D": D, "ms": ours, "torch_fp32_ms": t_ref,
            "TFLOPs_useful": flop / ours / 1e9, "TFLOPs_issued_bf16": 3 * flop / ours / 1e9,
            "roofline": {"bound": "hbm", "achieved": nbytes / ours / 1e6, "peak": hbm, "unit": "GB/s",
                         "frac": nbytes / ours / 1e6 / hbm, "traffic": None,
                         "tensor": {"achieved": 3 * flop / ours / 1e9, "peak": peak, "unit": "TFLOP/s",
                                    "frac": 3 * flop / ours / 1e9 / peak},
                         "note": "algorithmic bytes (embeddings read once, theta / A written once) over the measured copy "
                                 "peak; tensor: issued bf16 FLOPs (three passes of the hi/lo split, done inside the GEMM "
                                 "kernel) over the measured cuBLAS bf16 burst peak"},
            "max_abs_err_theta_vs_torch_fp32": float((th - th_r).abs().max()),
            "max_abs_err_A_vs_torch_fp32": float((a - a_r).abs().max())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs-per-gpu", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline workload")
    ap.add_argument("--no-graph", action="store_true", help="time eager calls only")
    ap.add_argument("--no-bind", action="store_true", help="do not bind the process to the GPU-local CPU cores")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: deepblast_b200 has no CPU path")
    cx = Ctx()
    cx.world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cx.dev = torch.device("cuda", local)
    cx.sms = torch.cuda.get_device_properties(cx.dev).multi_processor_count
    affinity = bind_to_gpu_numa_node(local) if not args.no_bind else "unbound (--no-bind)"
    if cx.world > 1:
        dist.init_process_group("nccl", device_id=cx.dev)
    if cx.world != args.gpus and cx.rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={cx.world}", file=sys.stderr)
    cx.side = torch.cuda.Stream(device=cx.dev)
    torch.cuda.set_stream(cx.side)                     # everything below: one non-default stream
    traffic_tab = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic_tab = json.load(f)
    steps, warmup = args.steps, max(args.warmup, 3)

    # ---- headline ---------------------------------------------------------------------------------
    w = make_workload(cx, args.workload, args.pairs_per_gpu)
    sampler = ClockSampler(local)
    if cx.rank == 0:
        sampler.start()
        time.sleep(0.3)
    r = time_workload(cx, w, steps, warmup, use_graph=not args.no_graph)
    clocks = sampler.stop() if cx.rank == 0 else None
    value = r["total_cells"] * steps / (r["ms_total"] * 1e-3)

    def e2e_of(wk, nst):
        ms, h2d, d2h, note = (e2e_packed if wk.plan is not None and wk.plan.packed else e2e_dense)(cx, wk, nst)
        cells = torch.tensor([float(wk.cells)], dtype=torch.float64, device=cx.dev)
        if cx.world > 1:
            dist.all_reduce(cells, op=dist.ReduceOp.SUM)
        ceil_ms = host_ceiling(cx, h2d, d2h)
        return {"value": float(cells.item()) / (ms * 1e-3), "unit": "cell-updates/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms, "steps": nst, "host_affinity": affinity,
                # the same bytes as plain pinned copies, both directions at once, all ranks together
                "host_ceiling_ms_per_step": ceil_ms, "frac_of_host_ceiling": ceil_ms / ms, "note": note}

    e2e = None if args.no_e2e else e2e_of(w, max(3, min(steps, 10)))
    e2e_inf = None
    if not args.no_e2e and w.plan is None and w.M % 4 == 0:
        try:
            nst = max(3, min(steps, 10))
            ms, h2d, d2h = e2e_infer(cx, w, nst)
            cells = torch.tensor([float(w.cells)], dtype=torch.float64, device=cx.dev)
            if cx.world > 1:
                dist.all_reduce(cells, op=dist.ReduceOp.SUM)
            ceil_ms = host_ceiling(cx, h2d, d2h)
            e2e_inf = {"value": float(cells.item()) / (ms * 1e-3), "unit": "cell-updates/s", "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "ms_per_step": ms, "steps": nst,
                       "host_ceiling_ms_per_step": ceil_ms, "frac_of_host_ceiling": ceil_ms / ms,
                       "note": "pinned host theta/A -> align.HostAligner (chunked H2D | fwd+bwd+traceback on the device | "
                               "D2H of the alignment paths and scores only), per rank"}
        except Exception as e:
            e2e_inf = {"error": f"{type(e).__name__}: {str(e)[:200]}"}

    line = None
    if cx.rank == 0:
        rf = roofline(cx, w, r, traffic_tab)
        step_gbs = w.cells * (BYTES_FWD + BYTES_BWD) / ((r["ms_total"] / steps) * 1e-3) / 1e9
        rf["step"] = {"GBps_at_36B_per_cell": step_gbs, "frac": step_gbs / rf["peak"]}
        line = {
            "metric": METRIC, "value": value, "unit": "cell-updates/s", "n_gpus": cx.world,
            "steps": steps, "warmup": warmup, "ms_per_step": r["ms_total"] / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": w.desc, "mode": w.mode, "pairs_per_gpu": w.B, "N": w.N, "M": w.M,
                       "global_pairs": w.B * cx.world, "parallelism": f"batch-sliced x{cx.world}, no DP collective",
                       "l2": "inputs (theta+A %.0f MB, Q %.0f MB per GPU) exceed the 126 MB L2; no flush needed"
                             % (2 * w.theta.numel() * 4 / 1e6, w.cells * 8 / 1e6),
                       "launch": "forward and backward captured once in two CUDA graphs through the autograd API, "
                                 "replayed per step" if r["graphed"] else "eager calls",
                       **({"packing": w.stats} if w.stats else {})},
            "roofline": rf,
            "gpu_launches": 2 * steps,
            "host_enqueue_ms_per_step": r["host_enqueue_ms_per_step"],
            "step_ms_median": r["step_ms_median"], "step_ms_first3": r["step_ms_first3"],
            "step_ms_max": r["step_ms_max"], "eager": r["eager"],
            "clocks": clocks,
        }
        if "graph_error" in r:
            line["graph_error"] = r["graph_error"]
        if e2e:
            line["e2e"] = e2e
        if e2e_inf:
            line["e2e_inference"] = e2e_inf
    del r

    # ---- every other BASELINE config under the same clock ---------------------------------------------
    if not args.no_extras:
        extras = {}
        xs = max(3, min(steps, 10))
        for name in ("c2", "c3", "c4", "c5", "b32"):
            if name == args.workload:
                continue
            try:
                wk = make_workload(cx, name)
                rk = time_workload(cx, wk, xs, 3, use_graph=not args.no_graph)
                ent = {"workload": wk.desc, "pairs_per_gpu": wk.B,
                       "value": rk["total_cells"] * xs / (rk["ms_total"] * 1e-3), "unit": "cell-updates/s",
                       "ms_per_step": rk["ms_total"] / xs, "steps": xs, "graphed": rk["graphed"],
                       "host_enqueue_ms_per_step": rk["host_enqueue_ms_per_step"], "eager": rk["eager"]}
                if wk.stats:
                    ent["packing"] = wk.stats
                if cx.rank == 0:
                    ent["roofline"] = roofline(cx, wk, rk, traffic_tab)
                if name == "c5" and not args.no_e2e:
                    ent["e2e"] = e2e_of(wk, 5)
                    if cx.world == 1 and not args.no_cpu_baseline:
                        if _ALL_CPUS:
                            os.sched_setaffinity(0, _ALL_CPUS)
                        v, cores, sample, _ = cpu_port_throughput(wk.mode, wk.N, wk.M, wk.xl, wk.yl, target_s=5.0)
                        ent["cpu_baseline"] = {"value": v, "unit": "cell-updates/s", "cores": cores, "kind": "port",
                                               "sample": sample}
                extras[name] = ent
                del rk, wk
            except Exception as e:
                extras[name] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
        if cx.world == 1:
            for tag, (B, N, M) in (("train_step_c2", (1024, 256, 256)), ("train_step_b32x512", (32, 512, 512))):
                try:
                    t = train_step_bench(cx, B, N, M, 10)
                    t.pop("_keep", None)
                    extras[tag] = t
                except Exception as e:
                    extras[tag] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
                torch.cuda.synchronize()
                torch.cuda.empty_cache()
            try:
                extras["producer_theta_A"] = producer_bench(cx, 1024, 256, 1024, 5)
            except Exception as e:
                extras["producer_theta_A"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
        if cx.rank == 0:
            line["workloads"] = extras

    if cx.rank == 0:
        if cx.world == 1 and not args.no_cpu_baseline:
            if _ALL_CPUS:
                os.sched_setaffinity(0, _ALL_CPUS)        # the CPU baseline gets every host core again
            v, cores, sample, _ = cpu_port_throughput(w.mode, w.N, w.M, w.xl, w.yl, target_s=args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": "cell-updates/s", "cores": cores, "kind": "port",
                                    "sample": sample}
        print(json.dumps(line))
    if cx.world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
