#!/usr/bin/env python
"""bench.py -- DP cell-updates/s (forward + backward) of the batched soft-DP alignment
path on B200, next to the CPU oracle port timed on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5]
    python bench.py --impl reference ...      # CPU arm (oracle port, all host threads)
    torchrun --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch: NeedlemanWunschFunction.apply
(theta, A -> Vt) then Vt.sum().backward() (-> dVt/dtheta), i.e. the reference's
`_forward_pass_kernel` + `_backward_pass_kernel` (deepblast/nw_cuda.py:46-102) behind
the reference's autograd API.  Metric and byte accounting: SURVEY.md section 8(d):
36 algorithmic bytes per cell-update (fwd: theta 4 + A 4 read, Q 12 written;
bwd: Q 12 read, E 4 written).
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
    "c5": ("nw", 1024, 1024, 1024, "BASELINE configs[4]: variable lengths 64..1024 (Zipf over 16 buckets), 1024 pairs/GPU"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_name(dom, B, N, M, varlen):
    """Which kernel b200dp_fwd / b200dp_bwd dispatch to for this workload (softdp_api.cu):
    the chained single-warp kernels for large batches of equal-size lattices, the hand-off
    kernels otherwise."""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    chained = (not varlen) and B >= 4 * sms and N >= 32 and M >= 64 and M % (16 if dom == "fwd" else 32) == 0
    return f"softdp_{dom}{3 if chained else 2}_kernel"


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
        "config": {"workload": desc, "mode": mode, "N": N, "M": M,
                   "note": "reference is Python+Numba (not compilable to oracle/_ref); this arm times the C oracle "
                           "port of deepblast/nw.py with OpenMP over pairs on a bounded sample per step"},
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs-per-gpu", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=18.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-bind", action="store_true", help="do not bind the process to the GPU-local CPU cores")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from deepblast_b200 import ops
    from deepblast_b200.nw_cuda import NeedlemanWunschFunction
    from deepblast_b200.sw_cuda import SmithWatermanFunction

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: deepblast_b200 has no CPU path")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_numa_node(local) if not args.no_bind else "unbound (--no-bind)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    mode, Bg, N, M, desc = WORKLOADS[args.workload]
    if args.pairs_per_gpu:
        Bg = args.pairs_per_gpu
    Fn = NeedlemanWunschFunction if mode == "nw" else SmithWatermanFunction
    gen = torch.Generator(device=dev).manual_seed(2 + rank)
    theta = torch.rand(Bg, N, M, generator=gen, device=dev)
    A = -torch.rand(Bg, N, M, generator=gen, device=dev)
    theta.requires_grad_()
    xlen = ylen = None
    stats = {}
    if args.workload == "c5":
        from deepblast_b200.sharding import lpt_assign, packing_stats
        xl_all, yl_all = zipf_lengths(Bg * world, np.random.default_rng(0))
        asg = lpt_assign(xl_all * yl_all, world)
        stats = packing_stats(xl_all, yl_all, asg)
        mine = asg[rank][:Bg] if len(asg[rank]) >= Bg else asg[rank]
        Bg = len(mine)
        theta = theta.detach()[:Bg].requires_grad_()
        A = A[:Bg]
        xlen = torch.tensor(xl_all[mine], dtype=torch.int32, device=dev)
        ylen = torch.tensor(yl_all[mine], dtype=torch.int32, device=dev)
        cells_local = int((xl_all[mine].astype(np.int64) * yl_all[mine]).sum())
    elif mode == "sw":
        cells_local = Bg * (N - 1) * (M - 1)
    else:
        cells_local = Bg * N * M

    def step():
        if xlen is None:
            Vt = Fn.apply(theta, A, 'softmax')
        else:
            Vt = Fn.apply(theta, A, 'softmax', xlen, ylen)
        g, = torch.autograd.grad(Vt.sum(), theta)
        return Vt, g

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: W untimed steps, with the rendezvous barrier exercised in between so that
    # every lazy initialisation (NCCL communicator, allocator pools, autograd worker
    # threads) happens before the timed region.  The results of a step stay referenced
    # while the next one runs, exactly as in the timed loop below, so that the caching
    # allocator already owns every block the timed steps will ask for.
    keep = None
    for it in range(max(args.warmup, 3)):
        keep = step()
        if it == 0:
            sync_all()
    sync_all()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # ---- timed region: exactly K steps, CUDA events on the launching stream -------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    sync_all()
    # keep the device busy for ~2 ms (untimed) while the host enqueues the first steps, so
    # that the K timed steps run back to back on the GPU instead of waiting for Python
    torch.cuda._sleep(int(4e6))
    host_t0 = time.perf_counter()
    ev0.record()
    for it in range(args.steps):
        kev[it][0].record()
        if xlen is None:
            Vt = Fn.apply(theta, A, 'softmax')
        else:
            Vt = Fn.apply(theta, A, 'softmax', xlen, ylen)
        kev[it][1].record()
        g, = torch.autograd.grad(Vt.sum(), theta)
        kev[it][2].record()
        keep = (Vt, g)
    ev1.record()
    host_ms = (time.perf_counter() - host_t0) * 1e3 / args.steps     # enqueue time, GPU not waited for
    torch.cuda.synchronize()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    cells = torch.tensor([float(cells_local)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cells, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    total_cells = float(cells.item())
    value = total_cells * args.steps / (ms * 1e-3)
    fwd_ms = float(np.mean([k[0].elapsed_time(k[1]) for k in kev]))
    bwd_ms = float(np.mean([k[1].elapsed_time(k[2]) for k in kev]))

    # ---- end to end: host buffers in, host results out, every step --------------------
    e2e = None
    if not args.no_e2e:
        # the reference-facing call for a caller whose theta / A live in HOST memory:
        # Decoder.decode_host -> C ABI b200dp_decode_host (pinned host buffers in, pinned host
        # Vt and padded E out; chunked upload / fwd / bwd / download pipeline inside).
        h_theta = torch.empty((Bg, N, M), dtype=torch.float32, pin_memory=True).copy_(theta.detach().cpu())
        h_A = torch.empty((Bg, N, M), dtype=torch.float32, pin_memory=True).copy_(A.cpu())
        h_Vt = torch.empty(Bg, dtype=torch.float32, pin_memory=True)
        h_E = torch.empty((Bg, N + 2, M + 2), dtype=torch.float32, pin_memory=True)
        if xlen is None:
            def e2e_step():
                ops.decode_host_async(h_theta, h_A, mode, out=(h_Vt, h_E), device=dev)
            d2h = int(Bg * (N + 2) * (M + 2) * 4 + Bg * 4)
            note = ("pinned host theta/A -> b200dp_decode_host (chunked H2D | fwd+bwd | D2H pipeline, 3 streams) -> "
                    "pinned host Vt and padded E (dVt/dtheta = E[:,1:-1,1:-1]), per rank")
        else:
            h_g = torch.empty((Bg, N, M), dtype=torch.float32, pin_memory=True)

            def e2e_step():
                th = h_theta.to(dev, non_blocking=True).requires_grad_()
                a = h_A.to(dev, non_blocking=True)
                Vt = Fn.apply(th, a, 'softmax', xlen, ylen)
                gg, = torch.autograd.grad(Vt.sum(), th)
                h_Vt.copy_(Vt.detach(), non_blocking=True)
                h_g.copy_(gg, non_blocking=True)
            d2h = int(Bg * N * M * 4 + Bg * 4)
            note = "pinned host theta/A -> H2D -> Function fwd + backward (per-pair lengths) -> Vt and dVt/dtheta D2H, per rank"

        for _ in range(2):
            e2e_step()
        sync_all()
        nst = max(3, min(args.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(nst):
            e2e_step()
        e1.record()
        torch.cuda.synchronize()
        sync_all()
        te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": total_cells * nst / (float(te.item()) * 1e-3), "unit": "cell-updates/s",
               "h2d_bytes_per_step": int(2 * Bg * N * M * 4), "d2h_bytes_per_step": d2h,
               "ms_per_step": float(te.item()) / nst, "steps": nst, "host_affinity": affinity, "note": note}

    if rank == 0:
        peak, peak_src = peaks()
        n_cells_launch = cells_local
        fwd_gbs = n_cells_launch * BYTES_FWD / (fwd_ms * 1e-3) / 1e9
        bwd_gbs = n_cells_launch * BYTES_BWD / (bwd_ms * 1e-3) / 1e9
        dom = "fwd" if fwd_ms >= bwd_ms else "bwd"
        ach = fwd_gbs if dom == "fwd" else bwd_gbs
        step_gbs = n_cells_launch * (BYTES_FWD + BYTES_BWD) / ((ms / args.steps) * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(f"{args.workload}_{dom}")
        line = {
            "metric": METRIC, "value": value, "unit": "cell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": desc, "mode": mode, "pairs_per_gpu": Bg, "N": N, "M": M,
                       "global_pairs": Bg * world, "parallelism": f"batch-sliced x{world}, no DP collective",
                       "l2": "inputs (theta+A %.0f MB, Q %.0f MB per GPU) exceed the 126 MB L2; no flush needed"
                             % (2 * Bg * N * M * 4 / 1e6, Bg * (N + 2) * (M + 2) * 12 / 1e6),
                       **({"packing": stats} if stats else {})},
            "roofline": {"bound": "hbm", "kernel": kernel_name(dom, Bg, N, M, xlen is not None), "achieved": ach,
                         "peak": peak,
                         "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                         # measured DRAM bytes of that launch (ncu, profiles/traffic.json) over the live
                         # duration: the kernels store two of the three Q states, so they move fewer bytes
                         # than the 36 B/cell contract `achieved` is quoted on
                         "traffic_GBps": (traffic / ((fwd_ms if dom == "fwd" else bwd_ms) * 1e-3) / 1e9) if traffic else None,
                         "traffic_frac": (traffic / ((fwd_ms if dom == "fwd" else bwd_ms) * 1e-3) / 1e9 / peak) if traffic else None,
                         "algorithmic_bytes_per_cell": BYTES_FWD if dom == "fwd" else BYTES_BWD,
                         "fwd": {"ms": fwd_ms, "GBps": fwd_gbs, "frac": fwd_gbs / peak},
                         "bwd": {"ms": bwd_ms, "GBps": bwd_gbs, "frac": bwd_gbs / peak},
                         "step": {"GBps_at_36B_per_cell": step_gbs, "frac": step_gbs / peak}},
            "gpu_launches": 2 * args.steps,
            "host_enqueue_ms_per_step": host_ms,
            "step_ms_median": float(np.median([k[0].elapsed_time(k[2]) for k in kev])),
            "step_ms_first3": [round(k[0].elapsed_time(k[2]), 4) for k in kev[:3]],
            "step_ms_max": float(np.max([k[0].elapsed_time(k[2]) for k in kev])),
            "clocks": clocks,
        }
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            if _ALL_CPUS:
                os.sched_setaffinity(0, _ALL_CPUS)        # the CPU baseline gets every host core again
            xl = None if xlen is None else xlen.cpu().numpy()
            yl = None if ylen is None else ylen.cpu().numpy()
            v, cores, sample, _ = cpu_port_throughput(mode, N, M, xl, yl, target_s=args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": "cell-updates/s", "cores": cores, "kind": "port",
                                    "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
