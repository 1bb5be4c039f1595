"""CPU-side tests: the C-ABI library loads and exports every symbol the header
declares, host logic (layout, sharding), loud failure without a GPU, and a
world_size-2 gloo run of the sharding logic."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    from deepblast_b200 import _lib
    from deepblast_b200.build import build
    build()
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "b200dp.h")).read()
    declared = set(re.findall(r"\b(b200dp_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name)
    assert L.b200dp_version() >= 100


def test_q_layout_is_a_valid_strided_view():
    from deepblast_b200 import _lib
    for N, M in [(1, 1), (5, 4), (31, 33), (32, 32), (256, 256), (300, 77), (1000, 2047)]:
        K, ss, ps, pad = _lib.q_layout(N, M)
        assert K == (N + 31) // 32 and ss == M * 64 and ps == K * ss + 31 * 64 and pad >= 16 * 64
        # the 5-D view [K, 32, M, 2] (stored states x, y) with strides (ss, 65, 64, 32) addresses
        # distinct elements inside the pair's storage: cell (i, j, c) -> k*ss + ((j-1)+t)*64 + c*32 + t
        if K * 32 * M <= 40000:
            k, t, j0, s = np.meshgrid(np.arange(K), np.arange(32), np.arange(M), np.arange(2), indexing="ij")
            addr = k * ss + t * 65 + j0 * 64 + s * 32
            assert addr.min() >= 0 and addr.max() < ps
            assert len(np.unique(addr)) == addr.size
            np.testing.assert_array_equal(addr, k * ss + (j0 + t) * 64 + s * 32 + t)


def test_argument_errors_do_not_need_a_gpu():
    from deepblast_b200 import _lib
    L = _lib.lib()
    assert L.b200dp_q_layout(0, 5, None, None, None, None) < 0
    assert b"N >= 1" in L.b200dp_last_error()
    assert L.b200dp_fwd(0, 0, 0, 0, 0, 0, 1, 0, 4, 0, 0, 0) < 0
    assert L.b200dp_fwd(0, 0, 0, 0, 0, 0, 1, 4, 4, 7, 0, 0) < 0      # bad mode
    assert L.b200dp_traceback(0, 0, 0, 0, 0, 0, 1, 4, 4, 5, 0, 1, 0, 0) < 0


def test_cpu_tensors_fail_loudly():
    from deepblast_b200.nw_cuda import NeedlemanWunschDecoder, NeedlemanWunschFunction
    theta = torch.rand(1, 4, 4, requires_grad=True)
    A = -torch.ones(1, 4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        NeedlemanWunschFunction.apply(theta, A, 'softmax')
    with pytest.raises(NotImplementedError):
        NeedlemanWunschFunction.apply(theta, A, 'hardmax')
    with pytest.raises(TypeError):
        NeedlemanWunschFunction.apply(theta.double(), A.double(), 'softmax')
    assert NeedlemanWunschDecoder('softmax').operator == 'softmax'


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "deepblast_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle accumulates", ""), f


def test_sharding_helpers():
    from deepblast_b200.sharding import lpt_assign, packing_stats, shard_range
    for B, W in [(8192, 8), (1024, 3), (5, 8), (0, 2)]:
        ranges = [shard_range(B, W, r) for r in range(W)]
        assert ranges[0][0] == 0 and ranges[-1][1] == B
        assert all(ranges[r][1] == ranges[r + 1][0] for r in range(W - 1))
        sizes = [hi - lo for lo, hi in ranges]
        assert max(sizes) - min(sizes) <= 1
    rng = np.random.default_rng(0)
    k = rng.choice(np.arange(1, 17), size=8192, p=(1 / np.arange(1, 17)) / (1 / np.arange(1, 17)).sum())
    k2 = rng.choice(np.arange(1, 17), size=8192, p=(1 / np.arange(1, 17)) / (1 / np.arange(1, 17)).sum())
    xlen, ylen = 64 * k, 64 * k2
    asg = lpt_assign(xlen * ylen, 8)
    assert sorted(np.concatenate(asg).tolist()) == list(range(8192))
    st = packing_stats(xlen, ylen, asg)
    assert st["imbalance"] < 1.01 and st["packing_efficiency"] == 1.0
    for a in asg:
        c = (xlen * ylen)[a]
        assert (np.diff(c) <= 0).all()


WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from deepblast_b200.sharding import shard_range, lpt_assign
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2],
                        rank=int(sys.argv[3]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
B, N, M = 37, 8, 8
lo, hi = shard_range(B, world, rank)
cells = torch.tensor([float((hi - lo) * N * M)])
dist.all_reduce(cells)                      # the only collective: a scalar
assert cells.item() == B * N * M
owned = torch.zeros(B); owned[lo:hi] = 1
dist.all_reduce(owned)
assert bool((owned == 1).all())             # disjoint cover
rng = np.random.default_rng(0)
c = rng.integers(1, 1000, size=101)
mine = lpt_assign(c, world)[rank]
tot = torch.tensor([float(c[mine].sum())]); dist.all_reduce(tot)
assert tot.item() == float(c.sum())
t = torch.tensor([1.0 + rank]); dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == 2.0                      # max-over-ranks timing reduction
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_state_string_matches_revstate_f():
    """deepblast/dataset/utils.py:32-38 (revstate_f) + trainer.py:86-87."""
    from deepblast_b200.align import state_string
    decoded = [(0, 0, 1), (1, 0, 0), (1, 1, 2), (2, 2, 1)]
    assert state_string(decoded) == ":12:"
    assert state_string([]) == ""


def test_new_entry_points_fail_loudly_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    from deepblast_b200 import ops
    from deepblast_b200.losses import MatrixCrossEntropy
    t = torch.rand(2, 8, 8)
    with pytest.raises(RuntimeError):
        ops.decode_host(t, t, "nw")
    with pytest.raises(RuntimeError):
        MatrixCrossEntropy()(t, t, [8, 8], [8, 8], torch.ones_like(t))
    with pytest.raises(TypeError):
        ops.decode_host(t.double(), t.double(), "nw")


def test_host_entry_argument_errors():
    """b200dp_decode_host / b200dp_mxent_* reject bad arguments before touching the device."""
    from deepblast_b200 import _lib
    L = _lib.lib()
    assert L.b200dp_decode_host_workspace(256, 256, 64) > 3 * 64 * (2 * 256 * 256 * 4)
    assert L.b200dp_decode_host_workspace(0, 256, 64) == 0
    assert L.b200dp_decode_host(None, None, None, None, None, 4, 8, 8, 0, 2, None, 0, 0, None) != 0
    assert b"null pointer" in L.b200dp_last_error()
    assert L.b200dp_decode_host(None, None, None, None, None, 4, 8, 8, 7, 2, None, 0, 0, None) != 0
    assert b"bad mode" in L.b200dp_last_error()
    assert L.b200dp_mxent_fwd(None, None, 0, 0, None, None, None, 4, 8, 8, None, None, None) != 0
    assert b"null pointer" in L.b200dp_last_error()
    assert L.b200dp_mxent_fwd(None, None, 0, 0, None, None, None, 0, 8, 8, None, None, None) == 0     # empty batch


def test_bench_affinity_helper_never_raises():
    sys.path.insert(0, ROOT)
    import bench
    assert isinstance(bench.bind_to_gpu_numa_node(0), str)


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU port on the host cores) prints one JSON line with
    the keys the driver reads; under torchrun only rank 0 prints."""
    import json
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-seconds", "0.5"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "cell-updates/s"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["metric"].startswith("DP cell-updates/sec")
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-seconds", "0.5"], capture_output=True, text=True, env=env, timeout=60)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.parametrize("mode", ["nw", "sw"])
@pytest.mark.parametrize("name", ["r8x8", "r17x23", "r33x40", "r64x48"])
def test_strip_major_two_state_layout_round_trip_on_cpu(golden, name, mode):
    """The engine's Q layout (two stored states, implied m state, zero marks) against the
    reference-generated golden Q / Qd: conversion there and back on CPU tensors, element
    addresses as documented in include/b200dp.h."""
    from deepblast_b200 import _lib, ops
    Qref = torch.from_numpy(golden[f"{name}/{mode}/Q"])
    B, N2, M2, _ = Qref.shape
    N, M = N2 - 2, M2 - 2
    Q5 = ops.q_from_reference(Qref)
    K, ss, ps, pad = _lib.q_layout(N, M)
    assert tuple(Q5.shape) == (B, K, 32, M, 2) and tuple(Q5.stride()) == (ps, ss, 65, 64, 32)
    flat = Q5.as_strided((B * ps + pad,), (1,))                # the raw storage
    for (b, i, j) in [(0, 1, 1), (B - 1, N, M), (0, min(N, 33), 2), (B - 1, 2, M)]:
        k, t = (i - 1) // 32, (i - 1) % 32
        base = b * ps + k * ss + ((j - 1) + t) * 64 + t
        x, m, y = Qref[b, i, j].tolist()
        if x == 0 and m == 0 and y == 0:                         # sw.py first row / column: the mark
            assert flat[base].item() < 0
        else:
            assert flat[base].item() == x and flat[base + 32].item() == min(y, np.float32(1) - np.float32(x))
    back = ops.q_to_reference(Q5, N)
    np.testing.assert_array_equal(back[..., 0].numpy(), Qref[..., 0].numpy())
    np.testing.assert_allclose(back.numpy(), Qref.numpy(), rtol=0, atol=2e-7)
    assert np.array_equal((back.sum(-1) == 0).numpy(), (Qref.sum(-1) == 0).numpy())
    assert float(back.min()) >= 0.0                                # the implied state never goes negative
    Qd = torch.from_numpy(golden[f"{name}/{mode}/Qd"])
    backd = ops.q_to_reference(ops.q_from_reference(Qd, "qd"), N, "qd")
    scale = max(1.0, float(Qd.abs().max()))
    np.testing.assert_allclose(backd[:, 1:-1, 1:-1].numpy(), Qd[:, 1:-1, 1:-1].numpy(), rtol=0, atol=2e-6 * scale)


def test_header_is_plain_c_and_bindings_match_it(tmp_path):
    """include/b200dp.h compiles as C (the boundary is a C ABI), and every ctypes binding in
    deepblast_b200/_lib.py has as many arguments as the declaration it binds."""
    from deepblast_b200 import _lib
    hdr_path = os.path.join(ROOT, "include", "b200dp.h")
    src = tmp_path / "use_header.c"
    src.write_text('#include "b200dp.h"\nint main(void) { int (*f)(void) = b200dp_version; const char* (*g)(void) = b200dp_last_error; return f == 0 || g == 0; }\n')
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only",
                        "-I", os.path.dirname(hdr_path), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    hdr = re.sub(r"/\*.*?\*/", "", open(hdr_path).read(), flags=re.S)
    decls = dict(re.findall(r"\b(b200dp_[a-z_0-9]+)\s*\(([^;{]*)\)\s*;", hdr))
    assert set(decls) == set(_lib.EXPORTS)
    for name, args in decls.items():
        args = args.strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        assert n == len(_lib.EXPORTS[name][1]), (name, n, len(_lib.EXPORTS[name][1]))


REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "deepblast")), reason="reference checkout not present")
def test_install_on_the_real_reference_alignment_module():
    """The real caller: deepblast.alignment (imported from the reference checkout with a stub `Bio`,
    SURVEY.md section 8c) picks up our decoders through install() -- NeuralAligner(device='cuda').ddp
    is ours for both alignment modes (alignment.py:67-79) -- and install() leaves autograd's anomaly
    mode as it found it (importing deepblast.nw_cuda switches it on, nw_cuda.py:9).  Runs in a
    subprocess so that the reference modules do not leak into the other tests."""
    code = r"""
import sys, types
sys.path.insert(0, %r); sys.path.insert(0, %r)
bio = types.ModuleType("Bio"); seqio = types.ModuleType("Bio.SeqIO"); bio.SeqIO = seqio
sys.modules["Bio"] = bio; sys.modules["Bio.SeqIO"] = seqio
import torch
assert not torch.is_anomaly_enabled()
import deepblast_b200
patched = deepblast_b200.install()
assert not torch.is_anomaly_enabled(), "install() left anomaly mode on"
assert "deepblast.nw_cuda.NeedlemanWunschDecoder" in patched and "deepblast.sw_cuda.SmithWatermanDecoder" in patched
from deepblast.alignment import NeuralAligner            # binds NWDecoderCUDA / SWDecoderCUDA at import
nw = NeuralAligner(22, 16, 16, 16, n_layers=1, device='cuda')
sw = NeuralAligner(22, 16, 16, 16, n_layers=1, device='cuda', alignment_mode='smith-waterman')
assert type(nw.ddp).__module__ == "deepblast_b200.nw_cuda", type(nw.ddp)
assert type(sw.ddp).__module__ == "deepblast_b200.sw_cuda", type(sw.ddp)
assert nw.ddp.operator == 'softmax'
# the other order: alignment imported first, install() afterwards
import deepblast.alignment as ali
ali.NWDecoderCUDA = None
deepblast_b200.install()
assert ali.NWDecoderCUDA is deepblast_b200.nw_cuda.NeedlemanWunschDecoder
torch.autograd.set_detect_anomaly(True)
deepblast_b200.install()
assert torch.is_anomaly_enabled(), "install() must not switch a caller's anomaly mode off either"
print("ok")
""" % (ROOT, REFERENCE)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_plan_builder_orders_every_strip_after_its_dependency():
    """b200dp_plan_build (host code of the C ABI, no GPU): every strip of both tables comes after the
    strip it depends on, offsets do not overlap, lengths are clamped, empty pairs have no strips."""
    from deepblast_b200 import plan
    rec = np.dtype([('t_off', '<i8'), ('q_off', '<i8'), ('b_in', '<i8'), ('b_out', '<i8'), ('rows', '<i4'), ('m', '<i4'),
                    ('pitch', '<i4'), ('pair', '<i4'), ('flags', '<i4'), ('k', '<i4'), ('pad', '<i4', 2)])
    rng = np.random.default_rng(3)
    B, N, M = 40, 300, 260
    xl = rng.integers(-5, 340, B)
    yl = rng.integers(-5, 300, B)
    for packed in (False, True):
        if not packed:
            p = plan.Plan(B, N, M, xl, yl, packed=False, device="cpu", resident_warps=(37, 11))
        else:
            p = plan.Plan(B, N, M, xl, yl, packed=True, device="cpu", resident_warps=(37, 11))
        n = np.clip(xl, 0, N); m = np.clip(yl, 0, M)
        n[(n == 0) | (m == 0)] = 0
        K = (n + 31) // 32
        assert p.nstrips == int(K.sum()) and p.cells == int((n.astype(np.int64) * np.where(n > 0, m, 0)).sum())
        for d, tab in enumerate(p.tabs_host):
            t = tab.view(rec).reshape(-1)[:p.nstrips]
            seen = set()
            for r in t:
                b, k = int(r['pair']), int(r['k'])
                dep = k - 1 if d == 0 else k + 1
                if 0 <= dep < K[b]:
                    assert (b, dep) in seen, (d, b, k)
                assert (b, k) not in seen
                seen.add((b, k))
                assert r['rows'] == min(32, n[b] - 32 * k) and r['m'] == m[b]
                assert (r['b_in'] >= 0) == (0 <= dep < K[b])
            assert len(seen) == p.nstrips
        # packed operands: pair blocks do not overlap and fit
        if packed:
            ends = p.pair_off[:B] + n.astype(np.int64) * p.pitch
            order = np.argsort(p.pair_off[:B], kind="stable")
            assert np.all(p.pair_off[:B][order][1:] >= ends[order][:-1]) and int(ends.max()) <= p.packed_floats


def test_packed_layout_round_trip_on_cpu():
    from deepblast_b200 import plan
    B, N, M = 5, 40, 37
    xl, yl = [40, 3, 17, 0, 25], [37, 37, 5, 9, 36]
    p = plan.Plan(B, N, M, xl, yl, packed=True, device="cpu")
    g = torch.Generator().manual_seed(0)
    dense = torch.rand(B, N, M, generator=g)
    flat = p.pack(dense)
    back = p.unpack(flat)
    for b in range(B):
        n, m = int(p.xlen[b]), int(p.ylen[b])
        assert torch.equal(back[b, :n, :m], dense[b, :n, :m])
        assert float(back[b, n:].abs().sum()) == 0 and float(back[b, :, m:].abs().sum()) == 0
