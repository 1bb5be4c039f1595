"""deepblast_b200 -- B200 (sm_100a) engine for DeepBLAST's differentiable soft-DP
alignment path (NeedlemanWunsch / SmithWaterman forward, backward, adjoint sweeps
and traceback), behind the reference's own torch.autograd.Function / nn.Module API.

    from deepblast_b200.nw_cuda import NeedlemanWunschDecoder   # == deepblast.nw_cuda
    import deepblast_b200; deepblast_b200.install()             # patch an installed deepblast
"""
import sys

__version__ = "0.1.0"


def install():
    """Rebind the reference's CUDA decoder classes to this engine so that
    deepblast.alignment.NeuralAligner and deepblast.trainer run unchanged
    (deepblast/alignment.py:3-6 binds them at import time, alignment.py:67-79
    instantiates them).  Works before or after `deepblast.alignment` is imported.

    Importing deepblast.nw_cuda / sw_cuda switches torch's autograd anomaly mode on for the whole
    process as a side effect (`torch.autograd.set_detect_anomaly(True)` at module level,
    nw_cuda.py:9, sw_cuda.py:9), which slows every backward pass; the mode the caller had is
    restored here."""
    import importlib
    import torch
    from . import nw_cuda, sw_cuda
    patched = []
    anomaly = torch.is_anomaly_enabled()
    try:
        for ref_name, ours in (("deepblast.nw_cuda", nw_cuda), ("deepblast.sw_cuda", sw_cuda)):
            try:
                ref = importlib.import_module(ref_name)
            except Exception:      # reference (or numba) not importable: nothing to patch there
                continue
            for n in ours.__all__:
                setattr(ref, n, getattr(ours, n))
                patched.append(f"{ref_name}.{n}")
    finally:
        if torch.is_anomaly_enabled() != anomaly:
            torch.autograd.set_detect_anomaly(anomaly)
    ali = sys.modules.get("deepblast.alignment")
    if ali is not None:
        ali.NWDecoderCUDA = nw_cuda.NeedlemanWunschDecoder
        ali.SWDecoderCUDA = sw_cuda.SmithWatermanDecoder
        patched += ["deepblast.alignment.NWDecoderCUDA", "deepblast.alignment.SWDecoderCUDA"]
    return patched
