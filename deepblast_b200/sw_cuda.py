"""Drop-in for deepblast.sw_cuda (reference deepblast/sw_cuda.py:168-326):
SmithWatermanFunction, SmithWatermanFunctionBackward, SmithWatermanDecoder.
Numerics follow deepblast/sw.py (forward/backward loops start/stop at 2,
sw.py:54-55,107-109; the adjoint sweeps cover the full range, sw.py:150-151)."""
from ._functions import make_classes

(SmithWatermanFunction,
 SmithWatermanFunctionBackward,
 SmithWatermanDecoder) = make_classes("sw", "SmithWaterman")
for _c in (SmithWatermanFunction, SmithWatermanFunctionBackward, SmithWatermanDecoder):
    _c.__module__ = __name__

__all__ = ["SmithWatermanFunction", "SmithWatermanFunctionBackward", "SmithWatermanDecoder"]
