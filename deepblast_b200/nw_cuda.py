"""Drop-in for deepblast.nw_cuda (reference deepblast/nw_cuda.py:168-325):
NeedlemanWunschFunction, NeedlemanWunschFunctionBackward, NeedlemanWunschDecoder
backed by the sm_100a kernels of libb200dp.so.  Numerics follow deepblast/nw.py."""
from ._functions import make_classes

(NeedlemanWunschFunction,
 NeedlemanWunschFunctionBackward,
 NeedlemanWunschDecoder) = make_classes("nw", "NeedlemanWunsch")
for _c in (NeedlemanWunschFunction, NeedlemanWunschFunctionBackward, NeedlemanWunschDecoder):
    _c.__module__ = __name__

__all__ = ["NeedlemanWunschFunction", "NeedlemanWunschFunctionBackward", "NeedlemanWunschDecoder"]
