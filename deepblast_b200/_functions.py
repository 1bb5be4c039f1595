"""torch.autograd glue shared by deepblast_b200.nw_cuda and deepblast_b200.sw_cuda.

Mirrors deepblast/nw_cuda.py:168-325 (and sw_cuda.py:168-326) name for name:
same forward/backward signatures, same saved tensors, same returned tuples
(grad wrt A is A itself, nw_cuda.py:206-207; double backward returns None for A,
nw_cuda.py:262), same TypeError / NotImplementedError checks (nw_cuda.py:171-175).
Not replicated on purpose: torch.autograd.set_detect_anomaly(True) at import
(nw_cuda.py:9) and the A[last, j-1] indexing of the Numba kernels (nw_cuda.py:61);
results follow deepblast/nw.py / sw.py.
"""
import torch
import torch.nn as nn

from . import ops


def make_classes(mode, prefix):
    class FunctionBackward(torch.autograd.Function):
        @staticmethod
        def forward(ctx, theta, A, Et, Q, operator, interior=False, keep_interior=False, plan=None):
            # `interior` (not part of the reference's signature, nw_cuda.py:212): return
            # E[:, 1:-1, 1:-1] instead of the padded E, so that the double backward receives the
            # gradient of the interior directly instead of a zero-padded copy it would slice again.
            # `plan` (ours as well): the batch runs on the strip-queue kernels (ops.sq_*), whose E
            # IS the interior, in the plan's layout (dense [B, N, M] or a flat packed buffer).
            if operator != 'softmax':
                raise NotImplementedError(
                    "CUDA variant only supports 'softmax' operator")
            ctx.set_materialize_grads(False)
            ctx.others = operator
            ctx.interior = bool(interior)
            ctx.plan = plan
            ctx.uplan = None
            if plan is not None:
                E = ops.sq_backward(plan, Et, Q, mode)
                ctx.save_for_backward(Q, E, None)
                ctx.lens = (None, None)
                if interior or plan.packed:
                    return E, A
                Epad = torch.zeros((plan.B, plan.N + 2, plan.M + 2), dtype=E.dtype, device=E.device)
                Epad[:, 1:-1, 1:-1] = E
                Epad[:, plan.N + 1, plan.M + 1] = Et           # nw.py:125-127
                return Epad, A
            lens = getattr(Q, "_b200dp_lens", (None, None))
            # Q: the engine's strip-major view (from our forward) or a dense
            # reference-layout [B,N+2,M+2,3] tensor (converted on the fly)
            if Q.dim() == 4:
                Q = ops.q_from_reference(Q)
            ctx.uplan = None
            if interior and lens == (None, None) and Q.dim() == 5 and ops.SQ_MODE != "never" and theta.shape[2] % 4 == 0:
                # large equal-size batch whose forward ran on the chained kernel: the backward sweep
                # on the strip-queue kernel (the same strip-major Q) writes the contiguous interior
                # directly -- no padded E, no second copy for the double backward, and more warps per
                # SM than one per pair (C2: 0.170 against 0.185 ms)
                from . import plan as _plan
                B_, N_, M_ = theta.shape
                ctx.uplan = _plan.get_plan(B_, N_, M_, None, None, False, Q.device)
                ctx.dims = (B_, N_, M_)
                Ei = ops.sq_backward(ctx.uplan, Et, Q, mode)
                ctx.save_for_backward(Q, None, Ei)
                ctx.lens = lens
                return Ei, A
            # keep_interior (set by Function.backward when a double backward may follow, i.e.
            # under create_graph): the chained sweep also keeps a contiguous copy of E's
            # interior for the adjoint forward sweep
            if keep_interior:
                E, Ei = ops.backward_pass(Et, Q, mode, lens[0], lens[1], N=theta.shape[1], keep_interior=True)
            else:
                E, Ei = ops.backward_pass(Et, Q, mode, lens[0], lens[1], N=theta.shape[1]), None
            # an unused output gradient arrives as None, not as a tensor of zeros to be read
            ctx.save_for_backward(Q, E, Ei)
            ctx.lens = lens
            return (E[:, 1:-1, 1:-1] if interior else E), A

        @staticmethod
        def backward(ctx, Ztheta, ZA):
            Q, E, Ei = ctx.saved_tensors
            plan = ctx.plan
            if plan is not None:
                if Ztheta is None:
                    Zt = torch.zeros_like(E)
                elif ctx.interior or plan.packed:
                    Zt = Ztheta
                else:
                    Zt = Ztheta[:, 1:-1, 1:-1]
                Vtd, QdE = ops.sq_adjoint_forward(plan, Q, Zt, ZA, E)
                Ed = ops.sq_adjoint_backward(plan, Q, QdE)
                return Ed, None, Vtd, None, None, None, None, None
            if ctx.uplan is not None:
                fast = ops.adjoint_pair_fast(Q, None, Ztheta, ZA, interior=True, Ei=Ei, interior_out=True, dims=ctx.dims)
                if fast is not None:
                    Vtd, Ed = fast
                    return Ed, None, Vtd, None, None, None, None, None
                Zt = torch.zeros_like(Ei) if Ztheta is None else Ztheta
                Vtd, QdE = ops.sq_adjoint_forward(ctx.uplan, Q, Zt, ZA, Ei)
                Ed = ops.sq_adjoint_backward(ctx.uplan, Q, QdE)
                return Ed, None, Vtd, None, None, None, None, None
            xl, yl = ctx.lens
            if xl is None and yl is None and (Ztheta is None or Ztheta.dtype == torch.float32):
                # large batches of equal-size lattices: both sweeps on the chained kernels
                # (ZA stays None when the caller did not use the A passthrough: no zeros to read)
                fast = ops.adjoint_pair_fast(Q, E, Ztheta, ZA, interior=ctx.interior, Ei=Ei, interior_out=True)
                if fast is not None:
                    Vtd, Ed = fast                      # Ed: contiguous [B, N, M], nothing to slice or clone
                    return Ed, None, Vtd, None, None, None, None, None
            if Ztheta is None:
                Ztheta = torch.zeros_like(E)
            elif ctx.interior:
                Zpad = torch.zeros_like(E)
                Zpad[:, 1:-1, 1:-1] = Ztheta
                Ztheta = Zpad
            B, ZN, ZM = Ztheta.shape
            if ZA is None:
                ZA = torch.zeros((B, ZN - 2, ZM - 2), dtype=Ztheta.dtype, device=Ztheta.device)
            Vtd, Qd = ops.adjoint_forward_pass(Q, Ztheta, ZA, xl, yl)
            Ed = ops.adjoint_backward_pass(E, Q, Qd, xl, yl)
            Ed = Ed[:, 1:-1, 1:-1]
            return Ed, None, Vtd, None, None, None, None, None

    class Function(torch.autograd.Function):
        @staticmethod
        def forward(ctx, theta, A, operator, xlen=None, ylen=None, plan=None):
            # xlen / ylen / plan are ours (the reference's forward takes theta, A, operator,
            # nw_cuda.py:170): per-pair lattice sizes, and an explicit plan.Plan (required for the
            # packed layout, where theta and A are flat buffers)
            if operator != 'softmax':
                raise NotImplementedError(
                    "CUDA variant only supports 'softmax' operator")
            if theta.dtype != torch.float32:
                raise TypeError("CUDA variant only supports torch.float32 type")
            if plan is None:
                plan = ops.route_plan(theta, xlen, ylen)
            if plan is None and not (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
                # nothing can ask for a gradient (NeuralAligner.score, alignment.py:127-137): the score-only
                # strip-queue forward writes no Q at all, also for the large batches that otherwise take the
                # chained kernel (1024 x 256^2: 0.152 against 0.202 ms)
                plan = ops.score_plan(theta, xlen, ylen)
            ctx.plan = plan
            ctx.others = operator
            if plan is not None:
                # no Q when nothing can ask for a gradient (NeuralAligner.score, alignment.py:127-137)
                need_q = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
                Vt, Q = ops.sq_forward(plan, theta, A, mode, need_q=need_q)
                ctx.save_for_backward(theta, A, Q)
                ctx.lens = (None, None)
                return Vt
            Vt, Q = ops.forward_pass(theta, A, mode, xlen, ylen)
            if xlen is not None or ylen is not None:
                Q._b200dp_lens = (xlen, ylen)
            ctx.save_for_backward(theta, A, Q)
            ctx.lens = (xlen, ylen)
            return Vt

        @staticmethod
        def backward(ctx, Et):
            theta, A, Q = ctx.saved_tensors
            operator = ctx.others
            if ctx.plan is not None:
                E, A = FunctionBackward.apply(theta, A, Et, Q, operator, True, False, ctx.plan)
                return E, A, None, None, None, None
            if ctx.lens != (None, None):
                Q._b200dp_lens = ctx.lens
            # a second copy of E's interior only when a double backward can follow AND reach theta
            # (create_graph with a differentiable Et or theta); plain inference never pays for it
            keep = torch.is_grad_enabled() and (Et.requires_grad or theta.requires_grad)
            E, A = FunctionBackward.apply(theta, A, Et, Q, operator, True, keep)
            return E, A, None, None, None, None

    class Decoder(nn.Module):
        def __init__(self, operator):
            super().__init__()
            self.operator = operator

        def forward(self, theta, A, xlen=None, ylen=None, plan=None):
            if xlen is None and ylen is None and plan is None:
                return Function.apply(theta, A, self.operator)
            return Function.apply(theta, A, self.operator, xlen, ylen, plan)

        def traceback(self, grad):
            """Greedy walk over one [N, M] expected-alignment matrix; the rule of
            deepblast/nw_cuda.py:273-317, evaluated by one kernel launch."""
            if not torch.is_tensor(grad):
                grad = torch.as_tensor(grad)
            if not grad.is_cuda:
                grad = grad.cuda()
            return ops.traceback_batch(grad.unsqueeze(0), variant="cuda")[0]

        def traceback_batch(self, grad, xlen=None, ylen=None, variant="cuda"):
            """All pairs of a [B, N, M] batch in one launch."""
            return ops.traceback_batch(grad, xlen, ylen, variant)

        def decode(self, theta, A, xlen=None, ylen=None, plan=None):
            """ Shortcut for doing inference.  (xlen / ylen: per-pair lattice sizes; plan: a
            plan.Plan, e.g. a packed one -- theta, A and the result are then flat packed buffers.) """
            with torch.enable_grad():
                nll = self.forward(theta, A, xlen, ylen, plan)
                v = torch.sum(nll)
                v_grad, _ = torch.autograd.grad(v, (theta, A), create_graph=True)
            return v_grad

        def decode_host(self, theta_h, A_h, **kw):
            """`decode` for HOST tensors: (Vt, dVt/dtheta) as pinned host tensors, with the
            PCIe copies of consecutive chunks overlapped with the sweeps (ops.decode_host)."""
            return ops.decode_host(theta_h, A_h, mode, **kw)

    for cls, suffix in ((FunctionBackward, "FunctionBackward"), (Function, "Function"),
                        (Decoder, "Decoder")):
        cls.__name__ = prefix + suffix
        cls.__qualname__ = prefix + suffix
    return Function, FunctionBackward, Decoder
