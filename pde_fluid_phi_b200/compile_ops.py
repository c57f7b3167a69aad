"""torch.library custom ops for the two spectral layers: what torch.compile sees.

The reference wraps its models in ``torch.compile(model, mode='max-autotune')``
(optimization/performance_optimization.py:466-486).  dynamo cannot trace the ctypes calls into libpdephi_b200.so, so the
autograd Functions of functional.py run behind ``torch.compiler.disable`` -- correct, but one graph break per layer.  Here the
same kernel chains are registered as opaque custom ops with fake (meta) implementations and autograd formulas, and the mirror
modules route through them whenever ``torch.compiler.is_compiling()``: a compiled model is then ONE graph (``fullgraph=True``
works) whose spectral layers are calls into the library and whose pointwise part (lift, 1x1x1 convolution, activation,
projection) is left to the compiler.  Eager execution does not use these ops (it keeps the fused block route, the transfer
cache and the slab decomposition, none of which fit an op schema of tensors and scalars).

  pdephi_b200::rational_spectral_fwd / _bwd   RationalFourierLayer.forward        (rational_fourier.py:150-172)
  pdephi_b200::complex_spectral_fwd / _bwd    SpectralConv3D.forward              (spectral_layers.py:46-78)
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import Tensor

from . import _cabi
from . import functional as pf

_HAS_CUSTOM_OP = hasattr(torch.library, "custom_op")


def _modes_h(modes):
    mx, my, mz = (int(v) for v in modes)
    return (mx, my, mz // 2 + 1)


if _HAS_CUSTOM_OP:

    # ------------------------------------------------------------------------------------------------
    # RationalFourierLayer
    # ------------------------------------------------------------------------------------------------
    @torch.library.custom_op("pdephi_b200::rational_spectral_fwd", mutates_args=())
    def rational_spectral_fwd(x: Tensor, P: Tensor, Q: Tensor, mask: Tensor, monoP: Tensor, monoQ: Tensor, modes: List[int],
                              eps: float, precision: int, skip0: Optional[Tensor], residual: bool, proj_eps: float, energy: bool,
                              keep: bool) -> List[Tensor]:
        """[out, X, Y, stats, R, Qe]: out = irfftn(StabilityProjection(R(k) mix(rfftn(x)[corner]))) [+ skip0 + x] and what its
        backward reads (R, Qe empty unless `keep`)."""
        for t, n in ((x, "x"), (P, "P_coeffs"), (Q, "Q_coeffs")):
            pf._require_cuda_f32(t, n)
        modes_h = _modes_h(modes)
        grid = tuple(x.shape[2:])
        X = pf.analysis(x, modes_h, _cabi.FORWARD, pf._tf32_route(precision, "forward_analysis", x.device, grid, modes_h))
        Y, stats, R, Qe = pf.rational_mix_fwd(X, P.contiguous(), Q.contiguous(), monoP, monoQ, eps, mask, proj_eps, keep_transfer=keep)
        if not energy:
            stats = pf._unit_stats(stats)
        out = pf.synthesis(Y, grid, mask, stats[:, 2], skip0, x if residual else None, _cabi.FORWARD, precision)
        if not keep:
            R = Qe = x.new_empty(0)
        return [out, X, Y, stats, R, Qe]

    @rational_spectral_fwd.register_fake
    def _(x, P, Q, mask, monoP, monoQ, modes, eps, precision, skip0, residual, proj_eps, energy, keep):
        B, C = x.shape[:2]
        mh = _modes_h(modes)
        O, I = P.shape[:2]
        tf = (O, I, *mh) if keep else (0,)
        return [x.new_empty((B, O, *x.shape[2:])), x.new_empty((B, C, *mh), dtype=torch.complex64),
                x.new_empty((B, O, *mh), dtype=torch.complex64), x.new_empty((B * O, 4)), x.new_empty(tf), x.new_empty(tf)]

    @torch.library.custom_op("pdephi_b200::rational_spectral_bwd", mutates_args=())
    def rational_spectral_bwd(g: Tensor, X: Tensor, Y: Tensor, stats: Tensor, R: Tensor, Qe: Tensor, mask: Tensor, monoP: Tensor,
                              monoQ: Tensor, grid: List[int], precision: int, residual: bool, order_p: int, order_q: int,
                              need_x: bool) -> List[Tensor]:
        """[dx, dP, dQ] (dx empty unless need_x); formulas: SURVEY.md section 8a."""
        g = g.contiguous()
        modes_h = tuple(X.shape[2:])
        grid = tuple(grid)
        O, I = R.shape[:2]
        dZ = pf.analysis(g, modes_h, _cabi.ADJOINT, pf._tf32_route(precision, "backward_analysis", g.device, grid, modes_h))
        dY = pf.projection_bwd(dZ, Y, mask, stats)
        dX, dP, dQ = pf.rational_mix_bwd(X, dY, R, Qe, monoP, monoQ, ((O, I, order_p, order_p, order_p), (O, I, order_q, order_q, order_q)))
        if need_x:
            dx = pf.synthesis(dX, grid, None, None, g if residual else None, None, _cabi.ADJOINT,
                              pf._tf32_route(precision, "adjoint_synthesis", g.device, grid, modes_h))
        else:
            dx = g.new_empty(0)
        return [dx, dP, dQ]

    @rational_spectral_bwd.register_fake
    def _(g, X, Y, stats, R, Qe, mask, monoP, monoQ, grid, precision, residual, order_p, order_q, need_x):
        O, I = R.shape[:2]
        B = X.shape[0]
        dx = g.new_empty((B, I, *grid)) if need_x else g.new_empty(0)
        return [dx, g.new_empty((O, I, order_p, order_p, order_p)), g.new_empty((O, I, order_q, order_q, order_q))]

    def _rs_setup(ctx, inputs, output):
        x, P, Q, mask, monoP, monoQ, modes, eps, precision, skip0, residual, proj_eps, energy, keep = inputs
        _, X, Y, stats, R, Qe = output
        ctx.save_for_backward(X, Y, stats, R, Qe, mask, monoP, monoQ)
        ctx.meta = ([int(v) for v in x.shape[2:]], int(precision), bool(residual), int(P.shape[-1]), int(Q.shape[-1]),
                    skip0 is not None, bool(keep))

    def _rs_backward(ctx, grads):
        g = grads[0]
        grid, precision, residual, op, oq, has0, keep = ctx.meta
        dx = dP = dQ = None
        if keep and g is not None and any(ctx.needs_input_grad[:3]):
            X, Y, stats, R, Qe, mask, monoP, monoQ = ctx.saved_tensors
            need_x = bool(ctx.needs_input_grad[0])
            dx, dP, dQ = rational_spectral_bwd(g, X, Y, stats, R, Qe, mask, monoP, monoQ, grid, precision, residual, op, oq, need_x)
            if not need_x:
                dx = None
        return (dx, dP if ctx.needs_input_grad[1] else None, dQ if ctx.needs_input_grad[2] else None, None, None, None, None,
                None, None, g if has0 else None, None, None, None, None)

    rational_spectral_fwd.register_autograd(_rs_backward, setup_context=_rs_setup)

    # ------------------------------------------------------------------------------------------------
    # SpectralConv3D
    # ------------------------------------------------------------------------------------------------
    @torch.library.custom_op("pdephi_b200::complex_spectral_fwd", mutates_args=())
    def complex_spectral_fwd(x: Tensor, W: Tensor, modes: List[int], precision: int, skip0: Optional[Tensor],
                             residual: bool) -> List[Tensor]:
        """[out, X]: out = irfftn(pad(einsum(rfftn(x)[corner], W[..., :mzh]))) [+ skip0 + x]."""
        pf._require_cuda_f32(x, "x")
        pf._require_cuda_f32(W, "weights", complex_ok=True)
        modes_h = _modes_h(modes)
        grid = tuple(x.shape[2:])
        X = pf.analysis(x, modes_h, _cabi.FORWARD, pf._tf32_route(precision, "forward_analysis", x.device, grid, modes_h))
        Y = pf.complex_mix_fwd(X, W.contiguous(), modes_h[2])
        out = pf.synthesis(Y, grid, None, None, skip0, x if residual else None, _cabi.FORWARD, precision)
        return [out, X]

    @complex_spectral_fwd.register_fake
    def _(x, W, modes, precision, skip0, residual):
        B, C = x.shape[:2]
        return [x.new_empty((B, W.shape[0], *x.shape[2:])), x.new_empty((B, C, *_modes_h(modes)), dtype=torch.complex64)]

    @torch.library.custom_op("pdephi_b200::complex_spectral_bwd", mutates_args=())
    def complex_spectral_bwd(g: Tensor, X: Tensor, W: Tensor, grid: List[int], precision: int, residual: bool,
                             need_x: bool) -> List[Tensor]:
        """[dx, dW] (dx empty unless need_x)."""
        g = g.contiguous()
        modes_h = tuple(X.shape[2:])
        grid = tuple(grid)
        dY = pf.analysis(g, modes_h, _cabi.ADJOINT, pf._tf32_route(precision, "backward_analysis", g.device, grid, modes_h))
        dX, dW = pf.complex_mix_bwd(X, W.contiguous(), dY, modes_h[2])
        if need_x:
            dx = pf.synthesis(dX, grid, None, None, g if residual else None, None, _cabi.ADJOINT,
                              pf._tf32_route(precision, "adjoint_synthesis", g.device, grid, modes_h))
        else:
            dx = g.new_empty(0)
        return [dx, dW]

    @complex_spectral_bwd.register_fake
    def _(g, X, W, grid, precision, residual, need_x):
        dx = g.new_empty((X.shape[0], W.shape[1], *grid)) if need_x else g.new_empty(0)
        return [dx, torch.empty_like(W)]

    def _cs_setup(ctx, inputs, output):
        x, W, modes, precision, skip0, residual = inputs
        ctx.save_for_backward(output[1], W)
        ctx.meta = ([int(v) for v in x.shape[2:]], int(precision), bool(residual), skip0 is not None)

    def _cs_backward(ctx, grads):
        g = grads[0]
        grid, precision, residual, has0 = ctx.meta
        dx = dW = None
        if g is not None and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            X, W = ctx.saved_tensors
            need_x = bool(ctx.needs_input_grad[0])
            dx, dW = complex_spectral_bwd(g, X, W, grid, precision, residual, need_x)
            if not need_x:
                dx = None
        return dx, dW if ctx.needs_input_grad[1] else None, None, None, g if has0 else None, None

    complex_spectral_fwd.register_autograd(_cs_backward, setup_context=_cs_setup)


def compiling() -> bool:
    """True while dynamo traces the caller (and the custom ops exist): the modules then route through the ops above."""
    is_compiling = getattr(getattr(torch, "compiler", None), "is_compiling", None)
    return bool(_HAS_CUSTOM_OP and is_compiling is not None and is_compiling())


def rational_spectral(x, P, Q, mask, monoP, monoQ, modes, eps, precision, skip0, residual, proj):
    """RationalSpectralFn through the custom op (no slab, no transfer cache): the route of a traced RationalFourierLayer."""
    proj_eps, energy = (float(eps), True) if proj is None else (float(proj[0]), bool(proj[1]))
    keep = torch.is_grad_enabled() and any(t.requires_grad for t in (x, P, Q))
    return rational_spectral_fwd(x, P, Q, mask, monoP, monoQ, [int(m) for m in modes], float(eps), int(precision), skip0,
                                 bool(residual), proj_eps, energy, bool(keep))[0]


def complex_spectral(x, W, modes, precision, skip0, residual):
    """ComplexSpectralFn through the custom op: the route of a traced SpectralConv3D."""
    return complex_spectral_fwd(x, W, [int(m) for m in modes], int(precision), skip0, bool(residual))[0]
