"""Host side of the hot path: thin wrappers over the C ABI and the two autograd Functions.

torch is used for what the reference uses it for around this seam -- device memory, the current
stream, autograd bookkeeping -- and nothing else: every FLOP of the spectral layer runs in
libpdephi_b200.so.  CPU tensors raise; there is no fallback.

Replaces the bodies of RationalFourierLayer.forward (operators/rational_fourier.py:150-172) and
SpectralConv3D.forward (operators/spectral_layers.py:46-78) and the autograd graph behind them.
"""
from __future__ import annotations

import itertools
from ctypes import c_void_p
from typing import Optional, Sequence, Tuple

import torch
from torch.autograd.function import once_differentiable

from . import _cabi


def _p(t: Optional[torch.Tensor]):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda_f32(t: torch.Tensor, name: str, complex_ok: bool = False):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}: the pdephi_b200 spectral path is CUDA-only (no CPU fallback)")
    want = (torch.float32, torch.complex64) if complex_ok else (torch.float32,)
    if t.dtype not in want:
        raise RuntimeError(f"{name} has dtype {t.dtype}; the pdephi_b200 spectral path computes in float32 only")


def monomial_exponents(order: int):
    """(a,b,c), a+b+c < order, in the loop order of rational_fourier.py:135-138."""
    return [e for e in itertools.product(range(order), repeat=3) if sum(e) < order]


def monomial_table(modes: Sequence[int], order: int) -> torch.Tensor:
    """mono[t, k] = kx^a ky^b kz^c over the retained box, built with the same torch ops as
    rational_fourier.py:144 so every entry is the float the reference multiplies by."""
    mx, my, mz = modes
    kx, ky, kz = torch.meshgrid(torch.arange(mx, dtype=torch.float32), torch.arange(my, dtype=torch.float32),
                                torch.arange(mz // 2 + 1, dtype=torch.float32), indexing="ij")
    rows = [((kx ** a) * (ky ** b) * (kz ** c)).reshape(-1) for a, b, c in monomial_exponents(order)]
    return torch.stack(rows).contiguous()


# ------------------------------------------------------------------------------------------------
# raw ops (no autograd)
# ------------------------------------------------------------------------------------------------
def analysis(x: torch.Tensor, modes_h: Tuple[int, int, int], direction: int = _cabi.FORWARD,
             precision: int = _cabi.PREC_AUTO, slab=None) -> torch.Tensor:
    """K1.  x [B,C,Nx,Ny,Nz] float32 -> [B,C,mx,my,mzh] complex64 (pruned rfftn / adjoint of synthesis).

    slab = (Nx_total, x_offset): x is the slab [x_offset, x_offset+Nx) of a larger grid and the result is
    that slab's partial sum of the modes."""
    _require_cuda_f32(x, "x")
    if x.dim() != 5:
        raise ValueError(f"expected a [B,C,Nx,Ny,Nz] tensor, got shape {tuple(x.shape)}")
    x = x.contiguous()
    B, C = x.shape[:2]
    grid = tuple(x.shape[2:])
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        plan = _cabi.get_plan(x.device.index, grid, modes_h, slab)
        precision = _cabi.resolve_precision(plan, precision)
        X = torch.empty((B, C, *modes_h), dtype=torch.complex64, device=x.device)
        nbytes = lib.pdephi_plan_workspace_bytes(plan, B * C)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        _cabi.check(lib.pdephi_analysis_r2c(plan, _p(x), _p(X), B * C, direction, precision, _p(ws), nbytes, _stream()),
                    "analysis_r2c")
    return X


def synthesis(Z: torch.Tensor, grid: Tuple[int, int, int], mode_scale: Optional[torch.Tensor] = None,
              row_scale: Optional[torch.Tensor] = None, addend0: Optional[torch.Tensor] = None,
              addend1: Optional[torch.Tensor] = None, direction: int = _cabi.FORWARD,
              precision: int = _cabi.PREC_AUTO, slab=None) -> torch.Tensor:
    """K3.  Z [B,C,mx,my,mzh] complex64 -> [B,C,Nx,Ny,Nz] float32, out = S(Z*mode_scale*row_scale) + addends.

    row_scale: float32 1-D (possibly strided) view with one element per volume, e.g. ``stats[:, 2]``.
    """
    _require_cuda_f32(Z, "Z", complex_ok=True)
    if Z.dim() != 5 or not Z.is_complex():
        raise ValueError(f"expected a complex [B,C,mx,my,mzh] tensor, got {Z.dtype} {tuple(Z.shape)}")
    Z = Z.contiguous()
    B, C = Z.shape[:2]
    modes_h = tuple(Z.shape[2:])
    rs_stride = 0
    if row_scale is not None:
        if row_scale.numel() != B * C or row_scale.dim() != 1:
            raise ValueError("row_scale must be a 1-D view with one element per (batch, channel)")
        rs_stride = row_scale.stride(0)
    for a in (addend0, addend1):
        if a is not None:
            _require_cuda_f32(a, "addend")
            if tuple(a.shape) != (B, C, *grid):
                raise ValueError(f"addend shape {tuple(a.shape)} != {(B, C, *grid)}")
    a0 = addend0.contiguous() if addend0 is not None else None
    a1 = addend1.contiguous() if addend1 is not None else None
    lib = _cabi.load()
    with torch.cuda.device(Z.device):
        plan = _cabi.get_plan(Z.device.index, grid, modes_h, slab)
        precision = _cabi.resolve_precision(plan, precision)
        out = torch.empty((B, C, *grid), dtype=torch.float32, device=Z.device)
        nbytes = lib.pdephi_plan_workspace_bytes(plan, B * C)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=Z.device)
        _cabi.check(lib.pdephi_synthesis_c2r(plan, _p(Z), _p(out), _p(a0), _p(a1), _p(mode_scale), _p(row_scale),
                                             rs_stride, B * C, direction, precision, _p(ws), nbytes, _stream()),
                    "synthesis_c2r")
    return out


def rational_mix_fwd(X, P, Q, monoP, monoQ, eps: float, mask, proj_eps: float, keep_transfer: bool = False):
    """K2.  Returns Y [B,O,mx,my,mzh] complex64 (unprojected), stats [B*O,4] = {a, b, s, 0} and, when
    ``keep_transfer``, the materialised R and Q+eps [O,I,mx,my,mzh] the backward streams."""
    B, I = X.shape[:2]
    O = P.shape[0]
    M = X[0, 0].numel()
    lib = _cabi.load()
    with torch.cuda.device(X.device):
        Y = torch.empty((B, O, *X.shape[2:]), dtype=torch.complex64, device=X.device)
        stats = torch.empty((B * O, 4), dtype=torch.float32, device=X.device)
        R = Qe = None
        if keep_transfer:
            R = torch.empty((O, I, *X.shape[2:]), dtype=torch.float32, device=X.device)
            Qe = torch.empty_like(R)
        _cabi.check(lib.pdephi_rational_mix_fwd(_p(X), _p(P), _p(Q), _p(monoP), _p(monoQ), P.shape[-1], Q.shape[-1],
                                                eps, _p(mask), proj_eps, _p(Y), _p(stats), _p(R), _p(Qe), B, I, O, M,
                                                _stream()), "rational_mix_fwd")
    return Y, stats, R, Qe


def projection_bwd(dZ, Y, mask, stats):
    rows = Y.shape[0] * Y.shape[1]
    M = Y[0, 0].numel()
    lib = _cabi.load()
    with torch.cuda.device(Y.device):
        dY = torch.empty_like(Y)
        _cabi.check(lib.pdephi_projection_bwd(_p(dZ), _p(Y), _p(mask), _p(stats), _p(dY), rows, M, _stream()),
                    "projection_bwd")
    return dY


def rational_mix_bwd(X, dY, R, Qe, monoP, monoQ, coeff_shapes):
    """Backward of K2 from the materialised transfer function; coeff_shapes = (P.shape, Q.shape)."""
    B, I = X.shape[:2]
    O = R.shape[0]
    M = X[0, 0].numel()
    p_shape, q_shape = coeff_shapes
    lib = _cabi.load()
    with torch.cuda.device(X.device):
        dX = torch.empty_like(X)
        dP = torch.empty(p_shape, dtype=torch.float32, device=X.device)
        dQ = torch.empty(q_shape, dtype=torch.float32, device=X.device)
        nbytes = lib.pdephi_rational_mix_bwd_workspace_bytes(p_shape[-1], q_shape[-1], I, O, M)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=X.device)
        _cabi.check(lib.pdephi_rational_mix_bwd(_p(X), _p(dY), _p(R), _p(Qe), _p(monoP), _p(monoQ), p_shape[-1],
                                                q_shape[-1], _p(dX), _p(dP), _p(dQ), B, I, O, M, _p(ws), nbytes,
                                                _stream()), "rational_mix_bwd")
    return dX, dP, dQ


def complex_mix_fwd(X, W, mzh: int):
    B, I = X.shape[:2]
    O = W.shape[0]
    mx, my, mz = W.shape[2:]
    lib = _cabi.load()
    with torch.cuda.device(X.device):
        Y = torch.empty((B, O, mx, my, mzh), dtype=torch.complex64, device=X.device)
        _cabi.check(lib.pdephi_complex_mix_fwd(_p(X), _p(W), _p(Y), B, I, O, mx * my, mzh, mz, _stream()),
                    "complex_mix_fwd")
    return Y


def complex_mix_bwd(X, W, dY, mzh: int):
    B, I = X.shape[:2]
    O = W.shape[0]
    mx, my, mz = W.shape[2:]
    lib = _cabi.load()
    with torch.cuda.device(X.device):
        dX = torch.empty_like(X)
        dW = torch.empty_like(W)
        _cabi.check(lib.pdephi_complex_mix_bwd(_p(X), _p(W), _p(dY), _p(dX), _p(dW), B, I, O, mx * my, mzh, mz,
                                               _stream()), "complex_mix_bwd")
    return dX, dW


def block_tail_fwd(h, spec, Wc, bc):
    """u = spec + (Wc h + bc) + h, hout = gelu(u) on [B,64,...] tensors (tcgen05 TF32 channel GEMM, fused epilogue)."""
    B, C = h.shape[:2]
    S = h[0, 0].numel()
    lib = _cabi.load()
    with torch.cuda.device(h.device):
        u = torch.empty_like(h)
        hout = torch.empty_like(h)
        _cabi.check(lib.pdephi_block_tail_fwd(_p(h), _p(spec), _p(Wc), _p(bc), _p(u), _p(hout), B, C, S, _stream()),
                    "block_tail_fwd")
    return u, hout


def block_tail_bwd(g, u, h, Wc, want_bias: bool):
    """gu = g gelu'(u), t = gu + Wc^T gu, dWc = sum gu (x) h, dbc = sum gu."""
    B, C = h.shape[:2]
    S = h[0, 0].numel()
    lib = _cabi.load()
    with torch.cuda.device(h.device):
        gu = torch.empty_like(h)
        t = torch.empty_like(h)
        dW = torch.empty((C, C), dtype=torch.float32, device=h.device)
        db = torch.empty(C, dtype=torch.float32, device=h.device) if want_bias else None
        nbytes = lib.pdephi_block_tail_bwd_workspace_bytes(C)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=h.device)
        _cabi.check(lib.pdephi_block_tail_bwd(_p(g), _p(u), _p(h), _p(Wc), _p(gu), _p(t), _p(dW), _p(db), B, C, S, _p(ws),
                                              nbytes, _stream()), "block_tail_bwd")
    return gu, t, dW, db


def block_tail_supported(h: torch.Tensor, precision: int) -> bool:
    """The fused block tail contracts the 1x1x1 convolution in single-pass TF32 on the tensor cores.  That is the
    arithmetic torch's own cuDNN convolution uses while ``torch.backends.cudnn.allow_tf32`` is True (the PyTorch
    default, so also what the reference's Conv3d does on this GPU): the 'tf32' route always takes it, the
    fp32-grade tensor-core routes take it exactly when torch would run the conv in TF32 too, 'fp32' never does."""
    if not (h.is_cuda and h.dtype == torch.float32 and h.dim() == 5 and h.shape[1] == 64 and h[0, 0].numel() % 128 == 0):
        return False
    if precision == _cabi.PREC_TF32:
        return True
    return precision in (_cabi.PREC_AUTO, _cabi.PREC_TF32X3) and bool(torch.backends.cudnn.allow_tf32)


# ------------------------------------------------------------------------------------------------
# autograd Functions
# ------------------------------------------------------------------------------------------------
def _slab_args(slab, x):
    """slab = None | (process_group, Nx_total): this rank holds x planes [rank*Nx, (rank+1)*Nx) of the grid."""
    if slab is None:
        return False, None, 1
    import torch.distributed as dist
    group, nx_total = slab
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nx = x.shape[2]
    if nx * world != nx_total:
        raise ValueError(f"slab decomposition needs Nx_total == world * local Nx, got {nx_total} vs {world} * {nx}")
    return (group if group is not None else dist.group.WORLD), (nx_total, rank * nx), world


def _allreduce_modes(X, group):
    """Sum the slabs' partial spectra: the one collective of the slab-decomposed transform (NCCL over NVLink)."""
    import torch.distributed as dist
    dist.all_reduce(torch.view_as_real(X), op=dist.ReduceOp.SUM, group=group)
    return X



class PrunedAnalysisFn(torch.autograd.Function):
    """K1 alone with autograd: X = rfftn(x)[..., :mx, :my, :mzh].  Backward is the ADJOINT synthesis."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, modes_h, precision):
        ctx.meta = (tuple(x.shape[2:]), precision)
        return analysis(x, tuple(modes_h), _cabi.FORWARD, precision)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    @once_differentiable
    def backward(ctx, gX):
        grid, precision = ctx.meta
        return synthesis(gX.contiguous(), grid, direction=_cabi.ADJOINT, precision=precision), None, None


class PrunedSynthesisFn(torch.autograd.Function):
    """K3 alone with autograd: out = irfftn(zero-pad(Z), s=grid).  Backward is the ADJOINT analysis."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, Z, grid, precision):
        ctx.meta = (tuple(Z.shape[2:]), precision)
        return synthesis(Z, tuple(grid), direction=_cabi.FORWARD, precision=precision)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    @once_differentiable
    def backward(ctx, g):
        modes_h, precision = ctx.meta
        return analysis(g.contiguous(), modes_h, _cabi.ADJOINT, precision), None, None


def spectral_resample(x: torch.Tensor, modes_h: Tuple[int, int, int], grid: Tuple[int, int, int],
                      precision: int = _cabi.PREC_AUTO) -> torch.Tensor:
    """irfftn(rfftn(x)[..., :mx, :my, :mzh] zero-padded to `grid`'s half-spectrum, s=grid): one K1 + K3 pair with
    different input and output grids.  Covers the reference's spectral down-sampling (rfno.py:119-130: modes = the whole
    coarse half-spectrum), up-sampling (rfno.py:132-149: modes = the whole input half-spectrum) and the 2/3-rule
    de-aliasing of the rollout loop (rational_fourier.py:392-412: same grid, modes = the retained box), without
    materialising a full spectrum."""
    return PrunedSynthesisFn.apply(PrunedAnalysisFn.apply(x, tuple(modes_h), precision), tuple(grid), precision)


class RationalSpectralFn(torch.autograd.Function):
    """out = irfftn(StabilityProjection(R(k) mix(rfftn(x)[corner]))) [+ skip0 + x].

    Saves only the retained modes (X, Y: 71 MB each at the 128^3 / width-64 / batch-8 config) and the
    transfer function (R, Q+eps: 285 MB each) instead of the reference's full spectra and polynomial graph
    (~19 GB per layer).  Backward formulas: SURVEY.md section 8a.
    """

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, P, Q, mask, monoP, monoQ, modes, eps, precision, skip0, residual, slab=None):
        for t, n in ((x, "x"), (P, "P_coeffs"), (Q, "Q_coeffs")):
            _require_cuda_f32(t, n)
        mx, my, mz = modes
        modes_h = (mx, my, mz // 2 + 1)
        grid = tuple(x.shape[2:])
        group, slab_plan, world = _slab_args(slab, x)
        P = P.contiguous()
        Q = Q.contiguous()
        X = analysis(x, modes_h, _cabi.FORWARD, precision, slab_plan)
        if group:
            _allreduce_modes(X, group)          # partial sums over the x slabs -> every rank holds the full modes
        need_grad = any(ctx.needs_input_grad[:3])
        Y, stats, R, Qe = rational_mix_fwd(X, P, Q, monoP, monoQ, eps, mask, eps, keep_transfer=need_grad)
        out = synthesis(Y, grid, mask, stats[:, 2], skip0, x if residual else None, _cabi.FORWARD, precision, slab_plan)
        if need_grad:
            ctx.save_for_backward(X, Y, stats, R, Qe, mask, monoP, monoQ)
        ctx.meta = (modes_h, grid, (tuple(P.shape), tuple(Q.shape)), precision, skip0 is not None, bool(residual),
                    group, slab_plan, world)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    @once_differentiable
    def backward(ctx, g):
        X, Y, stats, R, Qe, mask, monoP, monoQ = ctx.saved_tensors
        modes_h, grid, coeff_shapes, precision, has0, residual, group, slab_plan, world = ctx.meta
        g = g.contiguous()
        need_x, need_P, need_Q = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        dx = dP = dQ = None
        if need_x or need_P or need_Q:
            dZ = analysis(g, modes_h, _cabi.ADJOINT, precision, slab_plan)
            if group:
                _allreduce_modes(dZ, group)
            dY = projection_bwd(dZ, Y, mask, stats)
            dX, dP, dQ = rational_mix_bwd(X, dY, R, Qe, monoP, monoQ, coeff_shapes)
            if world > 1:
                # every rank computed the complete coefficient gradient from the replicated modes; pre-divide so the
                # SUM all-reduce that completes the slab-local gradients of the pointwise layers leaves it unchanged
                dP.div_(world)
                dQ.div_(world)
            if need_x:   # the residual branch's gradient (g itself) rides in the synthesis epilogue
                dx = synthesis(dX, grid, None, None, g if residual else None, None, _cabi.ADJOINT, precision, slab_plan)
        return (dx, dP if need_P else None, dQ if need_Q else None, None, None, None, None, None, None,
                g if has0 else None, None, None)


class ComplexSpectralFn(torch.autograd.Function):
    """out = irfftn(pad(einsum(rfftn(x)[corner], W[..., :mzh]))) [+ skip0 + x]  (SpectralConv3D)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, W, modes, precision, skip0, residual, slab=None):
        _require_cuda_f32(x, "x")
        _require_cuda_f32(W, "weights", complex_ok=True)
        mx, my, mz = modes
        modes_h = (mx, my, mz // 2 + 1)
        grid = tuple(x.shape[2:])
        group, slab_plan, world = _slab_args(slab, x)
        W = W.contiguous()
        X = analysis(x, modes_h, _cabi.FORWARD, precision, slab_plan)
        if group:
            _allreduce_modes(X, group)
        Y = complex_mix_fwd(X, W, modes_h[2])
        out = synthesis(Y, grid, None, None, skip0, x if residual else None, _cabi.FORWARD, precision, slab_plan)
        ctx.save_for_backward(X, W)
        ctx.meta = (modes_h, grid, precision, skip0 is not None, bool(residual), group, slab_plan, world)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    @once_differentiable
    def backward(ctx, g):
        X, W = ctx.saved_tensors
        modes_h, grid, precision, has0, residual, group, slab_plan, world = ctx.meta
        g = g.contiguous()
        dx = dW = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            dY = analysis(g, modes_h, _cabi.ADJOINT, precision, slab_plan)
            if group:
                _allreduce_modes(dY, group)
            dX, dW = complex_mix_bwd(X, W, dY, modes_h[2])
            if world > 1:
                dW.div_(world)
            if ctx.needs_input_grad[0]:
                dx = synthesis(dX, grid, None, None, g if residual else None, None, _cabi.ADJOINT, precision, slab_plan)
        return dx, dW if ctx.needs_input_grad[1] else None, None, None, g if has0 else None, None, None


class PointwiseLinearFn(torch.autograd.Function):
    """Channels-first pointwise linear map y[b,o,s] = bias[o] + sum_c W[o,c] x[b,c,s] (the models' lift / project).

    Equivalent to the reference's rearrange -> nn.Linear -> rearrange without the two transposed copies."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, weight, bias):
        _require_cuda_f32(x, "x")
        if x.dim() < 3 or weight.dim() != 2 or x.shape[1] != weight.shape[1]:
            raise RuntimeError(f"pointwise linear: input {tuple(x.shape)} does not match weight {tuple(weight.shape)}")
        x = x.contiguous()
        weight = weight.contiguous()
        B, Cin = x.shape[:2]
        Cout = weight.shape[0]
        S = x[0, 0].numel()
        lib = _cabi.load()
        with torch.cuda.device(x.device):
            out = torch.empty((B, Cout, *x.shape[2:]), dtype=torch.float32, device=x.device)
            _cabi.check(lib.pdephi_pointwise_linear(_p(x), _p(weight), _p(bias), _p(out), B, Cin, Cout, S, 0, _stream()),
                        "pointwise_linear")
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    @once_differentiable
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g = g.contiguous()
        B, Cin = x.shape[:2]
        Cout = weight.shape[0]
        S = x[0, 0].numel()
        lib = _cabi.load()
        dx = dW = db = None
        with torch.cuda.device(x.device):
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(x)
                _cabi.check(lib.pdephi_pointwise_linear(_p(g), _p(weight), _p(None), _p(dx), B, Cout, Cin, S, 1, _stream()),
                            "pointwise_linear(dx)")
            if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
                dW = torch.empty_like(weight)
                db = torch.empty(Cout, dtype=torch.float32, device=x.device) if ctx.has_bias else None
                nbytes = lib.pdephi_pointwise_wgrad_workspace_bytes(Cin, Cout, S)
                ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
                _cabi.check(lib.pdephi_pointwise_wgrad(_p(x), _p(g), _p(dW), _p(db), B, Cin, Cout, S, _p(ws), nbytes,
                                                       _stream()), "pointwise_wgrad")
        return dx, dW, db


def pointwise_linear_supported(x: torch.Tensor, weight: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() >= 3 and x[0, 0].numel() % 4 == 0
            and min(weight.shape[0], weight.shape[1]) <= 8)


class RationalBlockFn(torch.autograd.Function):
    """One whole block of RationalFourierOperator3D: h' = gelu(RationalFourierLayer(h) + Conv1x1(h) + h)
    (rational_fourier.py:263-271) as K1 -> K2 -> K3 -> fused block tail, with a hand-written backward that
    needs no autograd-side accumulation passes (the three branch gradients of h meet in K3's epilogue)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, h, P, Q, mask, monoP, monoQ, modes, eps, precision, Wc, bc, slab=None):
        for t, n in ((h, "x"), (P, "P_coeffs"), (Q, "Q_coeffs"), (Wc, "conv weight")):
            _require_cuda_f32(t, n)
        h = h.contiguous()
        mx, my, mz = modes
        modes_h = (mx, my, mz // 2 + 1)
        grid = tuple(h.shape[2:])
        group, slab_plan, world = _slab_args(slab, h)
        P, Q, Wc = P.contiguous(), Q.contiguous(), Wc.contiguous()
        X = analysis(h, modes_h, _cabi.FORWARD, precision, slab_plan)
        if group:
            _allreduce_modes(X, group)
        need_grad = any(ctx.needs_input_grad)
        Y, stats, R, Qe = rational_mix_fwd(X, P, Q, monoP, monoQ, eps, mask, eps, keep_transfer=need_grad)
        spec = synthesis(Y, grid, mask, stats[:, 2], None, None, _cabi.FORWARD, precision, slab_plan)
        u, hout = block_tail_fwd(h, spec, Wc, bc)
        if need_grad:
            ctx.save_for_backward(X, Y, stats, R, Qe, mask, monoP, monoQ, u, h, Wc)
        ctx.meta = (modes_h, grid, (tuple(P.shape), tuple(Q.shape)), precision, bc is not None, group, slab_plan, world)
        return hout

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    @once_differentiable
    def backward(ctx, g):
        X, Y, stats, R, Qe, mask, monoP, monoQ, u, h, Wc = ctx.saved_tensors
        modes_h, grid, coeff_shapes, precision, has_bias, group, slab_plan, world = ctx.meta
        gu, t, dWc, dbc = block_tail_bwd(g.contiguous(), u, h, Wc, has_bias)
        dZ = analysis(gu, modes_h, _cabi.ADJOINT, precision, slab_plan)
        if group:
            _allreduce_modes(dZ, group)
        dY = projection_bwd(dZ, Y, mask, stats)
        dX, dP, dQ = rational_mix_bwd(X, dY, R, Qe, monoP, monoQ, coeff_shapes)
        if world > 1:
            dP.div_(world)
            dQ.div_(world)
        dh = synthesis(dX, grid, None, None, t, None, _cabi.ADJOINT, precision, slab_plan)
        return (dh, dP, dQ, None, None, None, None, None, None, dWc.view_as(Wc), dbc, None)


class ComplexBlockFn(torch.autograd.Function):
    """One whole block of FNO3D: h' = gelu(SpectralConv3D(h) + Conv1x1(h) + h) (models/fno3d.py:91-97)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, h, W, modes, precision, Wc, bc, slab=None):
        _require_cuda_f32(h, "x")
        _require_cuda_f32(W, "weights", complex_ok=True)
        h = h.contiguous()
        mx, my, mz = modes
        modes_h = (mx, my, mz // 2 + 1)
        grid = tuple(h.shape[2:])
        group, slab_plan, world = _slab_args(slab, h)
        W, Wc = W.contiguous(), Wc.contiguous()
        X = analysis(h, modes_h, _cabi.FORWARD, precision, slab_plan)
        if group:
            _allreduce_modes(X, group)
        Y = complex_mix_fwd(X, W, modes_h[2])
        spec = synthesis(Y, grid, None, None, None, None, _cabi.FORWARD, precision, slab_plan)
        u, hout = block_tail_fwd(h, spec, Wc, bc)
        ctx.save_for_backward(X, W, u, h, Wc)
        ctx.meta = (modes_h, grid, precision, bc is not None, group, slab_plan, world)
        return hout

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    @once_differentiable
    def backward(ctx, g):
        X, W, u, h, Wc = ctx.saved_tensors
        modes_h, grid, precision, has_bias, group, slab_plan, world = ctx.meta
        gu, t, dWc, dbc = block_tail_bwd(g.contiguous(), u, h, Wc, has_bias)
        dY = analysis(gu, modes_h, _cabi.ADJOINT, precision, slab_plan)
        if group:
            _allreduce_modes(dY, group)
        dX, dW = complex_mix_bwd(X, W, dY, modes_h[2])
        if world > 1:
            dW.div_(world)
        dh = synthesis(dX, grid, None, None, t, None, _cabi.ADJOINT, precision, slab_plan)
        return dh, dW, None, None, dWc.view_as(Wc), dbc, None
