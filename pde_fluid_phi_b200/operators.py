"""Drop-in mirrors of the reference's spectral operator modules.

Same class names, constructor signatures, parameter / buffer names, shapes and RNG consumption
order as the reference (so ``torch.manual_seed`` reproduces its weights and ``state_dict``s
interchange), with ``forward`` rerouted to the sm_100a kernels behind ``pde_fluid_phi_b200.functional``.

  RationalFourierLayer   <- operators/rational_fourier.py:19-172
  StabilityProjection    <- operators/stability.py:14-103
  SpectralConv3D         <- operators/spectral_layers.py:17-78
  get_grid               <- utils/spectral_utils.py:16-45
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch
import torch.nn as nn

from . import _cabi, compile_ops
from .functional import (ComplexBlockFn, ComplexSpectralFn, RationalBlockFn, RationalSpectralFn, TransferCache, block_tail_supported,
                         library_call, monomial_exponents, monomial_table)

_default_precision = "auto"


def set_default_precision(precision: str) -> None:
    """Arithmetic of the transform stages.

    'auto'   (default) fp32-grade, parity gate 1e-5: split-TF32 tcgen05 contractions ('tf32x3') on the shapes
             the tensor-core kernels cover, FFMA ('fp32') everywhere else
    'fp32'   FFMA contractions, any shape
    'tf32x3' tcgen05, every operand as a (hi, lo) TF32 pair, three MMAs per product
    'tf32'   tcgen05 single-pass TF32, parity gate 2e-3; with width 64 the whole block runs fused
    """
    global _default_precision
    if precision not in _cabi.PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(_cabi.PRECISIONS)}")
    _default_precision = precision


def get_default_precision() -> str:
    return _default_precision


def get_grid(modes: Tuple[int, int, int], device: Union[str, torch.device] = "cpu",
             dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Integer wavenumbers of the retained box, [3, mx, my, mz//2+1] (kx, ky, kz planes)."""
    mx, my, mz = modes
    axes = [torch.arange(n, device=device, dtype=dtype) for n in (mx, my, mz // 2 + 1)]
    return torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=0)


def _lifted_ok(layer, x: torch.Tensor, conv: nn.Module, lift: nn.Module) -> bool:
    from .functional import block_tail_supported_shape, lift_supported, pointwise_linear_supported
    return (not compile_ops.compiling() and layer.slab is None and isinstance(lift, nn.Linear) and isinstance(x, torch.Tensor) and x.dim() == 5
            and x.shape[1] == lift.in_features and lift.out_features == layer.in_channels
            and pointwise_linear_supported(x, lift.weight) and lift_supported(lift)
            and isinstance(conv, nn.Conv3d) and conv.kernel_size == (1, 1, 1) and conv.groups == 1
            and conv.in_channels == conv.out_channels == layer.in_channels == layer.out_channels
            and block_tail_supported_shape(x.is_cuda, x.dtype, lift.out_features, x[0, 0].numel(), layer._precision_code())
            and all(m <= n for m, n in zip(layer.modes[:2], x.shape[2:4])) and layer.modes[2] // 2 + 1 <= x.shape[4] // 2 + 1)


def _check_input(x: torch.Tensor, channels: int, modes, who: str, slab=None):
    if not isinstance(x, torch.Tensor) or x.dim() != 5:
        raise ValueError(f"{who}: expected a [batch, channels, Nx, Ny, Nz] tensor, got "
                         f"{tuple(x.shape) if isinstance(x, torch.Tensor) else type(x)}")
    if x.shape[1] != channels:
        raise RuntimeError(f"{who}: input has {x.shape[1]} channels, layer was built for {channels}")
    mx, my, mz = modes
    nx, ny, nz = x.shape[2:]
    if slab is not None:
        nx = slab[1]                      # x is one slab of a grid with Nx_total planes
    if mx > nx or my > ny or mz // 2 + 1 > nz // 2 + 1:
        raise RuntimeError(f"{who}: modes {tuple(modes)} do not fit the half-spectrum of a {(nx, ny, nz)} grid")


class StabilityProjection(nn.Module):
    """Decay mask (1+|k|^2)^(-decay_rate/2) plus per-(batch, channel) energy-conserving rescale.

    Inside RationalFourierLayer.forward the projection is fused into the mix / synthesis kernels.
    Calling this module directly (a full-spectrum tensor in, full-spectrum tensor out, as the
    reference's adaptive subclass does) runs the same arithmetic as plain tensor ops.
    """

    def __init__(self, modes: Tuple[int, int, int], decay_rate: float = 2.0, eps: float = 1e-6,
                 energy_conserving: bool = True):
        super().__init__()
        self.modes = modes
        self.decay_rate = decay_rate
        self.eps = eps
        self.energy_conserving = energy_conserving
        self.register_buffer("decay_mask", self._create_decay_mask())

    def _create_decay_mask(self) -> torch.Tensor:
        kx, ky, kz = get_grid(self.modes)
        k_mag = torch.sqrt(kx ** 2 + ky ** 2 + kz ** 2)
        return (1 + k_mag ** 2) ** (-self.decay_rate / 2)

    def forward(self, x_ft: torch.Tensor) -> torch.Tensor:
        mx, my, mz = self.modes
        box = (slice(None), slice(None), slice(0, mx), slice(0, my), slice(0, mz // 2 + 1))
        kept = x_ft[box]
        damped = kept * self.decay_mask.to(x_ft.device)
        if self.energy_conserving:
            damped = self._enforce_energy_conservation(kept, damped)
        out = x_ft.clone()
        out[box] = damped
        return out

    def _enforce_energy_conservation(self, x_original: torch.Tensor, x_projected: torch.Tensor) -> torch.Tensor:
        dims = (-3, -2, -1)
        e0 = x_original.abs().pow(2).sum(dim=dims, keepdim=True)
        e1 = x_projected.abs().pow(2).sum(dim=dims, keepdim=True)
        return x_projected * torch.sqrt((e0 + self.eps) / (e1 + self.eps))


class RationalFourierLayer(nn.Module):
    """R(k) = P(k)/Q(k) spectral layer; forward runs K1 -> K2 -> K3 of libpdephi_b200.so."""

    def __init__(self, in_channels: int, out_channels: int, modes: Tuple[int, int, int],
                 rational_order: Tuple[int, int] = (4, 4), stability_eps: float = 1e-6, init_scale: float = 0.02):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.modes = modes
        self.rational_order = rational_order
        if not _cabi.rational_order_supported(tuple(rational_order)):
            raise NotImplementedError(f"rational_order {tuple(rational_order)}: the kernels are instantiated for every (p, q) "
                                      "with p, q in 2..4 and for (1, 1), (5, 5), (6, 6)")
        self.stability_eps = stability_eps
        self.precision: Optional[str] = None     # None -> module-level default
        self.slab = None                         # (process_group, Nx_total) for the slab-decomposed transform
        # R(k), Q(k)+eps of the last forward, reused while the coefficients are unchanged (functional.TransferCache); a plain
        # attribute: not a parameter, not a buffer, not in the state_dict
        self._transfer_cache = TransferCache()
        self._init_rational_coefficients(init_scale)
        self.stability_projection = StabilityProjection(modes=modes, decay_rate=2.0, eps=stability_eps)
        self.register_buffer("k_grid", get_grid(modes, device="cpu"))
        # kernel-side constants; non-persistent so state_dict keys equal the reference's
        self.register_buffer("_mono_p", monomial_table(modes, rational_order[0]), persistent=False)
        self.register_buffer("_mono_q", monomial_table(modes, rational_order[1]), persistent=False)

    def _init_rational_coefficients(self, scale: float):
        p, q = self.rational_order
        self.P_coeffs = nn.Parameter(torch.randn(self.out_channels, self.in_channels, p, p, p) * scale)
        self.Q_coeffs = nn.Parameter(torch.randn(self.out_channels, self.in_channels, q, q, q) * scale)
        with torch.no_grad():
            self.Q_coeffs[..., 0, 0, 0] = self.Q_coeffs[..., 0, 0, 0].abs() + 1.0

    # ---- torch-op methods kept for callers that use them directly on full spectra ---------------
    def _evaluate_polynomial(self, coeffs: torch.Tensor, k_x: torch.Tensor, k_y: torch.Tensor,
                             k_z: torch.Tensor) -> torch.Tensor:
        total = torch.zeros(coeffs.shape[:2] + k_x.shape, device=coeffs.device, dtype=coeffs.dtype)
        for a, b, c in monomial_exponents(coeffs.shape[-1]):
            total = total + coeffs[:, :, a, b, c, None, None, None] * ((k_x ** a) * (k_y ** b) * (k_z ** c))
        return total

    def rational_multiply(self, x_ft: torch.Tensor) -> torch.Tensor:
        mx, my, mz = self.modes
        box = (slice(None), slice(None), slice(0, mx), slice(0, my), slice(0, mz // 2 + 1))
        kx, ky, kz = self.k_grid.to(x_ft.device)
        transfer = self._evaluate_polynomial(self.P_coeffs, kx, ky, kz) / (
            self._evaluate_polynomial(self.Q_coeffs, kx, ky, kz) + self.stability_eps)
        kept = x_ft[box]
        out = torch.zeros_like(x_ft)
        out[box] = torch.einsum("bixyz,oixyz->boxyz", kept, transfer.to(kept.dtype))
        return out

    # ---- the hot path ---------------------------------------------------------------------------
    def _precision_code(self) -> int:
        return _cabi.PRECISIONS[self.precision or _default_precision]

    def _projection(self):
        """(eps, energy_conserving) of the stability_projection submodule, which the fused kernels implement
        (the reference calls the submodule, rational_fourier.py:167, so its attributes are honoured here)."""
        proj = self.stability_projection
        if type(proj) is not StabilityProjection:
            raise NotImplementedError(
                f"RationalFourierLayer.forward fuses the stock StabilityProjection into its kernels; the layer's "
                f"stability_projection is a {type(proj).__name__}.  Call rational_multiply / stability_projection (torch ops on "
                "full spectra) directly for a custom projection.")
        return (float(proj.eps), bool(proj.energy_conserving))

    def _run(self, x, skip0, residual):
        _check_input(x, self.in_channels, self.modes, "RationalFourierLayer", self.slab)
        if self.in_channels != self.out_channels:
            raise RuntimeError("RationalFourierLayer needs in_channels == out_channels (as the reference does)")
        if self.slab is None and compile_ops.compiling():      # torch.compile: one opaque custom op, no graph break
            return compile_ops.rational_spectral(x, self.P_coeffs, self.Q_coeffs, self.stability_projection.decay_mask,
                                                 self._mono_p, self._mono_q, tuple(self.modes), float(self.stability_eps),
                                                 self._precision_code(), skip0, residual, self._projection())
        return library_call(RationalSpectralFn, x, self.P_coeffs, self.Q_coeffs, self.stability_projection.decay_mask,
                                        self._mono_p, self._mono_q, tuple(self.modes), float(self.stability_eps),
                                        self._precision_code(), skip0, residual, self.slab, self._projection(), self._transfer_cache)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._run(x, None, False)

    def forward_fused(self, x: torch.Tensor, x_local: torch.Tensor) -> torch.Tensor:
        """layer(x) + x_local + x with both adds fused into the synthesis kernel's epilogue
        (the block body of rational_fourier.py:263-271 minus the activation)."""
        return self._run(x, x_local, True)

    def block_supported(self, x: torch.Tensor, conv: nn.Module) -> bool:
        # (not while torch.compile traces: the fused block is an eager route; a traced model runs the spectral layer as a custom
        # op with both adds in its epilogue and leaves the convolution and the activation to the compiler)
        return (not compile_ops.compiling() and isinstance(conv, nn.Conv3d) and conv.kernel_size == (1, 1, 1) and conv.groups == 1
                and conv.in_channels == conv.out_channels == self.in_channels == self.out_channels
                and block_tail_supported(x, self._precision_code()))

    def lifted_block_supported(self, x: torch.Tensor, conv: nn.Module, lift: nn.Module) -> bool:
        """Can the first block take the model input `x` and the lift with it? (the block route must apply to lift(x))"""
        return _lifted_ok(self, x, conv, lift)

    def forward_block(self, x: torch.Tensor, conv: nn.Conv3d, head: Optional[nn.Linear] = None,
                      lift: Optional[nn.Linear] = None, vz_in=None, want_vz: bool = False):
        """gelu(layer(x) + conv(x) + x): the whole block of rational_fourier.py:263-271 in the library
        (TF32 path, width 64): spectral layer, then the 1x1x1 convolution as a tcgen05 channel GEMM whose
        epilogue applies bias, both residual adds and the exact GELU.  `head`: the model's output projection, applied
        to the result inside the same kernel (the last block of the model): returns head(gelu(...)) channels-first.
        `lift`: the model's input lift; `x` is then the model input and the block starts from lift(x) (the first block:
        its analysis and, when x needs no gradient, the lift's backward run in mode space -- functional._lift_forward).
        `want_vz`: returns (result, vz) with vz the z stage of the forward analysis of the result, formed inside the block kernel
        (None where that route does not apply); `vz_in`: that tensor from the previous block -- this block's analysis starts from it."""
        if lift is None:
            _check_input(x, self.in_channels, self.modes, "RationalFourierLayer", self.slab)
        return library_call(RationalBlockFn, x, self.P_coeffs, self.Q_coeffs, self.stability_projection.decay_mask,
                            self._mono_p, self._mono_q, tuple(self.modes), float(self.stability_eps),
                            self._precision_code(), conv.weight, conv.bias, self.slab,
                            head.weight if head is not None else None, head.bias if head is not None else None,
                            lift.weight if lift is not None else None, lift.bias if lift is not None else None,
                            self._projection(), self._transfer_cache, vz_in, bool(want_vz))


class SpectralConv3D(nn.Module):
    """Classic FNO spectral convolution with a complex [O, I, mx, my, mz] weight.

    The einsum contracts against ``weights[..., :mz//2+1]`` (oracle convention 2 of SURVEY.md
    section 8c; the unmodified reference forward raises for mz > 2); entries with kz >= mz//2+1 receive zero
    gradient.  For mz in {1, 2} this is exactly the reference.
    """

    def __init__(self, in_channels: int, out_channels: int, modes: Tuple[int, int, int], init_scale: float = 0.02):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.modes = modes
        self.precision: Optional[str] = None
        self.slab = None
        self.weights = nn.Parameter(torch.randn(out_channels, in_channels, *modes, dtype=torch.cfloat) * init_scale)

    def _precision_code(self) -> int:
        return _cabi.PRECISIONS[self.precision or _default_precision]

    def _run(self, x, skip0, residual):
        _check_input(x, self.in_channels, self.modes, "SpectralConv3D", self.slab)
        if residual and self.in_channels != self.out_channels:
            raise RuntimeError("fused residual needs in_channels == out_channels")
        if self.slab is None and compile_ops.compiling():
            return compile_ops.complex_spectral(x, self.weights, tuple(self.modes), self._precision_code(), skip0, residual)
        return library_call(ComplexSpectralFn, x, self.weights, tuple(self.modes), self._precision_code(), skip0, residual,
                                       self.slab)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._run(x, None, False)

    def forward_fused(self, x: torch.Tensor, x_local: torch.Tensor) -> torch.Tensor:
        """conv(x) + x_local + x, adds fused into the synthesis epilogue (models/fno3d.py:91-97)."""
        return self._run(x, x_local, True)

    def block_supported(self, x: torch.Tensor, conv: nn.Module) -> bool:
        return (not compile_ops.compiling() and isinstance(conv, nn.Conv3d) and conv.kernel_size == (1, 1, 1) and conv.groups == 1
                and conv.in_channels == conv.out_channels == self.in_channels == self.out_channels
                and block_tail_supported(x, self._precision_code()))

    def lifted_block_supported(self, x: torch.Tensor, conv: nn.Module, lift: nn.Module) -> bool:
        return _lifted_ok(self, x, conv, lift)

    def forward_block(self, x: torch.Tensor, conv: nn.Conv3d, head: Optional[nn.Linear] = None,
                      lift: Optional[nn.Linear] = None, vz_in=None, want_vz: bool = False):
        """gelu(spectral_conv(x) + conv(x) + x): the whole block of models/fno3d.py:91-97 in the library; `head`, `lift`, `vz_in`,
        `want_vz` as in RationalFourierLayer.forward_block."""
        if lift is None:
            _check_input(x, self.in_channels, self.modes, "SpectralConv3D", self.slab)
        return library_call(ComplexBlockFn, x, self.weights, tuple(self.modes), self._precision_code(), conv.weight, conv.bias,
                            self.slab, head.weight if head is not None else None, head.bias if head is not None else None,
                            lift.weight if lift is not None else None, lift.bias if lift is not None else None, vz_in, bool(want_vz))
