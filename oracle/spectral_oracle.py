"""TEST INFRASTRUCTURE ONLY -- torch-CPU restatement of the reference spectral layers.

Functional (no nn.Module) restatement of the reference's hot path, op for op, so that
fp32 roundings happen in the same places.  Every function cites the reference lines it
follows (paths relative to /root/reference/src/pde_fluid_phi).  Gradients are obtained
by autograd through these functions, which is how the reference obtains them.

Pinned against the reference's own outputs by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import itertools
from typing import Dict, Sequence, Tuple

import torch
import torch.nn.functional as F

Modes = Tuple[int, int, int]


def wavenumber_grid(modes: Modes, dtype=torch.float32) -> torch.Tensor:
    """Integer wavenumber grid [3, mx, my, mz//2+1] (utils/spectral_utils.py:16-45)."""
    mx, my, mz = modes
    axes = [torch.arange(n, dtype=dtype) for n in (mx, my, mz // 2 + 1)]
    return torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=0)


def decay_mask(modes: Modes, decay_rate: float = 2.0) -> torch.Tensor:
    """(1+|k|^2)^(-decay_rate/2) on the retained box (operators/stability.py:40-57)."""
    kx, ky, kz = wavenumber_grid(modes)
    k_mag = torch.sqrt(kx ** 2 + ky ** 2 + kz ** 2)
    return (1 + k_mag ** 2) ** (-decay_rate / 2)


def monomial_exponents(order: int):
    """(a,b,c) with a+b+c < order in the reference's loop order (rational_fourier.py:135-138)."""
    return [e for e in itertools.product(range(order), repeat=3) if sum(e) < order]


def evaluate_polynomial(coeffs: torch.Tensor, kx, ky, kz) -> torch.Tensor:
    """sum_t coeffs[o,i,a,b,c] * kx^a ky^b kz^c, term by term (rational_fourier.py:121-148).

    The accumulation starts from zeros and adds one rounded product per monomial, as the
    reference's eager ops do; this ordering is what makes R(k) reproducible bit for bit.
    """
    order = coeffs.shape[-1]
    acc = torch.zeros(coeffs.shape[:2] + kx.shape, dtype=coeffs.dtype)
    for a, b, c in monomial_exponents(order):
        mono = (kx ** a) * (ky ** b) * (kz ** c)
        acc = acc + coeffs[:, :, a, b, c][:, :, None, None, None] * mono
    return acc


def rational_transfer(P: torch.Tensor, Q: torch.Tensor, modes: Modes, eps: float):
    """Real transfer function R = P(k) / (Q(k) + eps) (rational_fourier.py:92-104)."""
    kx, ky, kz = wavenumber_grid(modes, dtype=P.dtype)
    p_k = evaluate_polynomial(P, kx, ky, kz)
    q_k = evaluate_polynomial(Q, kx, ky, kz) + eps
    return p_k / q_k, p_k, q_k


def rational_multiply(x_ft: torch.Tensor, P, Q, modes: Modes, eps: float) -> torch.Tensor:
    """Channel mix with R(k) on the low-mode corner, zero elsewhere (rational_fourier.py:73-119)."""
    mx, my, mz = modes
    mzh = mz // 2 + 1
    R, _, _ = rational_transfer(P, Q, modes, eps)
    corner = x_ft[:, :, :mx, :my, :mzh]
    mixed = torch.einsum("bixyz,oixyz->boxyz", corner, R.to(corner.dtype))
    out = torch.zeros_like(x_ft)
    out[:, :, :mx, :my, :mzh] = mixed
    return out


def stability_projection(x_ft: torch.Tensor, modes: Modes, eps: float, decay_rate: float = 2.0):
    """Decay mask + per-(b,c) energy-conserving rescale (operators/stability.py:59-103)."""
    mx, my, mz = modes
    mzh = mz // 2 + 1
    mask = decay_mask(modes, decay_rate).to(x_ft.real.dtype)
    corner = x_ft[:, :, :mx, :my, :mzh]
    damped = corner * mask
    e_in = torch.sum(torch.abs(corner) ** 2, dim=(-3, -2, -1), keepdim=True)
    e_out = torch.sum(torch.abs(damped) ** 2, dim=(-3, -2, -1), keepdim=True)
    damped = damped * torch.sqrt((e_in + eps) / (e_out + eps))
    out = x_ft.clone()
    out[:, :, :mx, :my, :mzh] = damped
    return out


def rational_fourier_layer(x: torch.Tensor, P, Q, modes: Modes, eps: float = 1e-6) -> torch.Tensor:
    """RationalFourierLayer.forward (rational_fourier.py:150-172)."""
    x_ft = torch.fft.rfftn(x, dim=[-3, -2, -1])
    y_ft = rational_multiply(x_ft, P, Q, modes, eps)
    y_ft = stability_projection(y_ft, modes, eps)
    return torch.fft.irfftn(y_ft, s=x.shape[-3:], dim=[-3, -2, -1])


def spectral_conv3d(x: torch.Tensor, W: torch.Tensor, modes: Modes) -> torch.Tensor:
    """SpectralConv3D.forward (operators/spectral_layers.py:46-78), oracle convention 2.

    The unmodified reference einsum needs the weight's z extent (modes[2]) to equal the
    sliced spectrum's (modes[2]//2+1) and raises otherwise; the oracle convention (SURVEY.md
    section 8c) contracts against ``W[..., :modes[2]//2+1]``.  For modes[2] in {1, 2} this is
    the unmodified reference.
    """
    mx, my, mz = modes
    mzh = mz // 2 + 1
    x_ft = torch.fft.rfftn(x, dim=[-3, -2, -1])
    corner = x_ft[:, :, :mx, :my, :mzh]
    mixed = torch.einsum("bixyz,oixyz->boxyz", corner, W[..., :mzh])
    out_ft = torch.zeros(x.shape[0], W.shape[0], *x_ft.shape[-3:], dtype=x_ft.dtype)
    out_ft[:, :, :mx, :my, :mzh] = mixed
    return torch.fft.irfftn(out_ft, s=x.shape[-3:], dim=[-3, -2, -1])


def _lift(x, weight, bias):
    # rearrange 'b c h w d -> b h w d c', Linear, rearrange back (rational_fourier.py:258-260)
    return F.linear(x.permute(0, 2, 3, 4, 1), weight, bias).permute(0, 4, 1, 2, 3)


def rfno3d_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], modes: Modes, n_layers: int,
                   eps: float = 1e-6) -> torch.Tensor:
    """RationalFourierOperator3D.forward (rational_fourier.py:247-282) from a state_dict."""
    h = _lift(x, sd["input_proj.weight"], sd["input_proj.bias"])
    for i in range(n_layers):
        spec = rational_fourier_layer(h, sd[f"rational_layers.{i}.P_coeffs"],
                                      sd[f"rational_layers.{i}.Q_coeffs"], modes, eps)
        local = F.conv3d(h, sd[f"local_convs.{i}.weight"], sd[f"local_convs.{i}.bias"])
        h = F.gelu(spec + local + h)
    return _lift(h, sd["output_proj.weight"], sd["output_proj.bias"])


def fno3d_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], modes: Modes, n_layers: int) -> torch.Tensor:
    """FNO3D.forward (models/fno3d.py:75-108) from a state_dict, oracle convention 2."""
    h = _lift(x, sd["input_proj.weight"], sd["input_proj.bias"])
    for i in range(n_layers):
        spec = spectral_conv3d(h, sd[f"spectral_layers.{i}.weights"], modes)
        local = F.conv3d(h, sd[f"local_layers.{i}.weight"], sd[f"local_layers.{i}.bias"])
        h = F.gelu(spec + local + h)
    return _lift(h, sd["output_proj.weight"], sd["output_proj.bias"])


# ----------------------------------------------------------------------------------------
# Parameter construction with the reference's RNG order, for tests that need weights
# without importing the reference (rational_fourier.py:59-71, spectral_layers.py:42-44).
# ----------------------------------------------------------------------------------------
def init_rational_coeffs(channels_out: int, channels_in: int, rational_order: Sequence[int] = (4, 4),
                         scale: float = 0.02):
    p, q = rational_order
    P = torch.randn(channels_out, channels_in, p, p, p) * scale
    Q = torch.randn(channels_out, channels_in, q, q, q) * scale
    Q[..., 0, 0, 0] = torch.abs(Q[..., 0, 0, 0]) + 1.0
    return P, Q


def init_spectral_weights(channels_out: int, channels_in: int, modes: Modes, scale: float = 0.02):
    return torch.randn(channels_out, channels_in, *modes, dtype=torch.cfloat) * scale


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b|| over the whole tensor (complex ok)."""
    a = torch.as_tensor(a).detach()
    b = torch.as_tensor(b).detach()
    num = torch.linalg.vector_norm((a - b).reshape(-1).to(torch.complex128 if a.is_complex() else torch.float64))
    den = torch.linalg.vector_norm(b.reshape(-1).to(torch.complex128 if b.is_complex() else torch.float64))
    return float(num / den) if float(den) > 0 else float(num)
