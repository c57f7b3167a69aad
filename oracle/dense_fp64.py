"""TEST INFRASTRUCTURE ONLY -- float64 dense-DFT restatement with hand-derived backward.

This is the stage-by-stage specification the CUDA kernels implement:
pruned analysis transform A, per-mode channel mix with the rational transfer function,
StabilityProjection, pruned synthesis transform S, and the backward of each, written with
explicit DFT matrices in numpy float64 (no FFT library, no autograd).  It is checked against
the reference's own outputs and autograd gradients through tests/golden (see
tests/test_oracle_golden.py), and the -m gpu tests use it to localise a failure to one stage.

Reference lines restated: rational_fourier.py:73-172 (layer), stability.py:40-103
(projection), spectral_layers.py:46-78 (complex spectral conv), spectral_utils.py:16-45 (grid).
"""
from __future__ import annotations

import itertools
from typing import Tuple

import numpy as np

Modes = Tuple[int, int, int]


def twiddle(N: int, m: int, sign: float) -> np.ndarray:
    """F[k, n] = exp(sign * 2 pi i k n / N), k < m; the argument is reduced mod N exactly."""
    k = np.arange(m)[:, None]
    n = np.arange(N)[None, :]
    ang = 2.0 * np.pi * ((k * n) % N) / N
    return np.cos(ang) + 1j * sign * np.sin(ang)


def twiddle_slab(N_total: int, m: int, sign: float, offset: int, n_local: int) -> np.ndarray:
    """Rows [offset, offset+n_local) of twiddle(N_total, m, sign): the x twiddles of one slab."""
    return twiddle(N_total, m, sign)[:, offset:offset + n_local]


def analysis_slab(x_slab: np.ndarray, modes_h: Modes, nx_total: int, x_offset: int) -> np.ndarray:
    """Partial pruned spectrum contributed by the x slab [x_offset, x_offset+Nx_local): summing it over the
    slabs of a volume gives analysis(volume) (linearity of the x-stage sum)."""
    mx, my, mzh = modes_h
    nxl, Ny, Nz = x_slab.shape[-3:]
    t = np.einsum("...xyz,kz->...xyk", x_slab, twiddle(Nz, mzh, -1.0))
    t = np.einsum("...xyk,jy->...xjk", t, twiddle(Ny, my, -1.0))
    return np.einsum("...xjk,ix->...ijk", t, twiddle_slab(nx_total, mx, -1.0, x_offset, nxl))


def hermitian_weights(Nz: int, mzh: int) -> np.ndarray:
    """c_kz of the C2R step of irfftn: 1 at kz=0 and at the Nyquist bin, 2 elsewhere.

    torch.fft.irfftn (rational_fourier.py:170) ignores the imaginary part of the kz=0 and
    kz=Nz/2 bins; Re(c_kz * Z * e^{i...}) reproduces that exactly.
    """
    c = np.full(mzh, 2.0)
    c[0] = 1.0
    if Nz % 2 == 0 and mzh - 1 == Nz // 2:
        c[-1] = 1.0
    return c


def analysis(x: np.ndarray, modes_h: Modes) -> np.ndarray:
    """Pruned rfftn: X[..., kx, ky, kz] for kx<mx, ky<my, kz<mzh (rational_fourier.py:161 + :107)."""
    mx, my, mzh = modes_h
    Nx, Ny, Nz = x.shape[-3:]
    t = np.einsum("...xyz,kz->...xyk", x, twiddle(Nz, mzh, -1.0))
    t = np.einsum("...xyk,jy->...xjk", t, twiddle(Ny, my, -1.0))
    return np.einsum("...xjk,ix->...ijk", t, twiddle(Nx, mx, -1.0))


def synthesis(Z: np.ndarray, grid: Modes, hermitian: bool = True, normalise: bool = True) -> np.ndarray:
    """Zero-padded irfftn of a low-mode corner (rational_fourier.py:116-117,170).

    hermitian/normalise False gives the plain adjoint of ``analysis``.
    """
    Nx, Ny, Nz = grid
    mx, my, mzh = Z.shape[-3:]
    c = hermitian_weights(Nz, mzh) if hermitian else np.ones(mzh)
    t = np.einsum("...ijk,ix->...xjk", Z, twiddle(Nx, mx, +1.0))
    t = np.einsum("...xjk,jy->...xyk", t, twiddle(Ny, my, +1.0))
    out = np.einsum("...xyk,kz->...xyz", t * c, twiddle(Nz, mzh, +1.0)).real
    return out / (Nx * Ny * Nz) if normalise else out


def monomials(modes: Modes, order: int) -> np.ndarray:
    """mono[t, kx, ky, kz] = kx^a ky^b kz^c, (a,b,c) in the reference's loop order with a+b+c<order."""
    mx, my, mz = modes
    kx, ky, kz = np.meshgrid(np.arange(mx, dtype=np.float64), np.arange(my, dtype=np.float64),
                             np.arange(mz // 2 + 1, dtype=np.float64), indexing="ij")
    exps = [e for e in itertools.product(range(order), repeat=3) if sum(e) < order]
    return np.stack([kx ** a * ky ** b * kz ** c for a, b, c in exps]), exps


def pack_coeffs(C: np.ndarray):
    """[O,I,p,p,p] -> [O,I,T] picking the used entries, reference loop order."""
    order = C.shape[-1]
    exps = [e for e in itertools.product(range(order), repeat=3) if sum(e) < order]
    return np.stack([C[:, :, a, b, c] for a, b, c in exps], axis=-1), exps


def decay_mask(modes: Modes, decay_rate: float = 2.0) -> np.ndarray:
    mx, my, mz = modes
    kx, ky, kz = np.meshgrid(np.arange(mx, dtype=np.float64), np.arange(my, dtype=np.float64),
                             np.arange(mz // 2 + 1, dtype=np.float64), indexing="ij")
    return (1.0 + kx ** 2 + ky ** 2 + kz ** 2) ** (-decay_rate / 2)


def rational_layer_forward(x, P, Q, modes: Modes, eps: float = 1e-6, mask=None):
    """RationalFourierLayer.forward in float64; returns (out, cache for backward).

    ``mask``: the module's fp32 ``decay_mask`` buffer (a ``.double()`` reference module keeps the
    fp32-computed values); defaults to the float64 formula.
    """
    x = np.asarray(x, np.float64)
    P = np.asarray(P, np.float64)
    Q = np.asarray(Q, np.float64)
    mx, my, mz = modes
    mzh = mz // 2 + 1
    grid = x.shape[-3:]
    X = analysis(x, (mx, my, mzh))                                   # [B,I,mx,my,mzh]
    Pp, _ = pack_coeffs(P)
    Qp, _ = pack_coeffs(Q)
    monoP, _ = monomials(modes, P.shape[-1])
    monoQ, _ = monomials(modes, Q.shape[-1])
    Pk = np.einsum("oit,txyz->oixyz", Pp, monoP)
    Qe = np.einsum("oit,txyz->oixyz", Qp, monoQ) + eps
    R = Pk / Qe
    Y = np.einsum("bixyz,oixyz->boxyz", X, R)
    m = decay_mask(modes) if mask is None else np.asarray(mask, np.float64)
    a = np.sum(np.abs(Y) ** 2, axis=(-3, -2, -1)) + eps              # [B,O]
    b = np.sum((m * np.abs(Y)) ** 2, axis=(-3, -2, -1)) + eps
    s = np.sqrt(a / b)
    Z = Y * m * s[..., None, None, None]
    out = synthesis(Z, grid)
    cache = dict(X=X, R=R, Qe=Qe, Y=Y, m=m, a=a, b=b, s=s, grid=grid, modes=modes,
                 monoP=monoP, monoQ=monoQ, Pshape=P.shape, Qshape=Q.shape)
    return out, cache


def rational_layer_backward(g, cache):
    """Hand-derived backward: returns dx, dP, dQ (full [O,I,p,p,p] shapes)."""
    g = np.asarray(g, np.float64)
    X, R, Qe, Y, m = cache["X"], cache["R"], cache["Qe"], cache["Y"], cache["m"]
    a, b, s = cache["a"], cache["b"], cache["s"]
    Nx, Ny, Nz = cache["grid"]
    mx, my, mz = cache["modes"]
    mzh = mz // 2 + 1
    c = hermitian_weights(Nz, mzh)
    dZ = analysis(g, (mx, my, mzh)) * c / (Nx * Ny * Nz)            # backward of S is A with a diagonal scale
    G = np.sum((np.conj(dZ) * m * Y).real, axis=(-3, -2, -1))       # [B,O]
    s_ = s[..., None, None, None]
    dY = m * s_ * dZ + (G * s)[..., None, None, None] * (1.0 / a[..., None, None, None]
                                                         - m ** 2 / b[..., None, None, None]) * Y
    dX = np.einsum("oixyz,boxyz->bixyz", R, dY)
    dR = np.einsum("bixyz,boxyz->oixyz", np.conj(X), dY).real
    dPp = np.einsum("oixyz,txyz->oit", dR / Qe, cache["monoP"])
    dQp = -np.einsum("oixyz,txyz->oit", dR * R / Qe, cache["monoQ"])
    dx = synthesis(dX, (Nx, Ny, Nz), hermitian=False, normalise=False)
    dP = np.zeros(cache["Pshape"])
    dQ = np.zeros(cache["Qshape"])
    for t, (i, j, k) in enumerate(pack_coeffs(np.zeros(cache["Pshape"]))[1]):
        dP[:, :, i, j, k] = dPp[:, :, t]
    for t, (i, j, k) in enumerate(pack_coeffs(np.zeros(cache["Qshape"]))[1]):
        dQ[:, :, i, j, k] = dQp[:, :, t]
    return dx, dP, dQ


def spectral_conv_forward(x, W, modes: Modes):
    """SpectralConv3D.forward, oracle convention 2 (weight sliced to kz < mz//2+1)."""
    x = np.asarray(x, np.float64)
    W = np.asarray(W, np.complex128)
    mx, my, mz = modes
    mzh = mz // 2 + 1
    X = analysis(x, (mx, my, mzh))
    Y = np.einsum("bixyz,oixyz->boxyz", X, W[..., :mzh])
    return synthesis(Y, x.shape[-3:]), dict(X=X, W=W, grid=x.shape[-3:], modes=modes)


def spectral_conv_backward(g, cache):
    """dx and dW (torch complex-gradient convention: dW = sum_b conj(X) dY), zeros for kz >= mzh."""
    g = np.asarray(g, np.float64)
    X, W = cache["X"], cache["W"]
    Nx, Ny, Nz = cache["grid"]
    mx, my, mz = cache["modes"]
    mzh = mz // 2 + 1
    c = hermitian_weights(Nz, mzh)
    dY = analysis(g, (mx, my, mzh)) * c / (Nx * Ny * Nz)
    dX = np.einsum("oixyz,boxyz->bixyz", np.conj(W[..., :mzh]), dY)
    dW = np.zeros_like(W)
    dW[..., :mzh] = np.einsum("bixyz,boxyz->oixyz", np.conj(X), dY)
    dx = synthesis(dX, (Nx, Ny, Nz), hermitian=False, normalise=False)
    return dx, dW
