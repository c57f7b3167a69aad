"""Pin the oracle (oracle/) against the reference's own outputs stored in tests/golden.

CPU only.  The fixtures were produced by tests/golden/make_golden.py from the unmodified
reference modules; nothing here reads /root/reference.
"""
import numpy as np
import pytest
import torch

from oracle import dense_fp64 as d64
from oracle import spectral_oracle as so
from conftest import load_golden

RFL_CASES = [
    "rfl_c4_m4_n8_default",
    "rfl_c4_m4_n8_conditioned",
    "rfl_c3_m648_n161220_order43",
    "rfl_c2_m8_n8_nyquist_order22",
    "rfl_c8_m8_n16_default",
    "rfl_c2_m536_n9711_odd_order33",
]
SC_CASES = ["sc3d_c4_m4_n8_sliced", "sc3d_c3_m881_n16x16x1_unmodified", "sc3d_c2_m462_n8x12x6_unmodified"]


def _rel(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


@pytest.mark.parametrize("name", RFL_CASES)
def test_torch_restatement_matches_reference_fp32(name):
    """Same op order as the reference => same fp32 roundings: outputs and autograd grads agree to ~1 ulp."""
    z = load_golden(name)
    modes = tuple(int(v) for v in z["meta:modes"])
    x = torch.from_numpy(z["x"]).requires_grad_(True)
    P = torch.from_numpy(z["sd:P_coeffs"]).requires_grad_(True)
    Q = torch.from_numpy(z["sd:Q_coeffs"]).requires_grad_(True)
    out = so.rational_fourier_layer(x, P, Q, modes, float(z["meta:eps"]))
    out.backward(torch.from_numpy(z["g"]))
    assert _rel(out.detach().numpy(), z["out"]) < 1e-6
    assert _rel(x.grad.numpy(), z["dx"]) < 1e-6
    assert _rel(P.grad.numpy(), z["grad:P_coeffs"]) < 1e-6
    assert _rel(Q.grad.numpy(), z["grad:Q_coeffs"]) < 1e-6


@pytest.mark.parametrize("name", RFL_CASES)
def test_buffers_match_reference(name):
    z = load_golden(name)
    modes = tuple(int(v) for v in z["meta:modes"])
    assert np.array_equal(so.wavenumber_grid(modes).numpy(), z["sd:k_grid"])
    assert np.array_equal(so.decay_mask(modes).numpy(), z["sd:stability_projection.decay_mask"])
    assert _rel(d64.decay_mask(modes), z["sd:stability_projection.decay_mask"]) < 1e-6


@pytest.mark.parametrize("name", RFL_CASES)
def test_dense_fp64_matches_reference_fp64(name):
    """The stage-by-stage float64 spec (dense DFT + hand-derived backward) vs the reference's .double() run."""
    z = load_golden(name)
    modes = tuple(int(v) for v in z["meta:modes"])
    out, cache = d64.rational_layer_forward(z["x"], z["sd:P_coeffs"], z["sd:Q_coeffs"], modes, float(z["meta:eps"]),
                                            mask=z["sd:stability_projection.decay_mask"])
    dx, dP, dQ = d64.rational_layer_backward(z["g"], cache)
    assert _rel(out, z["out64"]) < 1e-9
    assert _rel(dx, z["dx64"]) < 1e-9
    assert _rel(dP, z["grad64:P_coeffs"]) < 1e-9
    assert _rel(dQ, z["grad64:Q_coeffs"]) < 1e-9
    # unused coefficient entries (a+b+c >= order) get exactly zero gradient, as in the reference
    assert np.count_nonzero(z["grad64:P_coeffs"][dP == 0]) == 0


@pytest.mark.parametrize("name", SC_CASES)
def test_spectral_conv_restatements(name):
    z = load_golden(name)
    modes = tuple(int(v) for v in z["meta:modes"])
    x = torch.from_numpy(z["x"]).requires_grad_(True)
    W = torch.from_numpy(z["sd:weights"]).requires_grad_(True)
    out = so.spectral_conv3d(x, W, modes)
    out.backward(torch.from_numpy(z["g"]))
    assert _rel(out.detach().numpy(), z["out"]) < 1e-6
    assert _rel(x.grad.numpy(), z["dx"]) < 1e-6
    assert _rel(W.grad.numpy(), z["grad:weights"]) < 1e-6
    # float64 spec against the reference's fp32 numbers (no fp64 reference run for the unmodified forward)
    o64, cache = d64.spectral_conv_forward(z["x"], z["sd:weights"], modes)
    dx64, dW64 = d64.spectral_conv_backward(z["g"], cache)
    ref_out = z["out64"] if "out64" in z else z["out"]
    tol = 1e-9 if "out64" in z else 2e-6
    assert _rel(o64, ref_out) < tol
    assert _rel(dx64, z["dx64"] if "dx64" in z else z["dx"]) < tol
    assert _rel(dW64, z["grad64:weights"] if "grad64:weights" in z else z["grad:weights"]) < tol


def test_rfno3d_model_restatement():
    z = load_golden("rfno3d_w8_m4_l2_n8")
    modes = tuple(int(v) for v in z["meta:modes"])
    sd = {k[3:]: torch.from_numpy(v).requires_grad_(v.dtype.kind in "fc") for k, v in z.items() if k.startswith("sd:")}
    x = torch.from_numpy(z["x"]).requires_grad_(True)
    out = so.rfno3d_forward(x, sd, modes, int(z["meta:n_layers"]))
    out.backward(torch.from_numpy(z["g"]))
    assert _rel(out.detach().numpy(), z["out"]) < 1e-6
    assert _rel(x.grad.numpy(), z["dx"]) < 1e-5
    for k, v in z.items():
        if k.startswith("grad:"):
            assert _rel(sd[k[5:]].grad.numpy(), v) < 1e-5, k


def test_fno3d_model_restatement():
    z = load_golden("fno3d_w8_m4_l2_n8_sliced")
    modes = tuple(int(v) for v in z["meta:modes"])
    sd = {k[3:]: torch.from_numpy(v).requires_grad_(True) for k, v in z.items() if k.startswith("sd:")}
    x = torch.from_numpy(z["x"]).requires_grad_(True)
    out = so.fno3d_forward(x, sd, modes, int(z["meta:n_layers"]))
    out.backward(torch.from_numpy(z["g"]))
    assert _rel(out.detach().numpy(), z["out"]) < 1e-6
    assert _rel(x.grad.numpy(), z["dx"]) < 1e-5
    for k, v in z.items():
        if k.startswith("grad:"):
            assert _rel(sd[k[5:]].grad.numpy(), v) < 1e-5, k


def test_R_is_bit_identical_to_reference_order():
    """H1: the stepwise fp32 evaluation order reproduces P, Q, R; an fp64-then-round evaluation does not."""
    torch.manual_seed(7)
    P, Q = so.init_rational_coeffs(4, 4)
    modes = (16, 16, 16)
    R, Pk, Qe = so.rational_transfer(P, Q, modes, 1e-6)
    # numpy stepwise fp32 emulation, independent of torch broadcasting
    mono, exps = d64.monomials(modes, 4)
    Pp, _ = d64.pack_coeffs(P.numpy())
    Qp, _ = d64.pack_coeffs(Q.numpy())
    accP = np.zeros(Pk.shape, np.float32)
    accQ = np.zeros(Pk.shape, np.float32)
    for t in range(len(exps)):
        m = mono[t].astype(np.float32)
        accP = (accP + (Pp[:, :, t][:, :, None, None, None] * m).astype(np.float32)).astype(np.float32)
        accQ = (accQ + (Qp[:, :, t][:, :, None, None, None] * m).astype(np.float32)).astype(np.float32)
    accQ = (accQ + np.float32(1e-6)).astype(np.float32)
    assert np.array_equal(accP, Pk.numpy())
    assert np.array_equal(accQ, Qe.numpy())
    assert np.array_equal((accP / accQ).astype(np.float32), R.numpy())
