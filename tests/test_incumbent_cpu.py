"""The stock-PyTorch incumbent that scripts/incumbent_torch_cufft.py times on the GPU is the reference: on the CPU it
reproduces the reference's own whole-operator fixture (tests/golden/make_golden.py) bit for bit."""
import importlib.util
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_script():
    spec = importlib.util.spec_from_file_location("incumbent_torch_cufft", os.path.join(ROOT, "scripts", "incumbent_torch_cufft.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_incumbent_restatement_reproduces_the_reference_fixture():
    inc = _load_script()
    d = np.load(os.path.join(ROOT, "tests", "golden", "rfno3d_w8_m4_l2_n8.npz"))
    sd = {k[3:].replace("stability_projection.decay_mask", "decay_mask"): torch.from_numpy(d[k]) for k in d.keys() if k.startswith("sd:")}
    order = sd["rational_layers.0.P_coeffs"].shape[-1], sd["rational_layers.0.Q_coeffs"].shape[-1]
    model = inc.EagerOperator((4, 4, 4), 8, 2, order)
    model.load_state_dict(sd, strict=True)              # buffers included: k_grid and decay_mask are the reference's
    x = torch.from_numpy(d["x"]).requires_grad_(True)
    out = model(x)
    out.backward(torch.from_numpy(d["g"]))
    assert torch.equal(out.detach(), torch.from_numpy(d["out"]))
    assert torch.equal(x.grad, torch.from_numpy(d["dx"]))
    fresh = inc.EagerOperator((4, 4, 4), 8, 2, order)   # ... and the buffers it builds itself are the same
    assert torch.equal(fresh.rational_layers[0].decay_mask, sd["rational_layers.0.decay_mask"])
    assert torch.equal(fresh.rational_layers[0].k_grid, sd["rational_layers.0.k_grid"])


def test_incumbent_spectral_conv_reproduces_the_reference_fixtures():
    inc = _load_script()
    for name in ("sc3d_c4_m4_n8_sliced", "sc3d_c3_m881_n16x16x1_unmodified", "sc3d_c2_m462_n8x12x6_unmodified"):
        d = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        W = torch.from_numpy(d["sd:weights"])
        layer = inc.EagerSpectralConv(W.shape[1], W.shape[0], tuple(int(v) for v in d["meta:modes"]))
        with torch.no_grad():
            layer.weights.copy_(W)
        x = torch.from_numpy(d["x"]).requires_grad_(True)
        out = layer(x)
        out.backward(torch.from_numpy(d["g"]))
        assert torch.equal(out.detach(), torch.from_numpy(d["out"])), name
        assert torch.equal(x.grad, torch.from_numpy(d["dx"])), name
        assert torch.equal(layer.weights.grad, torch.from_numpy(d["grad:weights"])), name


def test_incumbent_fno_reproduces_the_reference_fixture():
    inc = _load_script()
    d = np.load(os.path.join(ROOT, "tests", "golden", "fno3d_w8_m4_l2_n8_sliced.npz"))
    sd = {k[3:]: torch.from_numpy(d[k]) for k in d.keys() if k.startswith("sd:")}
    model = inc.EagerFNO((4, 4, 4), 8, 2)
    model.load_state_dict(sd, strict=True)
    x = torch.from_numpy(d["x"]).requires_grad_(True)
    out = model(x)
    out.backward(torch.from_numpy(d["g"]))
    assert torch.equal(out.detach(), torch.from_numpy(d["out"]))
    assert torch.equal(x.grad, torch.from_numpy(d["dx"]))
