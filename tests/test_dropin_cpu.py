"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every declared
symbol, the mirrored modules reproduce the reference's parameters / state_dict under the same
seed, the kept torch-op methods agree with the oracle, and CPU tensors are refused loudly."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import pde_fluid_phi_b200 as pb
from pde_fluid_phi_b200 import _cabi
from pde_fluid_phi_b200.functional import monomial_table
from oracle import dense_fp64 as d64
from oracle import spectral_oracle as so
from conftest import ROOT, load_golden


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "pdephi_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(pdephi_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 12
    lib = ctypes.CDLL(_cabi.lib_path())
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/pdephi_b200.h but not exported"
    assert sorted(_cabi.SIGNATURES) == declared        # the ctypes table mirrors the header one to one
    assert pb.load_library().pdephi_version() == 100


def test_argument_errors_without_gpu():
    lib = pb.load_library()
    # null-pointer / size validation happens before any CUDA call
    rc = lib.pdephi_rational_mix_fwd(None, None, None, None, None, 4, 4, 1e-6, None, 1e-6, None, None, None, None, 1, 1, 1, 1, None)
    assert rc == _cabi.ERR_INVALID
    assert b"null pointer" in lib.pdephi_last_error()
    assert lib.pdephi_plan_workspace_bytes(None, 4) == 0
    assert lib.pdephi_rational_mix_bwd_workspace_bytes(4, 4, 64, 64, 17408) > 0


@pytest.mark.parametrize("name,factory", [
    ("rfl_c4_m4_n8_default", lambda: pb.RationalFourierLayer(4, 4, (4, 4, 4))),
    ("rfl_c3_m648_n161220_order43", lambda: pb.RationalFourierLayer(3, 3, (6, 4, 8), rational_order=(4, 3))),
    ("sc3d_c4_m4_n8_sliced", lambda: pb.SpectralConv3D(4, 4, (4, 4, 4))),
    ("rfno3d_w8_m4_l2_n8", lambda: pb.RationalFourierOperator3D(modes=(4, 4, 4), width=8, n_layers=2, rational_order=(3, 3))),
    ("fno3d_w8_m4_l2_n8_sliced", lambda: pb.FNO3D(modes=(4, 4, 4), width=8, n_layers=2)),
])
def test_same_seed_same_state_dict_as_reference(name, factory):
    z = load_golden(name)
    torch.manual_seed(1234)
    module = factory()
    sd = module.state_dict()
    ref_keys = sorted(k[3:] for k in z if k.startswith("sd:"))
    assert sorted(sd.keys()) == ref_keys
    for k in ref_keys:
        assert np.array_equal(sd[k].numpy(), z["sd:" + k]), k


def test_state_dict_round_trip_from_reference_checkpoint():
    z = load_golden("rfno3d_w8_m4_l2_n8")
    torch.manual_seed(99)
    model = pb.RationalFourierOperator3D(modes=(4, 4, 4), width=8, n_layers=2, rational_order=(3, 3))
    model.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("sd:")}, strict=True)
    assert np.array_equal(model.rational_layers[1].Q_coeffs.detach().numpy(), z["sd:rational_layers.1.Q_coeffs"])


def test_cpu_tensors_are_refused():
    layer = pb.RationalFourierLayer(2, 2, (4, 4, 4))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        layer(torch.randn(1, 2, 8, 8, 8))
    conv = pb.SpectralConv3D(2, 2, (4, 4, 4))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        conv(torch.randn(1, 2, 8, 8, 8))
    with pytest.raises(RuntimeError):          # channel mismatch, as the reference (tests/test_integration.py:415-427)
        layer(torch.randn(1, 3, 8, 8, 8))
    with pytest.raises(ValueError):
        layer(torch.randn(2, 8, 8, 8))
    with pytest.raises(RuntimeError):          # grid smaller than modes
        layer(torch.randn(1, 2, 4, 4, 4)[:, :, :2])


def test_kept_torch_methods_match_oracle():
    """rational_multiply / stability_projection on full spectra (called directly by a reference subclass)."""
    z = load_golden("rfl_c4_m4_n8_default")
    layer = pb.RationalFourierLayer(4, 4, (4, 4, 4))
    layer.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("sd:")})
    x_ft = torch.fft.rfftn(torch.from_numpy(z["x"]), dim=[-3, -2, -1])
    with torch.no_grad():
        got = layer.stability_projection(layer.rational_multiply(x_ft))
        P, Q = torch.from_numpy(z["sd:P_coeffs"]), torch.from_numpy(z["sd:Q_coeffs"])
        want = so.stability_projection(so.rational_multiply(x_ft, P, Q, (4, 4, 4), 1e-6), (4, 4, 4), 1e-6)
        out = torch.fft.irfftn(got, s=(8, 8, 8), dim=[-3, -2, -1])
    assert torch.equal(got, want)
    assert so.rel_l2(out, torch.from_numpy(z["out"])) < 1e-6


def test_monomial_table_is_exact():
    modes = (6, 4, 8)
    mono64, exps = d64.monomials(modes, 4)
    t = monomial_table(modes, 4).numpy()
    assert t.shape == (20, 6 * 4 * 5)
    assert np.array_equal(t.astype(np.float64), mono64.reshape(20, -1))


def test_rational_fno_reproduces_reference_parameters():
    """N4: same constructor / sub-module names / RNG order as models/rfno.py -> same weights under the same seed,
    and the reference's state_dict loads strictly."""
    z = load_golden("rfno_model_w8_m4_l2_n16_multiscale")
    torch.manual_seed(1234)                                   # the seed tests/golden/make_golden.py used
    model = pb.RationalFNO(modes=(4, 4, 4), width=8, n_layers=2, rational_order=(3, 3), multi_scale=True)
    sd = model.state_dict()
    ref = {k[3:]: v for k, v in z.items() if k.startswith("sd:")}
    assert sorted(sd) == sorted(ref)
    for k, v in ref.items():
        assert np.array_equal(sd[k].numpy(), v), k
    assert model.coarse_modes == (2, 2, 2) and model.coarse_processor.width == 4
    with pytest.raises(RuntimeError):                         # CPU tensors are refused, never silently computed
        model(torch.zeros(1, 3, 16, 16, 16))


def test_rollout_fixture_is_self_consistent():
    """N3 fixture sanity on CPU: frame 0 is the initial condition, and the reference's 2/3 de-aliasing equals the fp64
    spec of a pruned analysis / synthesis pair on the [0, 2N/3) box (the mask is not Hermitian on the kz = 0 plane, so
    this also pins the irfftn convention of dropping that plane's imaginary residue)."""
    z = load_golden("rollout_rfno3d_w8_m4_l2_n12")
    assert np.array_equal(z["traj_adaptive"][:, 0], z["x"]) and np.array_equal(z["traj_fixed"][:, 0], z["x"])
    want = d64.synthesis(d64.analysis(z["x"].astype(np.float64), (8, 8, 4)), (12, 12, 12))
    assert so.rel_l2(torch.from_numpy(z["dealias"]).double(), torch.from_numpy(np.asarray(want))) < 1e-6


def test_rational_orders_the_kernels_cover():
    """Every (p, q) in 2..4 and the symmetric orders 1, 5, 6 construct (same parameter shapes as the reference for any
    order, rational_fourier.py:59-71); anything else is refused at construction, not at the first forward."""
    for order in [(2, 2), (4, 3), (3, 4), (1, 1), (5, 5), (6, 6)]:
        layer = pb.RationalFourierLayer(2, 2, (4, 4, 4), rational_order=order)
        assert layer.P_coeffs.shape == (2, 2, order[0], order[0], order[0])
        assert layer.Q_coeffs.shape == (2, 2, order[1], order[1], order[1])
        assert layer._mono_p.shape[0] == order[0] * (order[0] + 1) * (order[0] + 2) // 6
    for order in [(1, 3), (5, 4), (7, 7), (0, 2)]:
        with pytest.raises(NotImplementedError, match="rational_order"):
            pb.RationalFourierLayer(2, 2, (4, 4, 4), rational_order=order)


def test_fused_route_guards_are_host_side():
    """The lift the first block can absorb is capped at 4 input channels (MAXP_IN of block_tc.cu) and the projection at 4
    outputs: checked on the host so a 5..8-channel input takes the separate lift pass instead of a kernel error."""
    from pde_fluid_phi_b200.functional import head_supported, lift_supported
    assert lift_supported(torch.nn.Linear(3, 64)) is False      # CPU weights: never (the GPU cases are in test_parity_gpu.py)
    assert head_supported(torch.nn.Linear(64, 3)) is False
    layer = pb.RationalFourierLayer(4, 4, (4, 4, 4))

    class Custom(pb.StabilityProjection):
        pass
    layer.stability_projection = Custom((4, 4, 4))
    with pytest.raises(NotImplementedError, match="stability_projection"):
        layer._projection()
    layer.stability_projection = pb.StabilityProjection((4, 4, 4), eps=1e-3, energy_conserving=False)
    assert layer._projection() == (1e-3, False)


@pytest.mark.parametrize("cls", ["rfno", "fno"])
def test_models_trace_as_one_graph_under_torch_compile(cls):
    """The reference wraps its models in torch.compile (optimization/performance_optimization.py:466-486).  While dynamo
    traces, the mirror modules route their spectral layers through the pdephi_b200:: custom ops (compile_ops.py): the whole
    model is ONE graph (fullgraph=True: no graph break), forward and backward.  Traced here on meta tensors -- no GPU needed:
    the ops' fake implementations supply the shapes."""
    import pde_fluid_phi_b200 as pb
    torch.manual_seed(0)
    kw = dict(modes=(4, 4, 4), width=8, n_layers=2)
    model = (pb.RationalFourierOperator3D(**kw) if cls == "rfno" else pb.FNO3D(**kw)).to("meta")
    graphs = []

    def backend(gm, example_inputs):
        graphs.append(gm)
        return gm.forward

    compiled = torch.compile(model, backend=backend, fullgraph=True)
    out = compiled(torch.empty(2, 3, 8, 8, 8, device="meta"))
    out.sum().backward()
    assert tuple(out.shape) == (2, 3, 8, 8, 8) and len(graphs) == 1
    targets = [str(n.target) for n in graphs[0].graph.nodes if n.op == "call_function"]
    want = "rational_spectral_fwd" if cls == "rfno" else "complex_spectral_fwd"
    assert sum(want in t for t in targets) == 2, targets                       # one opaque op per spectral layer
    assert all(p.grad is not None and p.grad.shape == p.shape for p in model.parameters())


def test_transfer_cache_is_not_copied_or_pickled():
    """The layer's transfer-function cache (functional.TransferCache: R(k), Q(k)+eps of the last forward, hundreds of MB on
    the device) is not part of a deepcopy or a pickle of the layer, and never part of the state_dict."""
    import copy
    import pickle
    import pde_fluid_phi_b200 as pb
    layer = pb.RationalFourierLayer(2, 2, (4, 4, 4))
    layer._transfer_cache.R = torch.zeros(3)            # stand-in for a filled cache
    layer._transfer_cache.key = ("k",)
    assert not any("transfer" in k or k in ("R", "Qe") for k in layer.state_dict())
    for clone in (copy.deepcopy(layer), pickle.loads(pickle.dumps(layer))):
        assert clone._transfer_cache is not layer._transfer_cache
        assert clone._transfer_cache.key is None and clone._transfer_cache.R is None
        assert torch.equal(clone.P_coeffs, layer.P_coeffs)


def test_packed_block_derivative_layout_round_trips():
    """The host-side view of PDEPHI_OPT_BLOCK_SAVES_DERIVATIVE = 2 (block_tc.cu pack_du): one 32-bit word per channel pair and
    point, low half = even channel; pack / unpack are inverse up to the binary16 rounding and keep the logical [B,C,...] shape."""
    from pde_fluid_phi_b200 import functional as pf
    torch.manual_seed(0)
    du = torch.rand(2, 64, 4, 8, 16) * 1.26 - 0.13           # the range of gelu'
    packed = pf.pack_block_derivative(du)
    assert packed.dtype == torch.float16 and packed.shape == du.shape and packed.element_size() * packed.numel() == du.numel() * 2
    words = packed.reshape(2, 32, -1, 2)
    assert torch.equal(words[1, 5, 77, 0], du[1, 10].reshape(-1)[77].half()) and torch.equal(words[1, 5, 77, 1], du[1, 11].reshape(-1)[77].half())
    back = pf.unpack_block_derivative(packed)
    assert back.dtype == torch.float32 and torch.equal(back, du.half().float())
    assert float((back - du).abs().max()) <= 2.0 ** -11 * 1.13 and pf.unpack_block_derivative(du) is du


def test_block_derivative_option_governs_the_u_buffer():
    """PDEPHI_OPT_BLOCK_SAVES_DERIVATIVE (host-side state of the library, no GPU needed): default 2 = packed binary16 buffer of
    half the bytes; 0 / 1 = an fp32 activation; a buffer written under one setting is refused under another."""
    from pde_fluid_phi_b200 import functional as pf
    lib = _cabi.load()
    assert lib.pdephi_get_option(_cabi.OPT_BLOCK_SAVES_DERIVATIVE) == 2
    u2 = pf._block_u_empty(2, 64, (4, 4, 128), "cpu")
    assert u2.dtype == torch.float16 and tuple(u2.shape) == (2, 64, 4, 4, 128)
    pf._check_block_u(u2, "test")
    old = _cabi.set_option(_cabi.OPT_BLOCK_SAVES_DERIVATIVE, 1)
    try:
        assert old == 2
        u1 = pf._block_u_empty(2, 64, (4, 4, 128), "cpu")
        assert u1.dtype == torch.float32 and u1.numel() * u1.element_size() == 2 * u2.numel() * u2.element_size()
        pf._check_block_u(u1, "test")
        with pytest.raises(RuntimeError, match="SAVES_DERIVATIVE"):
            pf._check_block_u(u2, "test")
    finally:
        _cabi.set_option(_cabi.OPT_BLOCK_SAVES_DERIVATIVE, old)
    with pytest.raises(RuntimeError, match="SAVES_DERIVATIVE"):
        pf._check_block_u(u1, "test")
    assert lib.pdephi_get_option(_cabi.OPT_BLOCK_SAVES_DERIVATIVE) == 2
