"""The BASELINE.json configurations as parity cases: cfg 1 at its exact size, cfg 3 / cfg 4 at
sizes the CPU oracle finishes in seconds plus a full-size property check, all through the public modules."""
import math

import pytest
import torch
import torch.nn.functional as F

import pde_fluid_phi_b200 as pb
from pde_fluid_phi_b200 import _cabi
from pde_fluid_phi_b200 import functional as pf
from oracle import spectral_oracle as so

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    return so.rel_l2(a.detach().cpu(), b.detach().cpu())


def strict_fp32(fn):
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False       # torch's own 1x1x1 conv in fp32 for the 1e-5 whole-model gate
    try:
        return fn()
    finally:
        torch.backends.cudnn.allow_tf32 = old


def taylor_green(n: int, amplitudes):
    """u = sin X cos Y cos Z, v = -cos X sin Y cos Z, w = 0 on linspace(0, 2 pi, n) (endpoint included)."""
    g = torch.linspace(0, 2 * math.pi, n)
    X, Y, Z = torch.meshgrid(g, g, g, indexing="ij")
    u = torch.stack([torch.sin(X) * torch.cos(Y) * torch.cos(Z), -torch.cos(X) * torch.sin(Y) * torch.cos(Z), torch.zeros_like(X)])
    return torch.stack([a * u for a in amplitudes])


def compare_model(model, oracle_fn, x, y, out_tol=1e-5, grad_tol=5e-5, precision=None, spectral_grad_tol=None):
    """fwd + MSE + bwd on the GPU model and on the CPU oracle with the same state_dict."""
    if precision is not None:
        model.set_precision(precision)
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point or v.is_complex()) for k, v in model.state_dict().items()}
    out_o = oracle_fn(x, sd)
    loss_o = F.mse_loss(out_o, y)
    loss_o.backward()
    model = model.to(DEV)

    def run():
        out = model(x.to(DEV))
        loss = F.mse_loss(out, y.to(DEV))
        loss.backward()
        return out, loss
    out, loss = strict_fp32(run)
    assert rel(out, out_o) < out_tol
    assert abs(float(loss) - float(loss_o)) <= max(1e-5, 2 * out_tol) * abs(float(loss_o))
    for n, p in model.named_parameters():
        assert p.grad is not None, n
        g_o = sd[n].grad
        if float(g_o.abs().max()) == 0.0:
            assert float(p.grad.abs().max()) == 0.0, n
        else:
            spectral = any(k in n for k in ("P_coeffs", "Q_coeffs", "weights"))
            tol = spectral_grad_tol if (spectral and spectral_grad_tol is not None) else grad_tol
            assert rel(p.grad, g_o) < tol, (n, rel(p.grad, g_o))
    return float(loss)


def test_cfg1_fno3d_taylor_green_exact_size():
    """configs[0]: FNO3D modes (8,8,8) width 32, 4 layers, Taylor-Green 32^3, batch 4, fp32 fwd+bwd."""
    torch.manual_seed(1234)
    model = pb.FNO3D(modes=(8, 8, 8), width=32, n_layers=4)
    x = taylor_green(32, [1.0, 0.5, 0.25, 0.125])
    y = x * math.exp(-2 * 1e-3 * 1.0)
    compare_model(model, lambda xx, sd: so.fno3d_forward(xx, sd, (8, 8, 8), 4), x, y)


def test_cfg3_2d_kolmogorov_scaled_and_full_size():
    """configs[2]: 2-D FNO modes (64,64,1) on a 256 x 256 x 1 grid.  Against the oracle at width 16 / batch 2; at the
    full width 64 / batch 64 the transforms are checked through adjointness (size-independent)."""
    torch.manual_seed(1234)
    n = 256
    model = pb.FNO3D(modes=(64, 64, 1), width=16, n_layers=2, in_channels=2, out_channels=2)
    g = torch.Generator().manual_seed(42)
    yy = torch.linspace(0, 2 * math.pi, n)
    base = torch.sin(4 * yy)[None, :, None].expand(n, n, 1)
    x = torch.stack([torch.stack([base + 0.1 * torch.randn(n, n, 1, generator=g), 0.1 * torch.randn(n, n, 1, generator=g)])
                     for _ in range(2)])
    y = torch.roll(x, 3, dims=2)
    compare_model(model, lambda xx, sd: so.fno3d_forward(xx, sd, (64, 64, 1), 2), x, y)
    # full size: <A x, Z> == <x, A^H Z> on [64, 64, 256, 256, 1]
    torch.manual_seed(1)
    xf = torch.randn(64, 64, n, n, 1, device=DEV)
    Z = torch.randn(64, 64, 64, 64, 1, device=DEV, dtype=torch.complex64)
    lhs = (pf.analysis(xf, (64, 64, 1)).conj().to(torch.complex128) * Z.to(torch.complex128)).sum().real
    rhs = (xf.double() * pf.synthesis(Z, (n, n, 1), direction=_cabi.ADJOINT).double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-5 * abs(float(lhs))


def test_cfg3_full_width_model_block_route_vs_fp32_route():
    """configs[2] at its full width (64) and grid, 4 layers, batch 2: on a 2-D grid the width-64 model takes the block
    route (lift in the first block, 1x1x1 convolution + residual + GELU on the tensor cores, projection in the last
    block; per-axis transforms).  Outputs and every parameter gradient against the fp32 route of the same model, whose
    layers the width-16 case above pins to the oracle."""
    torch.manual_seed(1234)
    n = 256
    model = pb.FNO3D(modes=(64, 64, 1), width=64, n_layers=4, in_channels=2, out_channels=2).to(DEV)
    g = torch.Generator().manual_seed(42)
    yy = torch.linspace(0, 2 * math.pi, n)
    base = torch.sin(4 * yy)[None, :, None].expand(n, n, 1)
    x = torch.stack([torch.stack([base + 0.1 * torch.randn(n, n, 1, generator=g), 0.1 * torch.randn(n, n, 1, generator=g)])
                     for _ in range(2)]).to(DEV)
    y = torch.roll(x, 3, dims=2)

    def run(precision):
        model.set_precision(precision)
        model.zero_grad(set_to_none=True)
        out = model(x)
        torch.nn.functional.mse_loss(out, y).backward()
        return out.detach().clone(), {k: p.grad.clone() for k, p in model.named_parameters()}

    assert model.spectral_layers[0].block_supported(torch.empty(2, 64, n, n, 1, device=DEV), model.local_layers[0])
    out_tc, g_tc = run("tf32")
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        out_ref, g_ref = run("fp32")
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert torch.isfinite(out_tc).all()
    assert rel(out_tc, out_ref) < 3e-3
    for k in g_ref:
        assert rel(g_tc[k], g_ref[k]) < 5e-3, k
        assert g_tc[k].dtype == g_ref[k].dtype and g_tc[k].shape == g_ref[k].shape


@pytest.mark.parametrize("modes,order", [((16, 16, 16), (2, 2)), ((32, 32, 32), (3, 3)), ((64, 64, 64), (4, 4))])
def test_cfg4_multiscale_operators_on_128_cubed(modes, order):
    """configs[3] (oracle convention 3): the large / medium / small RationalFourierOperator3D(in = out = width) of
    MultiScaleFNO on a 128^3 grid, each against the oracle, at width 4 / one layer so the CPU side stays in seconds.
    Conditioned weights: the default init has poles of R(k) inside a 64^3 box."""
    torch.manual_seed(1234)
    w = 4
    model = pb.RationalFourierOperator3D(modes=modes, width=w, n_layers=1, in_channels=w, out_channels=w, rational_order=order)
    with torch.no_grad():
        layer = model.rational_layers[0]
        q0 = layer.Q_coeffs[..., 0, 0, 0].clone()
        layer.Q_coeffs.mul_(1e-6)
        layer.Q_coeffs[..., 0, 0, 0] = q0
    g = torch.Generator().manual_seed(42)
    x = torch.randn(1, w, 128, 128, 128, generator=g)
    y = torch.randn(1, w, 128, 128, 128, generator=g)
    weight = {16: 1.0, 32: 2.0, 64: 4.0}[modes[0]]          # MultiScaleLoss weights (multiscale_fno.py:49-54)
    # bias / pointwise-weight gradients are sums of 2M sign-alternating fp32 terms on a white-noise target (torch's own
    # reductions on both sides, different summation orders): 2e-4 on gradients here, 1e-5 on the outputs
    loss = compare_model(model, lambda xx, sd: so.rfno3d_forward(xx, sd, modes, 1), x, y, grad_tol=2e-4)
    assert math.isfinite(weight * loss)


CFG4_SCALES = {"large": ((16, 16, 16), 2, (2, 2), 1.0), "medium": ((32, 32, 32), 3, (3, 3), 2.0),
               "small": ((64, 64, 64), 4, (4, 4), 4.0)}     # modes, n_layers, rational_order, MultiScaleLoss weight


def _cfg4_operator(scale: str, width: int = 64):
    """One of MultiScaleFNO's default scale operators (multiscale_fno.py:93-127: in_channels = out_channels = width) with
    conditioned weights (Q(k) kept away from its poles, P(k) scaled so the spectral branch is level
    with the other two) and visible biases."""
    modes, n_layers, order, _ = CFG4_SCALES[scale]
    model = pb.RationalFourierOperator3D(modes=modes, width=width, n_layers=n_layers, in_channels=width, out_channels=width,
                                         rational_order=order)
    kmax = float(modes[0] - 1)
    with torch.no_grad():
        for layer in model.rational_layers:
            q0 = layer.Q_coeffs[..., 0, 0, 0].clone()
            layer.Q_coeffs.mul_(1e-6)
            layer.Q_coeffs[..., 0, 0, 0] = q0
            layer.P_coeffs.mul_(2.0 / (width * 0.02 * max(1.0, kmax ** (order[0] - 1))))
        for m in list(model.local_convs) + [model.input_proj, model.output_proj]:
            m.bias.normal_(0.0, 0.1)
    return model


def _band_limited(channels: int, n: int, kmax: float, seed: int) -> torch.Tensor:
    """[1, channels, n, n, n] random fields with a k^-5/3 energy spectrum over 1 <= |k| <= kmax, unit variance."""
    g = torch.Generator().manual_seed(seed)
    k1 = torch.fft.fftfreq(n, 1.0 / n)
    kx, ky, kz = torch.meshgrid(k1, k1, torch.fft.rfftfreq(n, 1.0 / n), indexing="ij")
    kmag = torch.sqrt(kx ** 2 + ky ** 2 + kz ** 2)
    amp = torch.where((kmag >= 1) & (kmag <= kmax), kmag.clamp(min=1.0) ** (-11.0 / 6.0), torch.zeros_like(kmag))
    out = torch.empty(1, channels, n, n, n)
    for c in range(channels):
        noise = torch.randn(2, n, n, n // 2 + 1, generator=g)
        u = torch.fft.irfftn(torch.complex(noise[0], noise[1]) * amp, s=(n, n, n))
        out[0, c] = u / u.std()
    return out


def test_cfg4_three_scale_operators_full_width_weighted_step():
    """configs[3] assembled: the large / medium / small operators of MultiScaleFNO at their full width 64 on one 128^3 sample,
    the scale-weighted loss (multiscale_fno.py:47-55: 1 / 2 / 4) summed and back-propagated in one step.  Every operator runs
    lift (64 -> 64, pointwise_wide.cu) -> tcgen05 block route -> projection in the library; the LARGE operator's output and
    every parameter gradient are checked against the CPU oracle at the full width (2e-3, the TF32 gate: the blocks'
    1x1x1 convolution runs in TF32 like torch's own with cudnn.allow_tf32)."""
    torch.manual_seed(4321)
    ops = {s: _cfg4_operator(s) for s in CFG4_SCALES}
    # band-limited random fields (k^-5/3 over 1 <= |k| <= 16, the large scale's own band).  With a WHITE-noise target the
    # loss gradient is white too, 99.8 % of its energy lies outside a 16^3 mode box, and the single-pass TF32 analysis -- whose
    # rounding error scales with the energy of the whole field -- leaves 1.3e-2 on the last layer's coefficient gradients
    # (measured); that is the arithmetic of the tf32 route, not a property of this configuration.
    x, y = _band_limited(64, 128, 16.0, 7), _band_limited(64, 128, 16.0, 1007)
    # oracle: the large-scale operator, full width
    modes, n_layers, _, w_large = CFG4_SCALES["large"]
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in ops["large"].state_dict().items()}
    out_o = so.rfno3d_forward(x, sd, modes, n_layers)
    (w_large * F.mse_loss(out_o, y)).backward()
    # library: all three, one weighted loss
    xd, yd = x.to(DEV), y.to(DEV)
    _cabi.profile(True, True)
    total = 0.0
    outs = {}
    for s, op in ops.items():
        op.to(DEV).set_precision("tf32")
        outs[s] = op(xd)
        total = total + CFG4_SCALES[s][3] * F.mse_loss(outs[s], yd)
    total.backward()
    torch.cuda.synchronize()
    rep = _cabi.profile_report(timed=False)
    _cabi.profile(False, True)
    assert math.isfinite(float(total))
    n_blocks = sum(v[1] for v in CFG4_SCALES.values())
    assert rep.get("block_spectral_fwd_kernel", (0,))[0] + rep.get("block_tail_fwd_kernel", (0,))[0] == n_blocks, rep
    assert rep.get("block_tail_bwd_kernel", (0,))[0] + rep.get("block_spectral_bwd_kernel", (0,))[0] >= n_blocks, rep
    assert rep.get("pointwise_linear_kernel", (0,))[0] >= 2 * 3, rep         # wide lift and projection of every operator
    assert not any(k in rep for k in ("analysis_zy_kernel", "synthesis_zy_kernel")), rep        # no FFMA transform
    errs = {"out": rel(outs["large"], out_o)}
    for n, p in ops["large"].named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
        errs[n] = rel(p.grad, sd[n].grad)
    print("\ncfg 4, large-scale operator at width 64 vs CPU oracle (rel L2):", {k: f"{v:.2e}" for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v < 2e-3}
    assert not bad, bad
    for s in ("medium", "small"):
        for n, p in ops[s].named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), (s, n)


@pytest.mark.parametrize("precision,out_tol,grad_tol", [("fp32", 1e-5, 5e-5), ("tf32", 2e-3, 4e-3)])
def test_cfg5_operator_on_256_cubed(precision, out_tol, grad_tol):
    """configs[4]: RationalFourierOperator3D modes (64,64,64) on a 256^3 grid, batch 1 -- against the oracle at width 4 /
    one layer (the CPU side stays in seconds), on the FFMA route and on the per-axis tcgen05 route (transforms_gen.cu)
    that this shape takes at TF32; the multi-GPU slab decomposition of the same shape is covered in test_parity_gpu.py.
    Conditioned weights as in cfg 4.  TF32 gradient gate: 2e-3 per transform, two transforms in series.  The gradients of
    the pointwise layers (torch's own conv / linear backward on both sides) are fp32 sums of 1.7e7 sign-alternating terms
    in different orders on CPU and GPU (6e-4 measured on the conv bias): they get 2e-3, the spectral coefficients the gate."""
    torch.manual_seed(1234)
    w, modes = 4, (64, 64, 64)
    model = pb.RationalFourierOperator3D(modes=modes, width=w, n_layers=1, in_channels=w, out_channels=w)
    with torch.no_grad():
        layer = model.rational_layers[0]
        q0 = layer.Q_coeffs[..., 0, 0, 0].clone()
        layer.Q_coeffs.mul_(1e-6)
        layer.Q_coeffs[..., 0, 0, 0] = q0
    g = torch.Generator().manual_seed(42)
    x = torch.randn(1, w, 256, 256, 256, generator=g)
    y = torch.randn(1, w, 256, 256, 256, generator=g)
    compare_model(model, lambda xx, sd: so.rfno3d_forward(xx, sd, modes, 1), x, y, out_tol=out_tol, grad_tol=max(grad_tol, 2e-3),
                  precision=precision, spectral_grad_tol=grad_tol)
    if precision == "tf32":
        assert _cabi.load().pdephi_plan_supports(_cabi.get_plan(torch.cuda.current_device(), (256, 256, 256), (64, 64, 33)),
                                                 _cabi.PREC_TF32) == 1


def test_cfg5_full_size_adjointness_on_the_tensor_core_route():
    """configs[4] at its full layer size [1, 64, 256^3], modes 64^3, on the per-axis tcgen05 route: <A x, Z> == <x, A^H Z>
    (size-independent property; the transform pair is what the layer's forward and backward are built from), and the
    normalised pair inverts itself on Hermitian-consistent modes."""
    torch.manual_seed(2)
    grid, mh = (256, 256, 256), (64, 64, 33)
    x = torch.randn(1, 64, *grid, device=DEV)
    Z = torch.randn(1, 64, *mh, device=DEV, dtype=torch.complex64)
    Ax = pf.analysis(x, mh, precision=_cabi.PREC_TF32)
    AhZ = pf.synthesis(Z, grid, direction=_cabi.ADJOINT, precision=_cabi.PREC_TF32)
    lhs = (Ax.conj().to(torch.complex128) * Z.to(torch.complex128)).sum().real
    rhs = (x.double() * AhZ.double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-3 * abs(float(lhs))
    del x, AhZ, Ax
    Zc = Z.clone()
    Zc[..., 0] = 0                     # kz = 0 plane is not Hermitian in general: irfftn drops its imaginary part
    back = pf.analysis(pf.synthesis(Zc, grid, precision=_cabi.PREC_TF32), mh, precision=_cabi.PREC_TF32)
    assert rel(back[..., 1:], Zc[..., 1:]) < 2e-3


def test_cfg4_operator_with_wide_input_and_output():
    """MultiScaleFNO's inner operators have in_channels = out_channels = width (multiscale_fno.py:100-127): the lift and
    the projection are then wide channel GEMMs, run channels-first without the reference's permuted copies."""
    torch.manual_seed(1234)
    w, modes = 16, (8, 8, 8)
    model = pb.RationalFourierOperator3D(modes=modes, width=w, n_layers=1, in_channels=w, out_channels=w, rational_order=(2, 2))
    with torch.no_grad():
        layer = model.rational_layers[0]
        q0 = layer.Q_coeffs[..., 0, 0, 0].clone()
        layer.Q_coeffs.mul_(1e-3)
        layer.Q_coeffs[..., 0, 0, 0] = q0
    g = torch.Generator().manual_seed(42)
    x = torch.randn(2, w, 16, 16, 16, generator=g)
    y = torch.randn(2, w, 16, 16, 16, generator=g)
    compare_model(model, lambda xx, sd: so.rfno3d_forward(xx, sd, modes, 1), x, y)
