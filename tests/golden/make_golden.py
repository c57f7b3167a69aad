#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules on CPU.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py

For each case the fixture holds the seeded input, the module parameters, the reference's
fp32 output and autograd gradients for a fixed upstream gradient, and the same quantities
from a ``.double()`` copy of the module (the reference's own fp32-vs-fp64 noise floor,
SURVEY.md section 0 fact 4).  The only non-reference code on the path is the two stub modules for
the missing third-party imports (GPUtil, h5py) and, for SpectralConv3D with modes[2] > 2,
the oracle-convention-2 subclass that slices the weight in the einsum --
the unmodified forward raises there.
"""
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PDEPHI_REF", "/root/reference/src")


def load_reference():
    class _Stub(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return type(k, (), {})
    for name in ("GPUtil", "h5py"):
        m = _Stub(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules.setdefault(name, m)
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    from pde_fluid_phi.operators.rational_fourier import RationalFourierLayer, RationalFourierOperator3D
    from pde_fluid_phi.operators.spectral_layers import SpectralConv3D
    from pde_fluid_phi.models.fno3d import FNO3D
    return RationalFourierLayer, RationalFourierOperator3D, SpectralConv3D, FNO3D


RationalFourierLayer, RationalFourierOperator3D, SpectralConv3D, FNO3D = load_reference()


class SlicedSpectralConv3D(SpectralConv3D):
    """Oracle convention 2: reference forward with weights[..., :mz//2+1] in the einsum."""

    def forward(self, x):
        mx, my, mz = self.modes
        x_ft = torch.fft.rfftn(x, dim=[-3, -2, -1])
        x_modes = x_ft[:, :, :mx, :my, :mz // 2 + 1]
        out_modes = torch.einsum('bixyz,oixyz->boxyz', x_modes, self.weights[..., :mz // 2 + 1])
        out_ft = torch.zeros(x.shape[0], self.out_channels, *x_ft.shape[-3:], dtype=x_ft.dtype, device=x.device)
        out_ft[:, :, :mx, :my, :mz // 2 + 1] = out_modes
        return torch.fft.irfftn(out_ft, s=x.shape[-3:], dim=[-3, -2, -1])


def run(module, x, g):
    """forward + backward; returns out, dx, {param: grad}."""
    module.zero_grad(set_to_none=True)
    x = x.clone().requires_grad_(True)
    out = module(x)
    out.backward(g)
    grads = {n: p.grad.detach().clone() for n, p in module.named_parameters()}
    return out.detach(), x.grad.detach(), grads


def to_np(t):
    t = t.detach()
    return t.numpy().copy()


def save_case(name, module, x, extra=None, fp64=True):
    torch.manual_seed(4321)
    with torch.no_grad():
        g = torch.randn_like(module(x))
    out, dx, grads = run(module, x, g)
    import copy
    blob = {"x": to_np(x), "g": to_np(g), "out": to_np(out), "dx": to_np(dx)}
    grads64, nf = {}, float("nan")
    if fp64:
        m64 = copy.deepcopy(module).double()
        for p in m64.parameters():                # nn.Module.double() leaves complex64 parameters alone
            if p.is_complex():
                p.data = p.data.to(torch.complex128)
        out64, dx64, grads64 = run(m64, x.double(), g.double())
        blob.update({"out64": to_np(out64), "dx64": to_np(dx64)})
        nf = float((out.double() - out64).norm() / out64.norm())
    for n, p in module.state_dict().items():
        blob["sd:" + n] = to_np(p)
    for n, v in grads.items():
        blob["grad:" + n] = to_np(v)
    for n, v in grads64.items():
        blob["grad64:" + n] = to_np(v)
    for k, v in (extra or {}).items():
        blob["meta:" + k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **blob)
    print(f"{name:32s} out{tuple(out.shape)}  ref fp32-vs-fp64 rel L2 = {nf:.2e}  ({os.path.getsize(path)/1024:.0f} KB)")


def condition(layer, factor=1e-4):
    """Weight set (ii) of SURVEY.md section 8c: shrink all non-constant Q coefficients so Q stays in [0.9, 3]."""
    with torch.no_grad():
        q0 = layer.Q_coeffs[..., 0, 0, 0].clone()
        layer.Q_coeffs.mul_(factor)
        layer.Q_coeffs[..., 0, 0, 0] = q0


def main():
    # --- RationalFourierLayer ---------------------------------------------------------------
    torch.manual_seed(1234)
    layer = RationalFourierLayer(4, 4, (4, 4, 4))
    torch.manual_seed(42)
    save_case("rfl_c4_m4_n8_default", layer, torch.randn(2, 4, 8, 8, 8),
              {"modes": (4, 4, 4), "rational_order": (4, 4), "eps": 1e-6})

    torch.manual_seed(1234)
    layer = RationalFourierLayer(4, 4, (4, 4, 4))
    condition(layer)
    torch.manual_seed(42)
    save_case("rfl_c4_m4_n8_conditioned", layer, torch.randn(2, 4, 8, 8, 8),
              {"modes": (4, 4, 4), "rational_order": (4, 4), "eps": 1e-6})

    torch.manual_seed(1234)
    layer = RationalFourierLayer(3, 3, (6, 4, 8), rational_order=(4, 3))
    torch.manual_seed(42)
    save_case("rfl_c3_m648_n161220_order43", layer, torch.randn(2, 3, 16, 12, 20),
              {"modes": (6, 4, 8), "rational_order": (4, 3), "eps": 1e-6})

    torch.manual_seed(1234)
    layer = RationalFourierLayer(2, 2, (8, 8, 8), rational_order=(2, 2))
    torch.manual_seed(42)
    save_case("rfl_c2_m8_n8_nyquist_order22", layer, torch.randn(3, 2, 8, 8, 8),
              {"modes": (8, 8, 8), "rational_order": (2, 2), "eps": 1e-6})

    torch.manual_seed(1234)
    layer = RationalFourierLayer(8, 8, (8, 8, 8))
    torch.manual_seed(42)
    save_case("rfl_c8_m8_n16_default", layer, torch.randn(1, 8, 16, 16, 16),
              {"modes": (8, 8, 8), "rational_order": (4, 4), "eps": 1e-6})

    torch.manual_seed(1234)
    layer = RationalFourierLayer(2, 2, (5, 3, 6), rational_order=(3, 3))
    torch.manual_seed(42)
    save_case("rfl_c2_m536_n9711_odd_order33", layer, torch.randn(1, 2, 9, 7, 11),
              {"modes": (5, 3, 6), "rational_order": (3, 3), "eps": 1e-6})

    # --- SpectralConv3D ---------------------------------------------------------------------
    torch.manual_seed(1234)
    conv = SlicedSpectralConv3D(4, 4, (4, 4, 4))
    torch.manual_seed(42)
    save_case("sc3d_c4_m4_n8_sliced", conv, torch.randn(2, 4, 8, 8, 8), {"modes": (4, 4, 4)})

    torch.manual_seed(1234)
    conv = SpectralConv3D(3, 3, (8, 8, 1))                      # unmodified reference (2-D config)
    torch.manual_seed(42)
    save_case("sc3d_c3_m881_n16x16x1_unmodified", conv, torch.randn(2, 3, 16, 16, 1), {"modes": (8, 8, 1)},
              fp64=False)   # the unmodified forward hard-codes cfloat (spectral_layers.py:69-72)

    torch.manual_seed(1234)
    conv = SpectralConv3D(2, 2, (4, 6, 2))                      # unmodified reference, mz == 2
    torch.manual_seed(42)
    save_case("sc3d_c2_m462_n8x12x6_unmodified", conv, torch.randn(2, 2, 8, 12, 6), {"modes": (4, 6, 2)},
              fp64=False)

    # --- whole operators --------------------------------------------------------------------
    torch.manual_seed(1234)
    op = RationalFourierOperator3D(modes=(4, 4, 4), width=8, n_layers=2, rational_order=(3, 3))
    torch.manual_seed(42)
    save_case("rfno3d_w8_m4_l2_n8", op, torch.randn(2, 3, 8, 8, 8),
              {"modes": (4, 4, 4), "rational_order": (3, 3), "n_layers": 2, "width": 8})

    torch.manual_seed(1234)
    fno = FNO3D(modes=(4, 4, 4), width=8, n_layers=2)
    for i in range(2):                                           # convention 2 for the inner layers
        fno.spectral_layers[i].__class__ = SlicedSpectralConv3D
    torch.manual_seed(42)
    save_case("fno3d_w8_m4_l2_n8_sliced", fno, torch.randn(2, 3, 8, 8, 8),
              {"modes": (4, 4, 4), "n_layers": 2, "width": 8})


def main_next_rows():
    """Fixtures for the 'next' rows of SURVEY.md section 8f: the rollout loop (N3) and RationalFNO's coarse branch (N4)."""
    from pde_fluid_phi.models.rfno import RationalFNO

    # --- N4: RationalFNO with the spectral down / up-sampling branch ---------------------------
    torch.manual_seed(1234)
    model = RationalFNO(modes=(4, 4, 4), width=8, n_layers=2, rational_order=(3, 3), multi_scale=True)
    torch.manual_seed(42)
    save_case("rfno_model_w8_m4_l2_n16_multiscale", model, torch.randn(2, 3, 16, 16, 16),
              {"modes": (4, 4, 4), "rational_order": (3, 3), "n_layers": 2, "width": 8})

    # --- N3: rollout (RK4 + CFL timestep + constraints + 2/3 de-aliasing) ----------------------
    torch.manual_seed(1234)
    op = RationalFourierOperator3D(modes=(4, 4, 4), width=8, n_layers=2, rational_order=(3, 3))
    torch.manual_seed(42)
    x = torch.randn(2, 3, 12, 12, 12)
    blob = {"x": to_np(x)}
    for n, p in op.state_dict().items():
        blob["sd:" + n] = to_np(p)
    with torch.no_grad():
        blob["traj_adaptive"] = to_np(op.rollout(x, steps=3, return_trajectory=True, adaptive_timestep=True))
        blob["traj_fixed"] = to_np(op.rollout(x, steps=3, return_trajectory=True, adaptive_timestep=False))
        blob["final_adaptive_cfl02"] = to_np(op.rollout(x, steps=4, adaptive_timestep=True, max_cfl=0.2))
        blob["dealias"] = to_np(op._apply_dealiasing(x))
    for k, v in {"modes": (4, 4, 4), "rational_order": (3, 3), "n_layers": 2, "width": 8}.items():
        blob["meta:" + k] = np.asarray(v)
    path = os.path.join(HERE, "rollout_rfno3d_w8_m4_l2_n12.npz")
    np.savez_compressed(path, **blob)
    print(f"rollout_rfno3d_w8_m4_l2_n12: trajectory {blob['traj_adaptive'].shape} ({os.path.getsize(path)/1024:.0f} KB)")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "next":
        main_next_rows()
    else:
        main()
