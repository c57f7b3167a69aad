"""Drop-in mirrors of the two operator models that own the hot loop.

  RationalFourierOperator3D  <- operators/rational_fourier.py:175-421
  FNO3D                      <- models/fno3d.py:17-142

Constructor signatures, sub-module / parameter names and the order in which the global RNG is
consumed match the reference, so ``torch.manual_seed(s); Model(...)`` gives the same weights and
checkpoints load either way.  The whole block -- spectral layer, 1x1x1 convolution, both residual adds,
activation -- runs in libpdephi_b200.so: at width 64 with GELU on the tcgen05 kernels of block_tc.cu, with
the input lift folded into the first block and the output projection into the last (SURVEY.md section 8f
rows N1, N2); at any other width <= 64, spatial size or usual activation on the fp32 kernels of
pointwise_wide.cu, which also carry the wide -> wide lift / projection of operators built with
in_channels = out_channels = width.  Only a width above 64 or an exotic activation keeps torch's Conv3d and
activation around the library's spectral layer, whose synthesis epilogue then carries both skip adds.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi
from . import compile_ops
from .functional import (PointwiseLinearFn, WideBlockTailFn, head_supported, library_call, pointwise_linear_supported,
                         spectral_resample, wide_activation_code, wide_block_tail_supported)
from .operators import RationalFourierLayer, SpectralConv3D


class StabilityConstraints:
    """Physical-space realizability clamp and energy cap used by ``rollout`` (operators/stability.py:106-197).

    Same public surface (``apply``, ``get_metrics``, the four metric keys); metrics are kept as device
    scalars and only converted when ``get_metrics`` is called, so ``apply`` does not sync the host.
    """

    def __init__(self, method: str = "rational_decay", decay_rate: float = 2.0, passivity_constraint: bool = True,
                 realizability: bool = True, max_spectral_radius: float = 0.99):
        self.method = method
        self.decay_rate = decay_rate
        self.passivity_constraint = passivity_constraint
        self.realizability = realizability
        self.max_spectral_radius = max_spectral_radius
        self.metrics: Dict[str, object] = {"spectral_radius": 0.0, "energy_drift": 0.0,
                                           "realizability_violations": 0, "passivity_violations": 0}

    def apply(self, x: torch.Tensor) -> torch.Tensor:
        y = x
        if self.realizability:
            self.metrics["realizability_violations"] = (x.abs() > 100.0).sum()
            y = y.clamp(min=-100.0, max=100.0)
        if self.passivity_constraint:
            energy = (y * y).sum(dim=(-3, -2, -1), keepdim=True)
            self.metrics["passivity_violations"] = (energy > 1000.0).sum()
            y = y * torch.sqrt(1000.0 / (energy + 1e-8)).clamp(max=1.0)
        e0, e1 = (x * x).sum(), (y * y).sum()
        self.metrics["energy_drift"] = (e1 - e0).abs() / (e0 + 1e-8)
        self.metrics["spectral_radius"] = torch.linalg.vector_norm(y - x) / (torch.linalg.vector_norm(x) + 1e-8)
        return y

    def get_metrics(self) -> Dict[str, float]:
        return {k: (float(v) if isinstance(v, torch.Tensor) else v) for k, v in self.metrics.items()}


def _lift(x: torch.Tensor, linear: nn.Linear) -> torch.Tensor:
    """The reference's rearrange -> Linear -> rearrange (rational_fourier.py:258-260,274-276) as one
    channels-first contraction on the [B,C,X,Y,Z] layout: same result, no permuted copies.  Runs in
    libpdephi_b200.so (pointwise.cu); shapes it does not cover use the reference's own torch formulation."""
    if x.dim() != 5:
        raise ValueError(f"expected a [batch, channels, Nx, Ny, Nz] field, got shape {tuple(x.shape)}")
    if x.shape[1] != linear.in_features:
        raise RuntimeError(f"input has {x.shape[1]} channels, the model was built for {linear.in_features}")
    if compile_ops.compiling():      # torch.compile: the reference's own formulation, for the compiler to lower
        return linear(x.permute(0, 2, 3, 4, 1)).permute(0, 4, 1, 2, 3)
    if pointwise_linear_supported(x, linear.weight):
        return library_call(PointwiseLinearFn, x, linear.weight, linear.bias)
    if x.is_cuda and x.dtype == torch.float32:
        # more than 64 channels on both sides: a plain library GEMM on the channels-first layout (cuDNN 1x1x1
        # convolution, fp32 like the reference's Linear) instead of two permuted copies around one
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            return F.conv3d(x, linear.weight.view(linear.out_features, linear.in_features, 1, 1, 1), linear.bias)
    return linear(x.permute(0, 2, 3, 4, 1)).permute(0, 4, 1, 2, 3)


def _generic_block(activation, spectral: nn.Module, local: nn.Module, h: torch.Tensor) -> torch.Tensor:
    """activation(spectral(h) + local(h) + h) where the tcgen05 block kernels do not apply: the block tail as one fp32 library
    kernel (any width <= 64, the usual activations: pointwise_wide.cu); beyond that torch's own convolution and activation
    around the spectral layer, with both adds in the synthesis kernel's epilogue."""
    if compile_ops.compiling():      # torch.compile: spectral layer = one custom op (adds in its epilogue), the rest is traced
        return activation(spectral.forward_fused(h, local(h)))
    code = wide_activation_code(activation)
    # (also under a slab decomposition: the tail is pointwise, its weight gradients are slab-local sums like every other
    # pointwise layer's, and the single-GPU and the slab run of one model then share the same fp32 arithmetic)
    if code is not None and wide_block_tail_supported(h, local):
        return library_call(WideBlockTailFn, spectral(h), h, local.weight, local.bias, code)
    return activation(spectral.forward_fused(h, local(h)))


def _reset_pointwise(module: nn.Module):
    # same per-module init calls, in module registration order, as the reference's _init_weights
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Conv3d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            if m.bias is not None:
                nn.init.zeros_(m.bias)


class RationalFourierOperator3D(nn.Module):
    """lift -> n x [RationalFourierLayer + 1x1x1 conv + residual + activation] -> project."""

    def __init__(self, modes: Tuple[int, int, int] = (32, 32, 32), width: int = 64, n_layers: int = 4,
                 in_channels: int = 3, out_channels: int = 3, rational_order: Tuple[int, int] = (4, 4),
                 activation: str = "gelu", final_activation: Optional[str] = None,
                 stability_constraints: Optional[object] = None):
        super().__init__()
        self.modes = modes
        self.width = width
        self.n_layers = n_layers
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.input_proj = nn.Linear(in_channels, width)
        self.output_proj = nn.Linear(width, out_channels)
        self.rational_layers = nn.ModuleList([
            RationalFourierLayer(in_channels=width, out_channels=width, modes=modes, rational_order=rational_order)
            for _ in range(n_layers)])
        self.local_convs = nn.ModuleList([nn.Conv3d(width, width, kernel_size=1, bias=True) for _ in range(n_layers)])
        self.activation = getattr(F, activation)
        self.final_activation = getattr(F, final_activation) if final_activation else None
        self.stability_constraints = stability_constraints or StabilityConstraints()
        _reset_pointwise(self)
        self.fuse_head = True      # last block + output projection in one kernel chain where the shapes allow (not a parameter)
        self.fuse_lift = True      # first block + input lift: mode-space analysis / lift backward (functional._lift_forward)

    def set_precision(self, precision: Optional[str]):
        """'auto' | 'fp32' | 'tf32x3' | 'tf32' | None (module default) for every spectral layer of this model."""
        for layer in self.rational_layers:
            layer.precision = precision
        return self

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        gelu = self.activation is F.gelu
        last = len(self.rational_layers) - 1
        out = vz = None
        first = 0
        if (gelu and self.fuse_lift and last > 0 and x.dim() == 5 and x.shape[1] == self.input_proj.in_features
                and self.rational_layers[0].lifted_block_supported(x, self.local_convs[0], self.input_proj)):
            # the lift rides in the first block: its analysis and (x needs no gradient) its backward run in mode space
            h, vz = self.rational_layers[0].forward_block(x, self.local_convs[0], None, self.input_proj, want_vz=True)
            first = 1
        else:
            h = _lift(x, self.input_proj)
        for i, (spectral, local) in enumerate(zip(self.rational_layers, self.local_convs)):
            if i < first:
                continue
            if gelu and spectral.block_supported(h, local):
                # vz: the z stage of this block's forward analysis, left by the previous block's fused kernel (None off that route)
                if i == last and self.fuse_head and head_supported(self.output_proj):
                    out = spectral.forward_block(h, local, self.output_proj, vz_in=vz)   # ... and the output projection with it
                else:
                    res = spectral.forward_block(h, local, vz_in=vz, want_vz=i < last)  # whole block in the library (width 64)
                    h, vz = res if i < last else (res, None)
            else:
                h, vz = _generic_block(self.activation, spectral, local, h), None
        if out is None:
            out = _lift(h, self.output_proj)
        return self.final_activation(out) if self.final_activation is not None else out

    # ---- inference-time rollout (rational_fourier.py:284-412) ---------------------
    # Same results as the reference loop, without its host round trips: the CFL timestep, the running time and the
    # "target time reached" test stay on the device (the reference reads max|u| back with .item() once per step and
    # the constraints read four metrics back per call), and the 2/3 de-aliasing is one pruned K1 + K3 pair instead of
    # a full rfftn / irfftn.  Once the target time is reached the remaining iterations leave the state untouched
    # (torch.where), which is what the reference's `break` does.
    def rollout(self, initial_condition: torch.Tensor, steps: int, return_trajectory: bool = False,
                adaptive_timestep: bool = True, max_cfl: float = 0.5) -> torch.Tensor:
        state = initial_condition
        trajectory = [state.clone()] if return_trajectory else None
        if not adaptive_timestep:
            for step in range(steps):
                with (torch.no_grad() if step > 0 else torch.enable_grad()):
                    state = self._apply_dealiasing(self.stability_constraints.apply(self.forward(state)))
                    if return_trajectory:
                        trajectory.append(state.clone())
            return torch.stack(trajectory, dim=1) if return_trajectory else state

        dt = self._adaptive_timestep(state, max_cfl)                 # float64 device scalars (the reference's Python floats)
        target_time = steps * dt
        total_time = torch.zeros_like(dt)
        actives = []
        for step in range(steps):
            dt = torch.minimum(dt, self._adaptive_timestep(state, max_cfl))
            active = total_time < target_time                        # the reference breaks out of the loop when this fails
            dt = torch.where(total_time + dt > target_time, target_time - total_time, dt)
            with (torch.no_grad() if step > 0 else torch.enable_grad()):
                new = self._rk4_step(state, dt) if step > 0 else self.forward(state)
                new = self._apply_dealiasing(self.stability_constraints.apply(new))
                state = torch.where(active, new, state)
                if return_trajectory:
                    trajectory.append(state.clone())
            total_time = torch.where(active, total_time + dt, total_time)
            actives.append(active)
        if return_trajectory:
            # the reference's trajectory stops growing at the break: drop the frozen tail (one host read, after the loop)
            taken = int(torch.stack(actives).sum()) if actives else 0
            return torch.stack(trajectory[:1 + taken], dim=1)
        return state

    def _adaptive_timestep(self, u: torch.Tensor, max_cfl: float) -> torch.Tensor:
        """CFL timestep as a float64 device scalar: max_cfl * dx / (max |u| + 1e-8), dx = 1 / Nx."""
        with torch.no_grad():
            u_max = torch.sqrt((u * u).sum(dim=1)).max().double()
            return (max_cfl * (1.0 / u.shape[-3])) / (u_max + 1e-8)

    def _compute_adaptive_timestep(self, u: torch.Tensor, max_cfl: float) -> float:
        """Host-float variant with the reference's signature (rational_fourier.py:356-372); syncs."""
        return float(self._adaptive_timestep(u, max_cfl))

    def _rk4_step(self, u: torch.Tensor, dt) -> torch.Tensor:
        k1 = self.forward(u)
        k2 = self.forward(u + 0.5 * dt * k1)
        k3 = self.forward(u + 0.5 * dt * k2)
        k4 = self.forward(u + dt * k3)
        return u + (dt / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)

    def _apply_dealiasing(self, u: torch.Tensor) -> torch.Tensor:
        """2/3 rule (rational_fourier.py:392-412): the reference zeroes every index >= 2N/3 of the half-spectrum, i.e.
        it keeps the box [0, 2Nx/3) x [0, 2Ny/3) x [0, 2(Nz/2+1)/3) -- exactly a pruned analysis / synthesis pair."""
        nx, ny, nz = u.shape[-3:]
        keep = ((2 * nx) // 3, (2 * ny) // 3, (2 * (nz // 2 + 1)) // 3)
        with torch.no_grad():
            if not u.is_cuda or min(keep) < 1:
                raise RuntimeError("rollout de-aliasing runs on the CUDA path only and needs at least a 2-point grid per axis")
            return spectral_resample(u.float(), keep, (nx, ny, nz))

    def get_stability_monitor(self) -> dict:
        return self.stability_constraints.get_metrics()

    def create_trainer(self, **kwargs):
        """The reference's StabilityTrainer drives this model unchanged (it only calls model(x) / backward)."""
        try:
            from pde_fluid_phi.training.stability_trainer import StabilityTrainer
        except ImportError as e:  # the trainer is reference code, out of scope for this package
            raise ImportError("create_trainer needs the reference package `pde_fluid_phi` on sys.path") from e
        return StabilityTrainer(self, **kwargs)


class FNO3D(nn.Module):
    """lift -> n x [SpectralConv3D + 1x1x1 conv + residual + activation] -> project."""

    def __init__(self, modes: Tuple[int, int, int] = (32, 32, 32), width: int = 64, n_layers: int = 4,
                 in_channels: int = 3, out_channels: int = 3, activation: str = "gelu",
                 final_activation: Optional[str] = None):
        super().__init__()
        self.modes = modes
        self.width = width
        self.n_layers = n_layers
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.input_proj = nn.Linear(in_channels, width)
        self.output_proj = nn.Linear(width, out_channels)
        self.spectral_layers = nn.ModuleList([SpectralConv3D(width, width, modes) for _ in range(n_layers)])
        self.local_layers = nn.ModuleList([nn.Conv3d(width, width, kernel_size=1, bias=True) for _ in range(n_layers)])
        self.activation = getattr(F, activation)
        self.final_activation = getattr(F, final_activation) if final_activation else None
        _reset_pointwise(self)
        self.fuse_head = True      # last block + output projection in one kernel chain where the shapes allow (not a parameter)
        self.fuse_lift = True      # first block + input lift: mode-space analysis / lift backward (functional._lift_forward)

    def set_precision(self, precision: Optional[str]):
        for layer in self.spectral_layers:
            layer.precision = precision
        return self

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        gelu = self.activation is F.gelu
        last = len(self.spectral_layers) - 1
        out = vz = None
        first = 0
        if (gelu and self.fuse_lift and last > 0 and x.dim() == 5 and x.shape[1] == self.input_proj.in_features
                and self.spectral_layers[0].lifted_block_supported(x, self.local_layers[0], self.input_proj)):
            h, vz = self.spectral_layers[0].forward_block(x, self.local_layers[0], None, self.input_proj, want_vz=True)
            first = 1
        else:
            h = _lift(x, self.input_proj)
        for i, (spectral, local) in enumerate(zip(self.spectral_layers, self.local_layers)):
            if i < first:
                continue
            if gelu and spectral.block_supported(h, local):
                if i == last and self.fuse_head and head_supported(self.output_proj):
                    out = spectral.forward_block(h, local, self.output_proj, vz_in=vz)
                else:
                    res = spectral.forward_block(h, local, vz_in=vz, want_vz=i < last)
                    h, vz = res if i < last else (res, None)
            else:
                h, vz = _generic_block(self.activation, spectral, local, h), None
        if out is None:
            out = _lift(h, self.output_proj)
        return self.final_activation(out) if self.final_activation is not None else out

    def rollout(self, initial_condition: torch.Tensor, steps: int, return_trajectory: bool = False) -> torch.Tensor:
        state = initial_condition
        trajectory = [state.clone()] if return_trajectory else None
        for step in range(steps):
            with (torch.no_grad() if step > 0 else torch.enable_grad()):
                state = self.forward(state)
                if return_trajectory:
                    trajectory.append(state.clone())
        return torch.stack(trajectory, dim=1) if return_trajectory else state


class RationalFNO(nn.Module):
    """RationalFourierOperator3D plus a half-resolution branch, blended 0.7 / 0.3 (models/rfno.py:18-117; SURVEY.md
    section 8f row N4 -- the model the reference's CLI trains by default).

    Same constructor, sub-module names (``rational_fno``, ``coarse_processor``) and RNG order as the reference.  The
    spectral down- and up-sampling of the coarse branch (rfno.py:119-149) are pruned K1 / K3 pairs with different
    input and output grids (``functional.spectral_resample``) instead of full rfftn / irfftn round trips.
    """

    def __init__(self, modes: Tuple[int, int, int] = (32, 32, 32), width: int = 64, n_layers: int = 4, in_channels: int = 3,
                 out_channels: int = 3, rational_order: Tuple[int, int] = (4, 4), activation: str = "gelu",
                 final_activation: Optional[str] = None, stability_weight: float = 0.01, spectral_reg_weight: float = 0.001,
                 multi_scale: bool = True, **kwargs):
        super().__init__()
        self.modes = modes
        self.width = width
        self.n_layers = n_layers
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.multi_scale = multi_scale
        self.rational_fno = RationalFourierOperator3D(modes=modes, width=width, n_layers=n_layers, in_channels=in_channels,
                                                      out_channels=out_channels, rational_order=rational_order,
                                                      activation=activation, final_activation=final_activation)
        if multi_scale:
            self.coarse_modes = tuple(m // 2 for m in modes)
            self.fine_modes = modes
            self.coarse_processor = RationalFourierOperator3D(modes=self.coarse_modes, width=width // 2, n_layers=2,
                                                              in_channels=in_channels, out_channels=out_channels,
                                                              rational_order=(2, 2))
        self.stability_constraints = StabilityConstraints(method="rational_decay", decay_rate=2.0, passivity_constraint=True,
                                                          realizability=True)
        self.stability_weight = stability_weight
        self.spectral_reg_weight = spectral_reg_weight
        self.training_metrics: Dict[str, float] = {}

    def set_precision(self, precision: Optional[str]):
        self.rational_fno.set_precision(precision)
        if self.multi_scale:
            self.coarse_processor.set_precision(precision)
        return self

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out = self.rational_fno(x)
        if self.multi_scale:
            coarse = self.coarse_processor(self._downsample(x, factor=2))
            out = 0.7 * out + 0.3 * self._upsample(coarse, target_shape=tuple(x.shape[-3:]))
        return self.stability_constraints.apply(out)

    def _downsample(self, x: torch.Tensor, factor: int) -> torch.Tensor:
        """Keep the [0, N/f) corner of the half-spectrum and invert it on the N/f grid (rfno.py:119-130)."""
        nx, ny, nz = (n // factor for n in x.shape[-3:])
        return spectral_resample(x, (nx, ny, nz // 2 + 1), (nx, ny, nz))

    def _upsample(self, x: torch.Tensor, target_shape: Tuple[int, int, int]) -> torch.Tensor:
        """Zero-pad the whole half-spectrum of x into the target grid's and invert there (rfno.py:132-149)."""
        nx, ny, nz = x.shape[-3:]
        return spectral_resample(x, (nx, ny, nz // 2 + 1), tuple(target_shape))

    # ---- loss terms of the reference trainer (rfno.py:151-232): callers of the path, plain tensor ops ----
    def compute_losses(self, predicted: torch.Tensor, target: torch.Tensor, input_field: torch.Tensor) -> Dict[str, torch.Tensor]:
        losses = {"data": F.mse_loss(predicted, target), "divergence": self._divergence_loss(predicted),
                  "energy_conservation": self._energy_conservation_loss(input_field, predicted),
                  "spectral_reg": self._spectral_regularization(predicted)}
        radius = self.stability_constraints.get_metrics()["spectral_radius"]
        losses["stability"] = torch.tensor(radius, device=predicted.device, requires_grad=True)
        losses["total"] = (losses["data"] + 0.1 * losses["divergence"] + 0.05 * losses["energy_conservation"]
                           + self.spectral_reg_weight * losses["spectral_reg"] + self.stability_weight * losses["stability"])
        return losses

    @staticmethod
    def _divergence_loss(u: torch.Tensor) -> torch.Tensor:
        div = torch.gradient(u[:, 0], dim=-3)[0] + torch.gradient(u[:, 1], dim=-2)[0] + torch.gradient(u[:, 2], dim=-1)[0]
        return div.pow(2).mean()

    @staticmethod
    def _energy_conservation_loss(u_initial: torch.Tensor, u_final: torch.Tensor) -> torch.Tensor:
        e0 = 0.5 * u_initial.pow(2).sum(dim=(-3, -2, -1))
        e1 = 0.5 * u_final.pow(2).sum(dim=(-3, -2, -1))
        return ((e1 - e0).abs() / (e0 + 1e-8)).mean()

    @staticmethod
    def _spectral_regularization(u: torch.Tensor) -> torch.Tensor:
        u_ft = torch.fft.rfftn(u, dim=[-3, -2, -1])
        density = u_ft.abs().pow(2).sum(dim=1)
        nx, ny, nz = u.shape[-3:]
        kx, ky, kz = torch.meshgrid(torch.fft.fftfreq(nx, device=u.device), torch.fft.fftfreq(ny, device=u.device),
                                    torch.fft.rfftfreq(nz, device=u.device), indexing="ij")
        high = (torch.sqrt(kx ** 2 + ky ** 2 + kz ** 2) > min(nx, ny, nz) // 4).to(density.dtype)
        return (density * high).sum() / (density.sum() + 1e-8)

    def rollout(self, initial_condition: torch.Tensor, steps: int, return_trajectory: bool = False,
                stability_check: bool = True) -> torch.Tensor:
        state = initial_condition
        trajectory = [state.clone()] if return_trajectory else None
        for step in range(steps):
            with torch.no_grad():
                state = self.forward(state)
                if stability_check and self.stability_constraints.get_metrics()["spectral_radius"] > 1.0:
                    print(f"Warning: Unstable at step {step}")
                    break
                if return_trajectory:
                    trajectory.append(state.clone())
        return torch.stack(trajectory, dim=1) if return_trajectory else state

    def get_stability_monitor(self) -> Dict[str, float]:
        return self.stability_constraints.get_metrics()

    def create_trainer(self, **kwargs):
        return RationalFourierOperator3D.create_trainer(self, **kwargs)
