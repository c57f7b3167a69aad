"""Drop-in mirrors of the two operator models that own the hot loop.

  RationalFourierOperator3D  <- operators/rational_fourier.py:175-421
  FNO3D                      <- models/fno3d.py:17-142

Constructor signatures, sub-module / parameter names and the order in which the global RNG is
consumed match the reference, so ``torch.manual_seed(s); Model(...)`` gives the same weights and
checkpoints load either way.  Per block the spectral layer, the two skip adds and (inside the
layer) the projection run in libpdephi_b200.so, as do the lift / project Linears (SURVEY.md section 8f
row N2); the 1x1x1 convolution and GELU stay with torch (row N1).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .functional import PointwiseLinearFn, pointwise_linear_supported
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
    if pointwise_linear_supported(x, linear.weight):
        return PointwiseLinearFn.apply(x, linear.weight, linear.bias)
    return linear(x.permute(0, 2, 3, 4, 1)).permute(0, 4, 1, 2, 3)


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

    def set_precision(self, precision: Optional[str]):
        """'auto' | 'fp32' | 'tf32x3' | 'tf32' | None (module default) for every spectral layer of this model."""
        for layer in self.rational_layers:
            layer.precision = precision
        return self

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = _lift(x, self.input_proj)
        gelu = self.activation is F.gelu
        for spectral, local in zip(self.rational_layers, self.local_convs):
            if gelu and spectral.block_supported(h, local):
                h = spectral.forward_block(h, local)          # whole block in the library (TF32 path, width 64)
            else:   # spectral(h) + local(h) + h in one pass: both adds ride in the synthesis kernel's epilogue
                h = self.activation(spectral.forward_fused(h, local(h)))
        out = _lift(h, self.output_proj)
        return self.final_activation(out) if self.final_activation is not None else out

    # ---- inference-time rollout (rational_fourier.py:284-412); same hot path, forward only -------
    def rollout(self, initial_condition: torch.Tensor, steps: int, return_trajectory: bool = False,
                adaptive_timestep: bool = True, max_cfl: float = 0.5) -> torch.Tensor:
        state = initial_condition
        trajectory = [state.clone()] if return_trajectory else None
        dt = total_time = target_time = 0.0
        if adaptive_timestep:
            dt = self._compute_adaptive_timestep(state, max_cfl)
            target_time = steps * dt
        for step in range(steps):
            if adaptive_timestep:
                dt = min(dt, self._compute_adaptive_timestep(state, max_cfl))
                if total_time >= target_time:
                    break
                if total_time + dt > target_time:
                    dt = target_time - total_time
            with (torch.no_grad() if step > 0 else torch.enable_grad()):
                state = self._rk4_step(state, dt) if (adaptive_timestep and step > 0) else self.forward(state)
                state = self.stability_constraints.apply(state)
                state = self._apply_dealiasing(state)
                if return_trajectory:
                    trajectory.append(state.clone())
            if adaptive_timestep:
                total_time += dt
        return torch.stack(trajectory, dim=1) if return_trajectory else state

    def _compute_adaptive_timestep(self, u: torch.Tensor, max_cfl: float) -> float:
        with torch.no_grad():
            u_max = float(torch.sqrt((u * u).sum(dim=1)).max())
        return float(max_cfl * (1.0 / u.shape[-3]) / (u_max + 1e-8))

    def _rk4_step(self, u: torch.Tensor, dt: float) -> torch.Tensor:
        k1 = self.forward(u)
        k2 = self.forward(u + 0.5 * dt * k1)
        k3 = self.forward(u + 0.5 * dt * k2)
        k4 = self.forward(u + dt * k3)
        return u + (dt / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)

    def _apply_dealiasing(self, u: torch.Tensor) -> torch.Tensor:
        """2/3 rule on the full spectrum (not part of the pruned hot path; rational_fourier.py:392-412)."""
        with torch.no_grad():
            u_ft = torch.fft.rfftn(u, dim=[-3, -2, -1])
            nx, ny, nzh = u_ft.shape[-3:]
            u_ft[:, :, (2 * nx) // 3:, :, :] = 0
            u_ft[:, :, :, (2 * ny) // 3:, :] = 0
            u_ft[:, :, :, :, (2 * nzh) // 3:] = 0
            return torch.fft.irfftn(u_ft, s=u.shape[-3:], dim=[-3, -2, -1])

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

    def set_precision(self, precision: Optional[str]):
        for layer in self.spectral_layers:
            layer.precision = precision
        return self

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = _lift(x, self.input_proj)
        gelu = self.activation is F.gelu
        for spectral, local in zip(self.spectral_layers, self.local_layers):
            if gelu and spectral.block_supported(h, local):
                h = spectral.forward_block(h, local)
            else:
                h = self.activation(spectral.forward_fused(h, local(h)))
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
