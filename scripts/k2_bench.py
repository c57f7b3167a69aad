"""GPU helper: the K2 kernels alone at the cfg-2 shape (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_fluid_phi_b200 as pb
from pde_fluid_phi_b200 import functional as pf, _cabi
torch.manual_seed(0)
dev = "cuda"
B, C, mh = 8, 64, (32, 32, 17)
layer = pb.RationalFourierLayer(C, C, (32, 32, 32)).to(dev)
X = torch.randn(B, C, *mh, device=dev, dtype=torch.complex64)
dZ = torch.randn(B, C, *mh, device=dev, dtype=torch.complex64)
mask = layer.stability_projection.decay_mask
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    Y, stats, R, Qe = pf.rational_mix_fwd(X, layer.P_coeffs.detach(), layer.Q_coeffs.detach(), layer._mono_p, layer._mono_q, 1e-6, mask, 1e-6, True)
    dY = pf.projection_bwd(dZ, Y, mask, stats)
    dX, dP, dQ = pf.rational_mix_bwd(X, dY, R, Qe, layer._mono_p, layer._mono_q, (tuple(layer.P_coeffs.shape), tuple(layer.Q_coeffs.shape)))
torch.cuda.synchronize()
_cabi.profile(enable=True, reset=True)
for _ in range(10):
    Y, stats, R, Qe = pf.rational_mix_fwd(X, layer.P_coeffs.detach(), layer.Q_coeffs.detach(), layer._mono_p, layer._mono_q, 1e-6, mask, 1e-6, True)
    dY = pf.projection_bwd(dZ, Y, mask, stats)
    dX, dP, dQ = pf.rational_mix_bwd(X, dY, R, Qe, layer._mono_p, layer._mono_q, (tuple(layer.P_coeffs.shape), tuple(layer.Q_coeffs.shape)))
torch.cuda.synchronize()
print({k: round(v[1] / v[0] * 1e3, 1) for k, v in _cabi.profile_report().items()})
