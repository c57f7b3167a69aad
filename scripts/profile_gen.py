"""GPU helper for ncu: a few fwd+bwd passes of a cfg-5-shaped RationalFourierLayer on the per-axis tcgen05 route."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pde_fluid_phi_b200 as pb

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
layer = pb.RationalFourierLayer(64, 64, (64, 64, 64)).cuda()
with torch.no_grad():
    q0 = layer.Q_coeffs[..., 0, 0, 0].clone()
    layer.Q_coeffs.mul_(1e-6)
    layer.Q_coeffs[..., 0, 0, 0] = q0
layer.precision = "tf32"
x = torch.randn(1, 64, 256, 256, 256, device="cuda", requires_grad=True)
g = torch.randn_like(x)
for _ in range(iters):
    x.grad = None
    layer.zero_grad(set_to_none=True)
    layer(x).backward(g)
torch.cuda.synchronize()
print("ok")
