"""GPU helper for ncu: a few fwd+bwd passes of one whole RFNO block at the cfg-2 shape (TF32 route)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_fluid_phi_b200 as pb
dev = torch.device("cuda")
torch.manual_seed(1234)
model = pb.RationalFourierOperator3D(modes=(32, 32, 32), width=64, n_layers=1).to(dev)
model.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32")
x = torch.randn(8, 3, 128, 128, 128, device=dev)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 4):
    model.zero_grad(set_to_none=True)
    model(x).square().mean().backward()
torch.cuda.synchronize()
print("ok")
