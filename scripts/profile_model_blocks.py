"""GPU helper for ncu: fwd+bwd passes of a 3-block RFNO at the cfg-2 shape, so that one pass launches every variant of the
block kernels (first block with the lift, middle block, last block with the projection)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_fluid_phi_b200 as pb
dev = torch.device("cuda")
torch.manual_seed(1234)
model = pb.RationalFourierOperator3D(modes=(32, 32, 32), width=64, n_layers=3).to(dev)
model.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32")
x = torch.randn(8, 3, 128, 128, 128, device=dev)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    model.zero_grad(set_to_none=True)
    model(x).square().mean().backward()
torch.cuda.synchronize()
print("ok")
