"""One fused block forward (middle-block variant) at the cfg-2 shape, three times: the program `ncu --set full -k
regex:block_tail_fwd` captures."""
import sys
import torch
sys.path.insert(0, ".")
import pde_fluid_phi_b200 as pb
from pde_fluid_phi_b200 import functional as Fn

dev = torch.device("cuda:0")
pb.load_library()
B, C, grid, mh = 8, 64, (128, 128, 128), (32, 32, 17)
torch.manual_seed(0)
h = torch.randn(B, C, *grid, device=dev)
Y = torch.randn(B, C, *mh, device=dev, dtype=torch.complex64) * 10
ms = torch.rand(mh, device=dev).reshape(-1) + 0.5
rs = torch.rand(B * C, 4, device=dev) + 0.5
Wc = torch.randn(C, C, device=dev) * 0.1
bc = torch.randn(C, device=dev)
for _ in range(3):
    out = Fn.block_spectral_fwd(Y, ms, rs[:, 2], h, Wc, bc)
torch.cuda.synchronize()
print("ok", float(out[1].abs().sum()))
