"""One fused block backward chain (middle-block variant) at the cfg-2 shape, three times: the program `ncu --set full -k
regex:block_tail_bwd` captures (scripts/_r2m.sh)."""
import sys
import torch
sys.path.insert(0, ".")
import pde_fluid_phi_b200 as pb
from pde_fluid_phi_b200 import functional as Fn

dev = torch.device("cuda:0")
pb.load_library()
B, C, grid, modes_h = 8, 64, (128, 128, 128), (32, 32, 17)
torch.manual_seed(0)
u = torch.randn(B, C, *grid, device=dev)
g = torch.randn(B, C, *grid, device=dev)
h = torch.randn(B, C, *grid, device=dev)
Wc = torch.randn(C, C, device=dev) * 0.1
for _ in range(3):
    out = Fn.block_spectral_bwd(g, u, h, Wc, modes_h, True)
torch.cuda.synchronize()
print("ok", float(out[3].abs().sum()))
