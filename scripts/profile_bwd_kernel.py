"""One fused block backward chain at the cfg-2 shape, three times: the program `ncu --set full -k regex:block_tail_bwd`
captures.  usage: python scripts/profile_bwd_kernel.py [middle|first|last]"""
import sys
import torch
sys.path.insert(0, ".")
import pde_fluid_phi_b200 as pb
from pde_fluid_phi_b200 import functional as Fn

dev = torch.device("cuda:0")
pb.load_library()
which = sys.argv[1] if len(sys.argv) > 1 else "middle"
B, C, grid, modes_h = 8, 64, (128, 128, 128), (32, 32, 17)
torch.manual_seed(0)
u = Fn.pack_block_derivative(torch.rand(B, C, *grid, device=dev))      # what the forward kernels leave (option value 2)
g = torch.randn(B, 3 if which == "last" else C, *grid, device=dev)
h = torch.randn(B, 3 if which == "first" else C, *grid, device=dev)
Wc = torch.randn(C, C, device=dev) * 0.1
kw = {}
if which == "last":
    kw = dict(head_w=torch.randn(3, C, device=dev) * 0.3)
if which == "first":
    kw = dict(write_t=False, lift=(torch.randn(C, 3, device=dev) * 0.3, torch.randn(C, device=dev) * 0.1))
for _ in range(3):
    out = Fn.block_spectral_bwd(g, u, h, Wc, modes_h, True, **kw)
torch.cuda.synchronize()
print("ok", which, float(out[3].abs().sum()))
