import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pde_fluid_phi_b200 import functional as pf, _cabi
torch.manual_seed(0)
dev = "cuda"
B, C, n = 8, 64, 128
h = torch.randn(B, C, n, n, n, device=dev); spec = torch.randn_like(h); g = torch.randn_like(h)
W = (torch.randn(C, C, device=dev) * 0.1).contiguous(); b = torch.randn(C, device=dev)
for _ in range(3):
    u, ho = pf.block_tail_fwd(h, spec, W, b)
    gu, t, dW, db = pf.block_tail_bwd(g, u, h, W, True)
torch.cuda.synchronize()
_cabi.profile(enable=True, reset=True)
for _ in range(10):
    u, ho = pf.block_tail_fwd(h, spec, W, b)
    gu, t, dW, db = pf.block_tail_bwd(g, u, h, W, True)
torch.cuda.synchronize()
print(os.environ.get("PDEPHI_BLOCK_DBG", "0"), {k: round(v[1] / v[0] * 1e3, 1) for k, v in _cabi.profile_report().items()})
