"""GPU debug helper for the fused block tail kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pde_fluid_phi_b200 import functional as pf, _cabi
from oracle.spectral_oracle import rel_l2
_cabi.set_option(_cabi.OPT_BLOCK_SAVES_DERIVATIVE, 0)      # this helper inspects the pre-activation itself in `u`
torch.manual_seed(0)
B, C, grid = 2, 64, (8, 8, 16)
dev = "cuda"
h = torch.randn(B, C, *grid, device=dev); spec = torch.randn(B, C, *grid, device=dev); g = torch.randn(B, C, *grid, device=dev)
def ref_fwd(W, b):
    conv = torch.einsum("oc,bcxyz->boxyz", W.double(), h.double()) + b.double()[None, :, None, None, None]
    u = spec.double() + conv + h.double()
    return u, torch.nn.functional.gelu(u)
for name, W in [("zero", torch.zeros(C, C, device=dev)), ("identity", torch.eye(C, device=dev)),
                ("shift", torch.roll(torch.eye(C, device=dev), 1, 0)), ("rand", torch.randn(C, C, device=dev) * 0.1)]:
    b = torch.randn(C, device=dev)
    u, ho = pf.block_tail_fwd(h, spec, W.contiguous(), b)
    torch.cuda.synchronize()
    ur, hr = ref_fwd(W, b)
    print(name, "u err %.3e" % rel_l2(u.cpu().double(), ur.cpu()), "hout err %.3e" % rel_l2(ho.cpu().double(), hr.cpu()),
          "| u vs (spec+h+b) %.3e" % rel_l2(u.cpu().double(), (spec.double() + h.double() + b.double()[None, :, None, None, None]).cpu()), flush=True)
    # backward
    gu, t, dW, db = pf.block_tail_bwd(g, ur.float().contiguous(), h, W.contiguous(), True)
    torch.cuda.synchronize()
    ud = ur.clone().requires_grad_(True)
    torch.nn.functional.gelu(ud).backward(g.double())
    gur = ud.grad
    tr = gur + torch.einsum("oc,boxyz->bcxyz", W.double(), gur)
    dWr = torch.einsum("boxyz,bcxyz->oc", gur, h.double())
    dbr = gur.sum(dim=(0, 2, 3, 4))
    print("   bwd: gu %.3e t %.3e dW %.3e db %.3e" % (rel_l2(gu.cpu().double(), gur.cpu()), rel_l2(t.cpu().double(), tr.cpu()),
                                                      rel_l2(dW.cpu().double(), dWr.cpu()), rel_l2(db.cpu().double(), dbr.cpu())), flush=True)
