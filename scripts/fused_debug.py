"""GPU debug helper: fused spectral-synthesis + block-tail forward (pdephi_block_spectral_fwd) vs synthesis + block_tail_fwd."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pde_fluid_phi_b200 import functional as pf, _cabi

def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())

def run(B, grid, modes, iters=0):
    mh = (modes[0], modes[1], modes[2] // 2 + 1)
    torch.manual_seed(0)
    C = 64
    h = torch.randn(B, C, *grid, device="cuda")
    Y = torch.randn(B, C, *mh, device="cuda", dtype=torch.complex64) * 100.0
    Wc = torch.randn(C, C, device="cuda") * 0.1
    bc = torch.randn(C, device="cuda")
    mask = (torch.rand(mh, device="cuda") + 0.5).reshape(-1).contiguous()
    rs = torch.rand(B * C, 4, device="cuda") + 0.5
    assert pf.block_spectral_supported(h, mh, _cabi.PREC_TF32), "fused route not supported for this shape"
    spec = pf.synthesis(Y, grid, mask, rs[:, 2], None, None, _cabi.FORWARD, _cabi.PREC_FP32)
    u_ref = spec.double() + torch.einsum("oc,bcxyz->boxyz", Wc.double(), h.double()) + bc.double().view(1, C, 1, 1, 1) + h.double()
    u, hout = pf.block_spectral_fwd(Y, mask, rs[:, 2], h, Wc, bc)
    spec_tc = pf.synthesis(Y, grid, mask, rs[:, 2], None, None, _cabi.FORWARD, _cabi.PREC_TF32)
    u2, hout2 = pf.block_tail_fwd(h, spec_tc, Wc, bc)
    torch.cuda.synchronize()
    print(B, grid, modes, "fused u %.2e hout %.2e | unfused u %.2e | spec-only part: fused %.2e" % (
        rel(u, u_ref), rel(hout, torch.nn.functional.gelu(u_ref)), rel(u2, u_ref),
        rel(u.double() - (u_ref - spec.double()), spec.double())), flush=True)
    if iters:
        _cabi.profile(True, True)
        for _ in range(iters):
            pf.block_spectral_fwd(Y, mask, rs[:, 2], h, Wc, bc)
            s2 = pf.synthesis(Y, grid, mask, rs[:, 2], None, None, _cabi.FORWARD, _cabi.PREC_TF32)
            pf.block_tail_fwd(h, s2, Wc, bc)
        torch.cuda.synchronize()
        for k, (n, ms) in _cabi.profile_report().items():
            print("    %-32s %3d launches %8.3f ms each" % (k, n, ms / n))
        _cabi.profile(False, True)

if __name__ == "__main__":
    run(1, (32, 128, 128), (8, 16, 16))
    run(2, (128, 128, 128), (32, 32, 32))
    run(8, (128, 128, 128), (32, 32, 32), iters=5)
