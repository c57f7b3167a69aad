"""GPU debug helper: tcgen05 transform stages vs the float64 spec (prints relative L2 errors)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pde_fluid_phi_b200 import functional as pf, _cabi
from oracle import dense_fp64 as d64
from oracle.spectral_oracle import rel_l2

def run(grid, modes, bc=(1, 2)):
    mh = (modes[0], modes[1], modes[2] // 2 + 1)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((*bc, *grid))
    Z = rng.standard_normal((*bc, *mh)) + 1j * rng.standard_normal((*bc, *mh))
    xd = torch.from_numpy(x).float().cuda()
    Zd = torch.from_numpy(Z).to(torch.complex64).cuda()
    c = d64.hermitian_weights(grid[2], mh[2]); n = float(np.prod(grid))
    for prec in (_cabi.PREC_FP32, _cabi.PREC_TF32):
        a0 = pf.analysis(xd, mh, _cabi.FORWARD, prec); torch.cuda.synchronize()
        a1 = pf.analysis(xd, mh, _cabi.ADJOINT, prec); torch.cuda.synchronize()
        s0 = pf.synthesis(Zd, grid, direction=_cabi.FORWARD, precision=prec); torch.cuda.synchronize()
        s1 = pf.synthesis(Zd, grid, direction=_cabi.ADJOINT, precision=prec); torch.cuda.synchronize()
        print(grid, modes, "prec", prec,
              "A fwd %.2e" % rel_l2(a0.cpu(), torch.from_numpy(d64.analysis(x, mh))),
              "A adj %.2e" % rel_l2(a1.cpu(), torch.from_numpy(d64.analysis(x, mh) * c / n)),
              "S fwd %.2e" % rel_l2(s0.cpu().double(), torch.from_numpy(d64.synthesis(Z, grid))),
              "S adj %.2e" % rel_l2(s1.cpu().double(), torch.from_numpy(d64.synthesis(Z, grid, hermitian=False, normalise=False))),
              flush=True)

if __name__ == "__main__":
    print("TF32 comp env:", os.environ.get("PDEPHI_TF32_COMP"))
    run((2, 128, 128), (2, 32, 32))
    run((4, 128, 64), (4, 16, 24), bc=(2, 3))
    run((16, 128, 128), (8, 32, 32), bc=(2, 40))      # > 148 planes per launch: exercises the persistent loop
