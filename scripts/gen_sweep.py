"""GPU helper: randomised shape sweep of the per-axis tensor-core route (TF32) against the FFMA route (fp32)."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pde_fluid_phi_b200 import functional as pf, _cabi


def rel(a, b):
    a = a.to(torch.complex128) if a.is_complex() else a.double()
    b = b.to(torch.complex128) if b.is_complex() else b.double()
    return float((a - b).norm() / b.norm().clamp(min=1e-300))


def main(n_cases, seed):
    rnd = random.Random(seed)
    lib = _cabi.load()
    axes = [32, 64, 96, 128, 160, 192, 256]
    done = bad = 0
    while done < n_cases:
        two_d = rnd.random() < 0.25
        Nx, Ny = rnd.choice(axes), rnd.choice(axes)
        Nz = 1 if two_d else rnd.choice([32, 64, 96, 128, 256])
        if Nx * Ny * Nz > 256 * 256 * 128:
            continue
        mx, my = rnd.randint(1, min(64, Nx)), rnd.randint(1, min(64, Ny))
        mzh = 1 if two_d else rnd.randint(1, min(40, Nz // 2 + 1))
        plan = _cabi.get_plan(0, (Nx, Ny, Nz), (mx, my, mzh))
        if not lib.pdephi_plan_supports(plan, _cabi.PREC_TF32):
            continue
        B, C = rnd.randint(1, 2), rnd.randint(1, 3)
        grid, mh = (Nx, Ny, Nz), (mx, my, mzh)
        g = torch.Generator(device="cuda").manual_seed(rnd.randint(0, 1 << 30))
        x = torch.randn(B, C, *grid, device="cuda", generator=g)
        Z = torch.complex(torch.randn(B, C, *mh, device="cuda", generator=g), torch.randn(B, C, *mh, device="cuda", generator=g))
        a0 = torch.randn(B, C, *grid, device="cuda", generator=g)
        ms = torch.rand(mh, device="cuda", generator=g).reshape(-1) + 0.5
        rs = torch.rand(B * C, 4, device="cuda", generator=g) + 0.5
        errs = []
        for d in (_cabi.FORWARD, _cabi.ADJOINT):
            errs.append(rel(pf.analysis(x, mh, d, _cabi.PREC_TF32), pf.analysis(x, mh, d, _cabi.PREC_FP32)))
            want = pf.synthesis(Z, grid, ms, rs[:, 2], a0, None, d, _cabi.PREC_FP32) - a0
            got = pf.synthesis(Z, grid, ms, rs[:, 2], a0, None, d, _cabi.PREC_TF32) - a0
            errs.append(rel(got, want))
            errs.append(rel(pf.synthesis(Z, grid, None, None, None, None, d, _cabi.PREC_TF32),
                            pf.synthesis(Z, grid, None, None, None, None, d, _cabi.PREC_FP32)))
        torch.cuda.synchronize()
        ok = max(errs) < 2e-3
        bad += (not ok)
        done += 1
        print(("ok  " if ok else "FAIL"), grid, mh, (B, C), " ".join("%.1e" % e for e in errs), flush=True)
    print("cases", done, "failures", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]) if len(sys.argv) > 1 else 40, int(sys.argv[2]) if len(sys.argv) > 2 else 0) else 0)
