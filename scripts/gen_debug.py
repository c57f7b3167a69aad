"""GPU debug helper: the per-axis tensor-core route (transforms_gen.cu) against torch.fft in float64 on the same device."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pde_fluid_phi_b200 import functional as pf, _cabi


def rel(a, b):
    a, b = a.to(torch.complex128) if a.is_complex() else a.double(), b
    return float((a - b).norm() / b.norm().clamp(min=1e-300))


def run(grid, modes, bc=(1, 2), prec=_cabi.PREC_TF32, timing=False):
    mh = (modes[0], modes[1], modes[2] // 2 + 1)
    torch.manual_seed(0)
    x = torch.randn(*bc, *grid, device="cuda")
    Z = torch.randn(*bc, *mh, device="cuda", dtype=torch.complex64)
    sl = (slice(None), slice(None), slice(0, mh[0]), slice(0, mh[1]), slice(0, mh[2]))
    ref_a = torch.fft.rfftn(x.double(), dim=[-3, -2, -1])[sl]
    full = torch.zeros(*bc, grid[0], grid[1], grid[2] // 2 + 1, dtype=torch.complex128, device="cuda")
    full[sl] = Z
    ref_s = torch.fft.irfftn(full, s=grid, dim=[-3, -2, -1])
    if grid[2] == 1:      # cuFFT's multi-dimensional C2R on a non-Hermitian kz = 0 plane: use the CPU semantics (Re of the C2C inverse)
        ref_s = torch.fft.ifftn(full, dim=[-3, -2]).real
    n = float(grid[0] * grid[1] * grid[2])
    c = torch.full((mh[2],), 2.0, dtype=torch.float64, device="cuda")
    c[0] = 1.0
    if grid[2] % 2 == 0 and mh[2] - 1 == grid[2] // 2:
        c[-1] = 1.0
    a0 = pf.analysis(x, mh, _cabi.FORWARD, prec)
    a1 = pf.analysis(x, mh, _cabi.ADJOINT, prec)
    s0 = pf.synthesis(Z, grid, direction=_cabi.FORWARD, precision=prec)
    add0, add1 = torch.randn_like(x), torch.randn_like(x)
    ms = torch.rand(mh, device="cuda") + 0.5
    rs = torch.rand(bc[0] * bc[1], 4, device="cuda") + 0.5
    s2 = pf.synthesis(Z, grid, ms.reshape(-1), rs[:, 2], add0, add1, _cabi.FORWARD, prec)
    full2 = torch.zeros_like(full)
    full2[sl] = Z * ms * rs[:, 2].reshape(*bc, 1, 1, 1)
    ref_s2 = (torch.fft.ifftn(full2, dim=[-3, -2]).real if grid[2] == 1 else torch.fft.irfftn(full2, s=grid, dim=[-3, -2, -1]))
    s2 = s2 - add0 - add1       # compare the transform part only (fp32 cancellation noise ~1e-7 of the addends)
    # adjointness: <A x, Z> == <x, S_adj Z>
    s1 = pf.synthesis(Z, grid, direction=_cabi.ADJOINT, precision=prec)
    lhs = (ref_a.conj() * Z.to(torch.complex128)).sum().real
    rhs = (x.double() * s1.double()).sum()
    torch.cuda.synchronize()
    print(grid, modes, bc, "prec", prec,
          "A fwd %.2e" % rel(a0, ref_a), "A adj %.2e" % rel(a1, ref_a * c / n),
          "S fwd %.2e" % rel(s0, ref_s), "(fp32 route %.2e)" % rel(pf.synthesis(Z, grid, precision=_cabi.PREC_FP32), ref_s), "S fused %.2e" % rel(s2, ref_s2),
          "S adj (adjointness) %.2e" % (abs(float(lhs - rhs)) / abs(float(lhs))), flush=True)
    if timing:
        _cabi.profile(True, True)
        for _ in range(5):
            pf.analysis(x, mh, _cabi.FORWARD, prec)
            pf.synthesis(Z, grid, None, None, add0, None, _cabi.FORWARD, prec)
        torch.cuda.synchronize()
        rep = _cabi.profile_report()
        _cabi.profile(False, True)
        gb = x.numel() * 4 / 1e9
        for k, (nl, ms_) in rep.items():
            print("    %-32s %3d launches  %8.3f ms each   (activation %.2f GB)" % (k, nl, ms_ / nl, gb), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "small"):
        run((128, 128, 32), (5, 8, 9))
        run((128, 128, 128), (64, 64, 64))
        run((128, 256, 64), (10, 40, 30), bc=(2, 3))
        run((128, 128, 1), (32, 16, 1), bc=(2, 3))
        run((256, 256, 1), (64, 64, 1), bc=(2, 4))
        run((256, 128, 96), (33, 20, 40), bc=(1, 3))
    if which in ("all", "big"):
        run((128, 128, 128), (64, 64, 64), bc=(1, 64), timing=True)
        run((256, 256, 256), (64, 64, 64), bc=(1, 16), timing=True)
        run((256, 256, 1), (64, 64, 1), bc=(16, 64), timing=True)
