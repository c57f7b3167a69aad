"""GPU helper: accuracy and time of the split-TF32 (fp32-grade) transforms against float64 torch.fft."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pde_fluid_phi_b200 import _cabi, functional as pf

def bias(a, b):
    a = a.double() if not a.is_complex() else a.to(torch.complex128)
    b = b.double() if not b.is_complex() else b.to(torch.complex128)
    return float(((a.conj() * b).sum().real / (b.conj() * b).sum().real) - 1.0)

def rel(a, b):
    a = a.double() if not a.is_complex() else a.to(torch.complex128)
    b = b.double() if not b.is_complex() else b.to(torch.complex128)
    return float((a - b).abs().pow(2).sum().sqrt() / b.abs().pow(2).sum().sqrt())

dev = "cuda"
torch.manual_seed(0)
for grid, modes in [((2, 128, 128), (2, 32, 32)), ((4, 128, 64), (4, 16, 24)), ((128, 128, 128), (32, 32, 32)), ((160, 128, 128), (8, 20, 38))]:
    mh = (modes[0], modes[1], modes[2] // 2 + 1)
    x = torch.randn(2, 3, *grid, device=dev)
    Z = torch.randn(2, 3, *mh, device=dev, dtype=torch.complex64)
    ref = torch.fft.rfftn(x.double(), dim=[-3, -2, -1])[:, :, :mh[0], :mh[1], :mh[2]]
    full = torch.zeros(2, 3, grid[0], grid[1], grid[2] // 2 + 1, dtype=torch.complex128, device=dev)
    full[:, :, :mh[0], :mh[1], :mh[2]] = Z
    refs = torch.fft.irfftn(full, s=grid, dim=[-3, -2, -1])
    a0 = torch.randn(2, 3, *grid, device=dev)
    for name, prec in [("fp32", _cabi.PREC_FP32), ("tf32", _cabi.PREC_TF32), ("tf32x3", _cabi.PREC_TF32X3)]:
        try:
            A = pf.analysis(x, mh, precision=prec); S = pf.synthesis(Z, grid, precision=prec)
            ea = rel(A, ref)
            es = rel(S, refs)
            print(f"   bias analysis {bias(A, ref):+.2e} synthesis {bias(S, refs):+.2e}")
            es2 = rel(pf.synthesis(Z, grid, addend0=a0, precision=prec), refs + a0.double())
            torch.cuda.synchronize()
            print(f"{grid} {modes} {name:7s} analysis {ea:.3e}  synthesis {es:.3e}  (+addend) {es2:.3e}", flush=True)
        except NotImplementedError as e:
            print(grid, modes, name, "unsupported")
# time at the cfg-2 shape
grid, mh = (128, 128, 128), (32, 32, 17)
x = torch.randn(8, 64, *grid, device=dev)
Z = torch.randn(8, 64, *mh, device=dev, dtype=torch.complex64)
for name, prec in [("tf32", _cabi.PREC_TF32), ("tf32x3", _cabi.PREC_TF32X3)]:
    for _ in range(3):
        pf.analysis(x, mh, precision=prec); pf.synthesis(Z, grid, precision=prec)
    torch.cuda.synchronize()
    _cabi.profile(enable=True, reset=True)
    for _ in range(10):
        pf.analysis(x, mh, precision=prec); o = pf.synthesis(Z, grid, precision=prec)
    torch.cuda.synchronize()
    rep = _cabi.profile_report()
    print(name, {k: round(v[1] / v[0] * 1e3, 1) for k, v in rep.items()}, flush=True)
    _cabi.profile(enable=True, reset=True)
    for _ in range(10):
        o2 = pf.synthesis(Z, grid, addend0=x, precision=prec)
    torch.cuda.synchronize()
    rep = _cabi.profile_report()
    _cabi.profile(enable=False, reset=True)
    print(name, "with addend", {k: round(v[1] / v[0] * 1e3, 1) for k, v in rep.items()}, flush=True)
