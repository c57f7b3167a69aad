"""Probe: torch.fft.irfftn on a spectrum whose kz = 0 plane is not Hermitian in (kx, ky) -- what the reference feeds it
(only kx, ky >= 0 retained, operators/rational_fourier.py:116-117,170).  pocketfft (CPU) computes C2C over x, y then C2R over z,
which ignores the anti-Hermitian part; a single 3-D cuFFT C2R plan is free to do otherwise.  Compares, per dtype:
CUDA irfftn vs CPU irfftn, and CUDA irfftn vs the explicit two-step (ifft2 over x, y then irfft over z) on CUDA."""
import json
import sys
import torch


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())


out = {}
for n, m in ((16, 4), (128, 32)):
    torch.manual_seed(0)
    mh = m // 2 + 1
    spec = torch.zeros(2, n, n, n // 2 + 1, dtype=torch.complex128)
    spec[:, :m, :m, :mh] = torch.randn(2, m, m, mh, dtype=torch.complex128)
    for dt in (torch.complex64, torch.complex128):
        s_cpu = spec.to(dt)
        s_gpu = s_cpu.cuda()
        cpu = torch.fft.irfftn(s_cpu, s=(n, n, n), dim=[-3, -2, -1])
        gpu = torch.fft.irfftn(s_gpu, s=(n, n, n), dim=[-3, -2, -1])
        two = torch.fft.irfft(torch.fft.ifft2(s_gpu, dim=[-3, -2]), n=n, dim=-1)
        herm = s_gpu.clone()                       # same input with the kz = 0 plane made Hermitian by hand
        p0 = herm[..., 0]
        flipped = torch.roll(torch.flip(p0, dims=[-2, -1]), shifts=(1, 1), dims=(-2, -1)).conj()
        herm[..., 0] = 0.5 * (p0 + flipped)
        gpu_h = torch.fft.irfftn(herm, s=(n, n, n), dim=[-3, -2, -1])
        out[f"n{n}_m{m}_{str(dt).split('.')[-1]}"] = {
            "cuda_irfftn_vs_cpu_irfftn": rel(gpu, cpu), "cuda_two_step_vs_cpu_irfftn": rel(two, cpu),
            "cuda_irfftn_of_symmetrised_input_vs_cpu_irfftn": rel(gpu_h, cpu)}
print(json.dumps(out, indent=1))
json.dump(out, open(sys.argv[1], "w"), indent=1) if len(sys.argv) > 1 else None
