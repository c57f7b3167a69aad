"""Fused block backward (pdephi_block_spectral_bwd) against the unfused pair block_tail_bwd + adjoint analysis, op level:
t / dWc / dbc / G must agree (same arithmetic), dZ to the TF32 gate; then the timing of both chains at the cfg-2 shape."""
import json
import sys
import torch

sys.path.insert(0, ".")
import pde_fluid_phi_b200 as pb
from pde_fluid_phi_b200 import _cabi, functional as Fn


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def case(B, grid, modes_h, variant, dev):
    torch.manual_seed(7)
    C = 64
    u = Fn.pack_block_derivative(torch.randn(B, C, *grid, device=dev))      # gelu'(u) as the forward kernels leave it
    Wc = torch.randn(C, C, device=dev) * 0.1
    head_w = lift = None
    write_t = True
    if variant == "head":
        head_w = torch.randn(3, C, device=dev) * 0.3
        g = torch.randn(B, 3, *grid, device=dev)
        h = torch.randn(B, C, *grid, device=dev)
    elif variant in ("lift", "lift_t"):
        lift = (torch.randn(C, 3, device=dev) * 0.3, torch.randn(C, device=dev) * 0.1)
        g = torch.randn(B, C, *grid, device=dev)
        h = torch.randn(B, 3, *grid, device=dev)
        write_t = variant == "lift_t"
    else:
        g = torch.randn(B, C, *grid, device=dev)
        h = torch.randn(B, C, *grid, device=dev)
    gu, t0, dW0, db0, *rest = Fn.block_tail_bwd(g, u, h, Wc, True, head_w, write_t=write_t, lift=lift)
    dZ_ref = Fn.analysis(gu, modes_h, _cabi.ADJOINT, _cabi.PREC_TF32X3)
    dZ_old = Fn.analysis(gu, modes_h, _cabi.ADJOINT, _cabi.PREC_TF32)
    t1, dW1, db1, dZ1, G1 = Fn.block_spectral_bwd(g, u, h, Wc, modes_h, True, head_w, write_t=write_t, lift=lift)
    torch.cuda.synchronize()
    out = {"variant": variant, "B": B, "grid": grid, "modes_h": modes_h,
           "dZ_fused_vs_fp32grade": rel(torch.view_as_real(dZ1), torch.view_as_real(dZ_ref)),
           "dZ_unfused_tf32_vs_fp32grade": rel(torch.view_as_real(dZ_old), torch.view_as_real(dZ_ref)),
           "dWc": rel(dW1, dW0), "dbc": rel(db1, db0)}
    if write_t:
        out["t"] = rel(t1, t0)
    if rest:
        out["G"] = rel(G1, rest[0])
    return out


def timing(dev, reps=5):
    B, C, grid, modes_h = 8, 64, (128, 128, 128), (32, 32, 17)
    u = Fn.pack_block_derivative(torch.randn(B, C, *grid, device=dev))      # gelu'(u) as the forward kernels leave it
    g = torch.randn(B, C, *grid, device=dev)
    h = torch.randn(B, C, *grid, device=dev)
    Wc = torch.randn(C, C, device=dev) * 0.1
    res = {}
    for name in ("unfused", "fused"):
        def run():
            if name == "unfused":
                gu, t, dW, db = Fn.block_tail_bwd(g, u, h, Wc, True)
                return Fn.analysis(gu, modes_h, _cabi.ADJOINT, _cabi.PREC_TF32)
            return Fn.block_spectral_bwd(g, u, h, Wc, modes_h, True)[3]
        for _ in range(2):
            run()
        _cabi.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        rep = _cabi.profile_report()
        _cabi.profile(False)
        res[name] = {"ms_per_chain": e0.elapsed_time(e1) / reps,
                     "kernels_ms": {k: round(v[1] / reps, 4) for k, v in rep.items()}}
    return res


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    pb.load_library()
    results = []
    for variant in ("middle", "head", "lift", "lift_t"):
        r = case(2, (32, 128, 128), (16, 32, 17), variant, dev)
        print(json.dumps(r), flush=True)
        results.append(r)
    r = case(1, (128, 128, 128), (32, 32, 17), "middle", dev)
    print(json.dumps(r), flush=True)
    results.append(r)
    tm = timing(dev)
    print(json.dumps(tm, indent=1))
    json.dump({"parity": results, "timing": tm}, open("gpurun_out/fused_bwd_debug.json", "w"), indent=1)
