"""GPU helper: fwd+bwd time of one spectral layer at the cfg-2 / cfg-1 shapes, with the per-kernel breakdown."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_fluid_phi_b200 as pb
from pde_fluid_phi_b200 import _cabi

def time_layer(layer, x, iters=20, warm=5):
    g = torch.randn_like(x)
    def step():
        x.grad = None
        layer.zero_grad(set_to_none=True)
        out = layer(x)
        out.backward(g)
    for _ in range(warm): step()
    torch.cuda.synchronize()
    _cabi.profile(enable=True, reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): step()
    e1.record(); torch.cuda.synchronize()
    prof = _cabi.profile_report()
    _cabi.profile(enable=False, reset=True)
    return e0.elapsed_time(e1) / iters * 1e3, {k: round(v[1] / v[0] * 1e3, 1) for k, v in prof.items()}

if __name__ == "__main__":
    prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    dev = "cuda"
    torch.manual_seed(0)
    res = {}
    if "--gen-only" not in sys.argv:
        rfl = pb.RationalFourierLayer(64, 64, (32, 32, 32)).to(dev); rfl.precision = prec
        x = torch.randn(8, 64, 128, 128, 128, device=dev, requires_grad=True)
        res["RationalFourierLayer cfg2 [8,64,128^3] modes 32^3"] = time_layer(rfl, x, iters)
        sc = pb.SpectralConv3D(64, 64, (32, 32, 32)).to(dev); sc.precision = prec
        res["SpectralConv3D cfg2-shape [8,64,128^3] modes 32^3"] = time_layer(sc, x, iters)
        del x
    sc1 = pb.SpectralConv3D(32, 32, (8, 8, 8)).to(dev); sc1.precision = prec
    x1 = torch.randn(4, 32, 32, 32, 32, device=dev, requires_grad=True)
    res[f"SpectralConv3D cfg1 [4,32,32^3] modes 8^3 ({prec})"] = time_layer(sc1, x1, iters)
    rf1 = pb.RationalFourierLayer(32, 32, (8, 8, 8)).to(dev); rf1.precision = prec
    res[f"RationalFourierLayer cfg1-shape [4,32,32^3] modes 8^3 ({prec})"] = time_layer(rf1, x1, iters)
    if "--gen" in sys.argv:
        def conditioned(layer):
            with torch.no_grad():
                q0 = layer.Q_coeffs[..., 0, 0, 0].clone(); layer.Q_coeffs.mul_(1e-6); layer.Q_coeffs[..., 0, 0, 0] = q0
            layer.precision = prec
            return layer
        del x1
        x3 = torch.randn(64, 64, 256, 256, 1, device=dev, requires_grad=True)
        sc3 = pb.SpectralConv3D(64, 64, (64, 64, 1)).to(dev); sc3.precision = prec
        res["SpectralConv3D cfg3 [64,64,256,256,1] modes (64,64,1)"] = time_layer(sc3, x3, iters)
        del x3
        x4 = torch.randn(1, 64, 128, 128, 128, device=dev, requires_grad=True)
        res["RationalFourierLayer cfg4-small [1,64,128^3] modes 64^3"] = time_layer(conditioned(pb.RationalFourierLayer(64, 64, (64, 64, 64)).to(dev)), x4, iters)
        del x4
        x5 = torch.randn(1, 64, 256, 256, 256, device=dev, requires_grad=True)
        res["RationalFourierLayer cfg5 [1,64,256^3] modes 64^3"] = time_layer(conditioned(pb.RationalFourierLayer(64, 64, (64, 64, 64)).to(dev)), x5, iters)
    for k, (us, kern) in res.items():
        print(f"{k}: {us:.1f} us fwd+bwd   per-kernel us: {kern}")
