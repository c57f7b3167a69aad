"""GPU helper: per-parameter gradient agreement of the cfg-2 model between the routes (tf32 with / without the fused end
blocks, tf32x3-grade 'auto')."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_fluid_phi_b200 as pb
sys.argv = sys.argv[:1]
import bench

dev = torch.device("cuda")
torch.manual_seed(1234)
model = pb.RationalFourierOperator3D(modes=(32, 32, 32), width=64, n_layers=4, rational_order=(4, 4)).to(dev)
with torch.no_grad():
    for layer in model.rational_layers:
        q0 = layer.Q_coeffs[..., 0, 0, 0].clone()
        layer.Q_coeffs.mul_(1e-4)
        layer.Q_coeffs[..., 0, 0, 0] = q0
        layer.P_coeffs.mul_(1e-4)
B = 4
x = bench.turbulence_batch(B, 128, 42, dev)
y = bench.turbulence_batch(B, 128, 1042, dev)
def compare(tag):
    res = {}
    for name, prec, fl, fh in (("auto", "auto", False, False), ("tf32 plain", "tf32", False, False), ("tf32 head", "tf32", False, True),
                           ("tf32 lift", "tf32", True, False), ("tf32 both", "tf32", True, True)):
        model.set_precision(prec)
        model.fuse_lift, model.fuse_head = fl, fh
        model.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(model(x), y)
        loss.backward()
        torch.cuda.synchronize()
        res[name] = (float(loss.detach()), {n: p.grad.clone() for n, p in model.named_parameters()})
    ref = res["auto"]
    for name, (loss, grads) in res.items():
        worst = sorted(((float((grads[n] - ref[1][n]).norm() / ref[1][n].norm().clamp(min=1e-30)), n) for n in grads), reverse=True)[:4]
        print("%s %-11s loss %.7f  worst grad rel diffs vs auto: %s" % (tag, name, loss, ", ".join("%s %.1e" % (n, e) for e, n in worst)), flush=True)


compare("init     ")
# away from the initial point: biases non-zero (the DC terms of the mode-space lift), every weight perturbed by 20 %
torch.manual_seed(7)
with torch.no_grad():
    for n, p in model.named_parameters():
        if "Q_coeffs" in n:
            continue                      # keep R(k) away from its poles
        p.add_(0.2 * p.abs().mean().clamp(min=1e-3) * torch.randn_like(p))
compare("perturbed")
