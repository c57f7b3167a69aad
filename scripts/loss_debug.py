import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_fluid_phi_b200 as pb
from bench import turbulence_batch, CFG
dev = torch.device("cuda")
for pscale in (1e-2, 1e-4):
    torch.manual_seed(1234)
    model = pb.RationalFourierOperator3D(modes=CFG["modes"], width=64, n_layers=4).to(dev).set_precision("tf32")
    with torch.no_grad():
        for layer in model.rational_layers:
            q0 = layer.Q_coeffs[..., 0, 0, 0].clone(); layer.Q_coeffs.mul_(1e-4); layer.Q_coeffs[..., 0, 0, 0] = q0
            layer.P_coeffs.mul_(pscale)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-6, weight_decay=1e-4)
    x = turbulence_batch(2, 128, 42, dev); y = turbulence_batch(2, 128, 43, dev)
    for i in range(12):
        opt.zero_grad(set_to_none=True)
        out = model(x)
        loss = torch.nn.functional.mse_loss(out, y); loss.backward()
        gn = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        for g in opt.param_groups: g["lr"] = min(1.0, (i + 2) / 1000.0) * 1e-3
        print(pscale, i, float(loss), float(gn), float(out.abs().max()), flush=True)
