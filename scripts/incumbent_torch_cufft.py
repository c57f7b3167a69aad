#!/usr/bin/env python
"""The incumbent on the same B200: the reference's op sequence in stock PyTorch eager (torch.fft -> cuFFT, einsum,
elementwise ops), timed beside the library at the cfg-2 workload ("also time the reference on
the same B200 (cuFFT path) -- that, not the CPU, is the incumbent the kernels must beat").

    python scripts/incumbent_torch_cufft.py [--steps 5] [--warmup 3] [--batch 8] [--out gpurun_out/incumbent.json]

The reference package cannot travel to the GPU box, so the module below restates it with plain torch ops in the
reference's own op order (operators/rational_fourier.py:73-172,185-276, operators/stability.py:59-103); it shares no code
with the product or with oracle/.  It measures
  (1) one RationalFourierLayer fwd+bwd at [B,64,128^3], modes 32^3 (CUDA events, median),
  (2) the bench.py training step (fwd + MSE + bwd + clip_grad_norm + AdamW) in samples/s,
  (3) outputs and parameter gradients of the two whole models on the same batch and weights (relative L2),
each for the stock incumbent and for pde_fluid_phi_b200.  A measurement script: nothing in the product imports it.
"""
from __future__ import annotations

import argparse
import itertools
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from bench import CFG, turbulence_batch  # noqa: E402


class EagerRationalLayer(nn.Module):
    """rfftn -> slice -> P(k)/Q(k) real mix -> zero-padded spectrum -> decay mask + energy rescale on a clone -> irfftn."""

    def __init__(self, width, modes, order=(4, 4), eps=1e-6):
        super().__init__()
        self.modes, self.eps = modes, eps
        p, q = order
        self.P_coeffs = nn.Parameter(torch.randn(width, width, p, p, p) * 0.02)
        self.Q_coeffs = nn.Parameter(torch.randn(width, width, q, q, q) * 0.02)
        with torch.no_grad():
            self.Q_coeffs[..., 0, 0, 0] = self.Q_coeffs[..., 0, 0, 0].abs() + 1.0
        mx, my, mz = modes
        axes = [torch.arange(n, dtype=torch.float32) for n in (mx, my, mz // 2 + 1)]
        k = torch.stack(torch.meshgrid(*axes, indexing="ij"))
        self.register_buffer("k_grid", k)
        self.register_buffer("decay_mask", (1 + torch.sqrt(k[0] ** 2 + k[1] ** 2 + k[2] ** 2) ** 2) ** (-2.0 / 2))

    def poly(self, coeffs):
        kx, ky, kz = self.k_grid
        n = coeffs.shape[-1]
        acc = torch.zeros(coeffs.shape[:2] + kx.shape, device=coeffs.device, dtype=coeffs.dtype)
        for a, b, c in itertools.product(range(n), repeat=3):
            if a + b + c < n:
                acc = acc + coeffs[:, :, a, b, c].unsqueeze(-1).unsqueeze(-1).unsqueeze(-1) * ((kx ** a) * (ky ** b) * (kz ** c))
        return acc

    def forward(self, x):
        mx, my, mz = self.modes
        mzh = mz // 2 + 1
        x_ft = torch.fft.rfftn(x, dim=[-3, -2, -1])
        R = self.poly(self.P_coeffs) / (self.poly(self.Q_coeffs) + self.eps)
        kept = x_ft[:, :, :mx, :my, :mzh]
        out_ft = torch.zeros_like(x_ft)
        out_ft[:, :, :mx, :my, :mzh] = torch.einsum("bixyz,oixyz->boxyz", kept, R.to(kept.dtype))
        box = out_ft[:, :, :mx, :my, :mzh]
        damped = box * self.decay_mask
        e0 = torch.sum(torch.abs(box) ** 2, dim=(-3, -2, -1), keepdim=True)
        e1 = torch.sum(torch.abs(damped) ** 2, dim=(-3, -2, -1), keepdim=True)
        damped = damped * torch.sqrt((e0 + self.eps) / (e1 + self.eps))
        proj = out_ft.clone()
        proj[:, :, :mx, :my, :mzh] = damped
        return torch.fft.irfftn(proj, s=x.shape[-3:], dim=[-3, -2, -1])


class EagerSpectralConv(nn.Module):
    """operators/spectral_layers.py:46-78 with the weight contracted on [..., :mz//2+1] (the unmodified forward raises
    for mz > 2; SURVEY.md section 8c convention 2)."""

    def __init__(self, cin, cout, modes):
        super().__init__()
        self.modes, self.cout = modes, cout
        self.weights = nn.Parameter(torch.randn(cout, cin, *modes, dtype=torch.cfloat) * 0.02)

    def forward(self, x):
        mx, my, mz = self.modes
        mzh = mz // 2 + 1
        x_ft = torch.fft.rfftn(x, dim=[-3, -2, -1])
        out_modes = torch.einsum("bixyz,oixyz->boxyz", x_ft[:, :, :mx, :my, :mzh], self.weights[..., :mzh])
        out_ft = torch.zeros(x.shape[0], self.cout, *x_ft.shape[-3:], dtype=torch.cfloat, device=x.device)
        out_ft[:, :, :mx, :my, :mzh] = out_modes
        return torch.fft.irfftn(out_ft, s=x.shape[-3:], dim=[-3, -2, -1])


class EagerOperator(nn.Module):
    """rearrange -> Linear -> rearrange, n x gelu(spectral + conv1x1 + x), rearrange -> Linear -> rearrange."""

    def __init__(self, modes, width, n_layers, order, cin=3, cout=3):
        super().__init__()
        self.input_proj = nn.Linear(cin, width)
        self.output_proj = nn.Linear(width, cout)
        self.rational_layers = nn.ModuleList([EagerRationalLayer(width, modes, order) for _ in range(n_layers)])
        self.local_convs = nn.ModuleList([nn.Conv3d(width, width, kernel_size=1) for _ in range(n_layers)])

    def forward(self, x):
        h = self.input_proj(x.permute(0, 2, 3, 4, 1)).permute(0, 4, 1, 2, 3)
        for spectral, local in zip(self.rational_layers, self.local_convs):
            h = F.gelu(spectral(h) + local(h) + h)
        return self.output_proj(h.permute(0, 2, 3, 4, 1)).permute(0, 4, 1, 2, 3)


class EagerFNO(nn.Module):
    """models/fno3d.py:75-108: the same loop around SpectralConv3D (names as in the reference: spectral_layers, local_layers)."""

    def __init__(self, modes, width, n_layers, cin=3, cout=3):
        super().__init__()
        self.input_proj = nn.Linear(cin, width)
        self.output_proj = nn.Linear(width, cout)
        self.spectral_layers = nn.ModuleList([EagerSpectralConv(width, width, modes) for _ in range(n_layers)])
        self.local_layers = nn.ModuleList([nn.Conv3d(width, width, kernel_size=1) for _ in range(n_layers)])

    def forward(self, x):
        h = self.input_proj(x.permute(0, 2, 3, 4, 1)).permute(0, 4, 1, 2, 3)
        for spectral, local in zip(self.spectral_layers, self.local_layers):
            h = F.gelu(spectral(h) + local(h) + h)
        return self.output_proj(h.permute(0, 2, 3, 4, 1)).permute(0, 4, 1, 2, 3)


def condition(layers):
    """bench.py's weight set: non-constant Q and all P coefficients x1e-4 (the default init overflows)."""
    with torch.no_grad():
        for layer in layers:
            q0 = layer.Q_coeffs[..., 0, 0, 0].clone()
            layer.Q_coeffs.mul_(1e-4)
            layer.Q_coeffs[..., 0, 0, 0] = q0
            layer.P_coeffs.mul_(1e-4)


def event_times(fn, warmup, reps):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return out


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=CFG["batch_per_gpu"])
    ap.add_argument("--fp64", action="store_true", help="also run the incumbent in float64 (noise floors)")
    ap.add_argument("--parity-only", action="store_true")
    ap.add_argument("--layers", action="store_true", help="fwd+bwd microseconds of single layers at the BASELINE shapes")
    ap.add_argument("--big", action="store_true", help="with --layers: the 64^3-mode layers of configs 4 and 5 instead")
    ap.add_argument("--cfg3", action="store_true", help="training step of the 2-D Kolmogorov FNO (configs[2]) instead")
    ap.add_argument("--diagnose", action="store_true", help="where the incumbent's fp32 run leaves its fp64 run, module by module")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "incumbent.json"))
    args = ap.parse_args()
    import pde_fluid_phi_b200 as pb
    pb.load_library()
    dev = torch.device("cuda", 0)
    n, B, W = CFG["grid"], args.batch, CFG["width"]
    res = {"device": torch.cuda.get_device_name(0), "torch": torch.__version__, "batch": B,
           "cudnn_allow_tf32": torch.backends.cudnn.allow_tf32, "matmul_allow_tf32": torch.backends.cuda.matmul.allow_tf32}

    def flush():
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)

    if args.cfg3:
        cfg3_step(res, dev, pb, args)
        flush()
        print(json.dumps(res))
        return
    if args.layers:
        single_layers(res, dev, pb, args.big)
        flush()
        print(json.dumps(res))
        return
    x = turbulence_batch(B, n, 42, dev)
    y = turbulence_batch(B, n, 1042, dev)
    if args.diagnose:
        diagnose(res, x[:1], W, dev, pb)
        flush()
        print(json.dumps(res))
        return
    if not args.parity_only:
        layer_times(res, B, W, n, dev, pb)
        flush()
    whole_model(res, args, x, y, B, W, dev, pb, flush)
    print(json.dumps(res))


def cfg3_step(res, dev, pb, args):
    """configs[2]: FNO modes (64,64,1), width 64, 4 layers, 256 x 256 x 1 grid, batch 64: fwd + MSE + bwd + AdamW."""
    B, n = 64, 256
    torch.manual_seed(1234)
    x = torch.randn(B, 2, n, n, 1, device=dev)
    y = torch.roll(x, 3, dims=3)
    step = {}
    for who, make in (("incumbent_torch_cufft", lambda: EagerFNO((64, 64, 1), 64, 4, 2, 2)),
                      ("pde_fluid_phi_b200[tf32]", lambda: pb.FNO3D(modes=(64, 64, 1), width=64, n_layers=4, in_channels=2,
                                                                     out_channels=2).set_precision("tf32"))):
        model = make().to(dev)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-5, weight_decay=1e-4)

        def run():
            opt.zero_grad(set_to_none=True)
            F.mse_loss(model(x), y).backward()
            opt.step()
        torch.cuda.reset_peak_memory_stats()
        ms = statistics.median(event_times(run, args.warmup, args.steps))
        step[who] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3, "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
        del model, opt
        torch.cuda.empty_cache()
    res["cfg3_train_step"] = step


def single_layers(res, dev, pb, big=False):
    """The 'SpectralConv3d fwd+bwd us' half of the BASELINE metric: incumbent vs library, median over CUDA-event-timed calls."""
    def conditioned(layer):
        condition([layer])
        return layer

    cases = [
        ("SpectralConv3D [4,32,32^3] modes 8^3 (cfg 1)", (4, 32, 32, 32, 32), lambda: EagerSpectralConv(32, 32, (8, 8, 8)),
         lambda: pb.SpectralConv3D(32, 32, (8, 8, 8)), "auto", 20, 100),
        ("RationalFourierLayer [4,32,32^3] modes 8^3 (cfg-1 shape)", (4, 32, 32, 32, 32),
         lambda: conditioned(EagerRationalLayer(32, (8, 8, 8))), lambda: conditioned(pb.RationalFourierLayer(32, 32, (8, 8, 8))), "auto", 20, 100),
        ("SpectralConv3D [8,64,128^3] modes 32^3 (cfg-2 shape)", (8, 64, 128, 128, 128), lambda: EagerSpectralConv(64, 64, (32, 32, 32)),
         lambda: pb.SpectralConv3D(64, 64, (32, 32, 32)), "tf32", 3, 10),
        ("SpectralConv3D [64,64,256,256,1] modes (64,64,1) (cfg 3)", (64, 64, 256, 256, 1), lambda: EagerSpectralConv(64, 64, (64, 64, 1)),
         lambda: pb.SpectralConv3D(64, 64, (64, 64, 1)), "tf32", 5, 20),
    ]
    if big:     # the 64^3-mode operator of configs 4 and 5 (R(k) alone is 2.2 GB per polynomial term on the incumbent)
        cases = [
            ("RationalFourierLayer [1,64,128^3] modes 64^3 (cfg 4 small-scale operator)", (1, 64, 128, 128, 128),
             lambda: conditioned(EagerRationalLayer(64, (64, 64, 64))), lambda: conditioned(pb.RationalFourierLayer(64, 64, (64, 64, 64))), "tf32", 2, 5),
            ("RationalFourierLayer [1,64,256^3] modes 64^3 (cfg 5)", (1, 64, 256, 256, 256),
             lambda: conditioned(EagerRationalLayer(64, (64, 64, 64))), lambda: conditioned(pb.RationalFourierLayer(64, 64, (64, 64, 64))), "tf32", 2, 5),
        ]
    rows = {}
    for name, shape, make_ref, make_new, prec, warm, reps in cases:
        torch.manual_seed(7)
        x = torch.randn(*shape, device=dev, requires_grad=True)
        g = torch.randn(*shape, device=dev)
        row = {}
        for who, make in (("incumbent_torch_cufft_us", make_ref), (f"pde_fluid_phi_b200[{prec}]_us", make_new)):
            layer = make().to(dev)
            if hasattr(layer, "precision"):
                layer.precision = prec

            def run():
                layer(x).backward(g)
                x.grad = None
                for p in layer.parameters():
                    p.grad = None
            row[who] = 1e3 * statistics.median(event_times(run, warm, reps))
            del layer
        row["speedup"] = row["incumbent_torch_cufft_us"] / row[f"pde_fluid_phi_b200[{prec}]_us"]
        rows[name] = row
        del x, g
        torch.cuda.empty_cache()
    res["layer_fwd_bwd_us"] = rows


def diagnose(res, x1, W, dev, pb):
    """One sample through the incumbent in fp32 (strict) and fp64 and through the library ('auto'), module outputs compared."""
    torch.manual_seed(1234)
    ref = EagerOperator(CFG["modes"], W, CFG["n_layers"], CFG["rational_order"]).to(dev)
    condition(ref.rational_layers)
    new = pb.RationalFourierOperator3D(modes=CFG["modes"], width=W, n_layers=CFG["n_layers"],
                                       rational_order=CFG["rational_order"]).to(dev)
    new.load_state_dict({k: v for k, v in ref.state_dict().items() if k in new.state_dict()}, strict=False)
    new.set_precision("auto")

    def trace(model, x):
        seen = {}
        hooks = [m.register_forward_hook(lambda mod, i, o, name=name: seen.__setitem__(name, o.detach()))
                 for name, m in model.named_modules() if name and "." in name or name in ("input_proj", "output_proj")]
        with torch.no_grad():
            seen["out"] = model(x)
        for h in hooks:
            h.remove()
        return seen

    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        t32 = trace(ref, x1)
        tnew = trace(new, x1)
    t64 = trace(ref.double(), x1.double())
    rows = {}
    for k, v in t64.items():
        if k in t32:
            rows[k] = {"rms_fp64": float(v.pow(2).mean().sqrt()), "incumbent_fp32_vs_fp64": rel(t32[k], v),
                       "mean_offset_over_rms": float((t32[k].double() - v).mean() / v.pow(2).mean().sqrt())}
            if k in tnew and tnew[k].shape == v.shape:
                rows[k]["library_auto_vs_fp64"] = rel(tnew[k], v)
    res["diagnose"] = rows


def layer_times(res, B, W, n, dev, pb):
    # ---- (1) one layer, fwd+bwd ----------------------------------------------------------------------------------
    torch.manual_seed(1234)
    theirs = EagerRationalLayer(W, CFG["modes"], CFG["rational_order"]).to(dev)
    ours = pb.RationalFourierLayer(W, W, CFG["modes"], rational_order=CFG["rational_order"]).to(dev)
    with torch.no_grad():
        ours.P_coeffs.copy_(theirs.P_coeffs)
        ours.Q_coeffs.copy_(theirs.Q_coeffs)
    condition([theirs, ours])
    h = torch.randn(B, W, n, n, n, device=dev, requires_grad=True)
    g = torch.randn(B, W, n, n, n, device=dev)

    def layer_step(layer):
        def run():
            layer(h).backward(g)
            h.grad = layer.P_coeffs.grad = layer.Q_coeffs.grad = None
        return run

    res["layer_fwd_bwd_ms"] = {"incumbent_torch_cufft": statistics.median(event_times(layer_step(theirs), 3, 10))}
    for prec in ("tf32", "auto"):
        ours.precision = prec
        res["layer_fwd_bwd_ms"][f"pde_fluid_phi_b200[{prec}]"] = statistics.median(event_times(layer_step(ours), 5, 20))
    res["layer_peak_mem_gb_incumbent"] = None
    torch.cuda.reset_peak_memory_stats()
    layer_step(theirs)()
    res["layer_peak_mem_gb_incumbent"] = torch.cuda.max_memory_allocated() / 1e9
    torch.cuda.reset_peak_memory_stats()
    layer_step(ours)()
    res["layer_peak_mem_gb_ours"] = torch.cuda.max_memory_allocated() / 1e9
    del theirs, ours, h, g
    torch.cuda.empty_cache()


def whole_model(res, args, x, y, B, W, dev, pb, flush):
    # ---- (2), (3) whole model: parity on one batch, then the training step --------------------------------------------
    torch.manual_seed(1234)
    ref = EagerOperator(CFG["modes"], W, CFG["n_layers"], CFG["rational_order"]).to(dev)
    new = pb.RationalFourierOperator3D(modes=CFG["modes"], width=W, n_layers=CFG["n_layers"],
                                       rational_order=CFG["rational_order"]).to(dev)
    new.load_state_dict({k: v for k, v in ref.state_dict().items() if k in new.state_dict()}, strict=False)
    with torch.no_grad():
        for a, b in zip(ref.rational_layers, new.rational_layers):
            b.P_coeffs.copy_(a.P_coeffs)
            b.Q_coeffs.copy_(a.Q_coeffs)
    condition(ref.rational_layers)
    condition(new.rational_layers)
    for a, b in zip(ref.parameters(), new.parameters()):
        assert a.shape == b.shape and torch.equal(a, b), "weight copy failed"

    def grads(model):
        model.zero_grad(set_to_none=True)
        out = model(x)
        F.mse_loss(out, y).backward()
        return out.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}

    try:
        # Four runs of the incumbent and two of the library on the same batch and weights.  The incumbent's own 1x1x1
        # convolution runs in TF32 under the stock cudnn.allow_tf32 = True, so its strict-fp32 run is the comparison
        # target, its stock run gives the TF32 noise floor of the incumbent itself, and its fp64 run (half the batch at a
        # time) the fp32 noise floor (new-vs-ref32, ref32-vs-ref64, new-vs-ref64).
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            o_ref, g_ref = grads(ref)
        o_stock, g_stock = grads(ref)
        ref.zero_grad(set_to_none=True)

        def versus(o, g, o0, g0):
            return {"output": rel(o, o0), "grad_max_over_params": max(rel(g[k], g0[k]) for k in g0),
                    "grad_P_coeffs_layer0": rel(g["rational_layers.0.P_coeffs"], g0["rational_layers.0.P_coeffs"]),
                    "grad_Q_coeffs_layer0": rel(g["rational_layers.0.Q_coeffs"], g0["rational_layers.0.Q_coeffs"])}

        parity = {"incumbent_stock_tf32_conv_vs_incumbent_strict_fp32": versus(o_stock, g_stock, o_ref, g_ref)}
        del o_stock, g_stock
        outs = {}
        for prec in ("auto", "tf32"):
            new.set_precision(prec)
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=(prec == "tf32")):   # 'auto' is the fp32-grade route
                outs[prec] = grads(new)
            parity[f"{prec}_vs_incumbent_strict_fp32"] = versus(*outs[prec], o_ref, g_ref)
        new.zero_grad(set_to_none=True)
        if args.fp64:
            ref64 = EagerOperator(CFG["modes"], W, CFG["n_layers"], CFG["rational_order"]).to(dev)
            ref64.load_state_dict(ref.state_dict())
            ref64 = ref64.double()
            ref64.zero_grad(set_to_none=True)
            o64 = []
            for i in range(B):                                   # one sample at a time: the loss is a mean over the batch
                out = ref64(x[i:i + 1].double())
                (F.mse_loss(out, y[i:i + 1].double()) / B).backward()
                o64.append(out.detach())
            o64 = torch.cat(o64)
            g64 = {k: p.grad.detach() for k, p in ref64.named_parameters()}
            parity["incumbent_strict_fp32_vs_incumbent_fp64"] = versus(o_ref, g_ref, o64, g64)
            for prec in outs:
                parity[f"{prec}_vs_incumbent_fp64"] = versus(*outs[prec], o64, g64)
            del ref64, o64, g64
        res["whole_model_parity"] = parity
        del o_ref, g_ref, outs
    except torch.OutOfMemoryError as e:          # the incumbent keeps ~15 GB per layer at B = 8
        res["whole_model_parity"] = f"out of memory: {str(e)[:120]}"
    torch.cuda.empty_cache()
    flush()
    if args.parity_only:
        return

    def train_step(model):
        params = list(model.parameters())
        opt = torch.optim.AdamW(params, lr=1e-8, weight_decay=1e-4)     # bench.py's first warm-up value

        def run():
            opt.zero_grad(set_to_none=True)
            loss = F.mse_loss(model(x), y)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
        return run

    step = {}
    try:
        torch.cuda.reset_peak_memory_stats()
        ms = statistics.median(event_times(train_step(ref), args.warmup, args.steps))
        step["incumbent_torch_cufft"] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3,
                                         "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    except torch.OutOfMemoryError as e:
        step["incumbent_torch_cufft"] = f"out of memory: {str(e)[:120]}"
    del ref
    torch.cuda.empty_cache()
    for prec in ("tf32", "auto"):
        new.set_precision(prec)
        torch.cuda.reset_peak_memory_stats()
        ms = statistics.median(event_times(train_step(new), args.warmup, args.steps))
        step[f"pde_fluid_phi_b200[{prec}]"] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3,
                                                "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    res["train_step"] = step
    flush()


if __name__ == "__main__":
    main()
