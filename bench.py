#!/usr/bin/env python
"""bench.py -- RFNO-3D 128^3 training-step throughput (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--precision fp32|tf32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one full training step (forward, MSE, backward, gradient all-reduce for N>1, AdamW)
of RationalFourierOperator3D(modes 32^3, width 64, rational order (4,4), 4 layers) on a batch of 8
synthetic random-spectrum turbulence fields per GPU at 128^3 (BASELINE.json configs[1]; weak scaling).
Rank 0 prints ONE JSON line.  --impl reference times the oracle's CPU restatement of the
reference's PyTorch path on the host cores (the reference is pure Python and is not on the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(modes=(32, 32, 32), width=64, rational_order=(4, 4), n_layers=4, grid=128, batch_per_gpu=8)
METRIC = "RFNO-3D 128^3 train samples/s"


# ------------------------------------------------------------------------------------------------
# synthetic data: random-spectrum turbulence (data/turbulence_dataset.py:155-194 of the reference)
# ------------------------------------------------------------------------------------------------
def turbulence_batch(batch: int, n: int, seed: int, device, kmax: float = 32.0) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    k1 = torch.fft.fftfreq(n, 1.0 / n)
    kz1 = torch.fft.rfftfreq(n, 1.0 / n)
    kx, ky, kz = torch.meshgrid(k1, k1, kz1, indexing="ij")
    kk = torch.stack([kx, ky, kz]).to(device)
    kmag = torch.sqrt((kk ** 2).sum(0))
    amp = torch.where((kmag >= 1) & (kmag <= kmax), kmag.clamp(min=1.0) ** (-5.0 / 6.0), torch.zeros_like(kmag))
    dealias = ((kx.abs() < n / 3) & (ky.abs() < n / 3) & (kz.abs() < n / 3)).to(device)
    out = []
    for _ in range(batch):
        noise = torch.randn(2, 3, n, n, n // 2 + 1, generator=g).to(device)
        u_hat = torch.complex(noise[0], noise[1]) * amp
        div = (kk * u_hat).sum(0) / (kmag ** 2).clamp(min=1.0)
        u_hat = (u_hat - kk * div) * dealias                      # Leray projection + 2/3 mask
        u_hat[:, 0, 0, 0] = 0                                     # zero mean flow
        u = torch.fft.irfftn(u_hat, s=(n, n, n), dim=[-3, -2, -1])
        out.append(u / u.std().clamp(min=1e-12))
    return torch.stack(out).float()


# ------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [s.strip() for s in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm_load = sorted(sm)[len(sm) // 4:] if sm else []          # drop the idle quarter
        return {"sm_mhz": statistics.median(sm_load) if sm_load else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference's PyTorch path, bounded sample
# ------------------------------------------------------------------------------------------------
CPU_OUT_SLICE = 8      # output channels of one layer evaluated per CPU step (of 64)


def cpu_layer_slice_seconds(reps: int, warmup: int):
    """One RationalFourierLayer fwd+bwd on ONE sample [1,64,128,128,128], restricted to CPU_OUT_SLICE of
    the 64 output channels (the cost is dominated by the O*I*M polynomial evaluation and its backward,
    which is linear in the number of output channels).  Returns per-rep seconds."""
    from oracle import spectral_oracle as so
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(1234)
    C, modes, n = CFG["width"], CFG["modes"], CFG["grid"]
    P, Q = so.init_rational_coeffs(CPU_OUT_SLICE, C, CFG["rational_order"])
    P.requires_grad_(True)
    Q.requires_grad_(True)
    x = turbulence_batch(1, n, 42, "cpu")[:, :1].expand(1, C, n, n, n).contiguous().requires_grad_(True)
    mx, my, mz = modes
    mzh = mz // 2 + 1
    mask = so.decay_mask(modes)

    def step():
        # same op sequence as the reference layer (rational_fourier.py:150-172) for a slice of out channels
        x_ft = torch.fft.rfftn(x, dim=[-3, -2, -1])
        R, _, _ = so.rational_transfer(P, Q, modes, 1e-6)
        corner = x_ft[:, :, :mx, :my, :mzh]
        mixed = torch.einsum("bixyz,oixyz->boxyz", corner, R.to(corner.dtype))
        damped = mixed * mask
        e0 = torch.sum(torch.abs(mixed) ** 2, dim=(-3, -2, -1), keepdim=True)
        e1 = torch.sum(torch.abs(damped) ** 2, dim=(-3, -2, -1), keepdim=True)
        damped = damped * torch.sqrt((e0 + 1e-6) / (e1 + 1e-6))
        out_ft = torch.zeros(1, CPU_OUT_SLICE, n, n, n // 2 + 1, dtype=x_ft.dtype)
        out_ft[:, :, :mx, :my, :mzh] = damped
        out = torch.fft.irfftn(out_ft, s=(n, n, n), dim=[-3, -2, -1])
        out.square().mean().backward()
        x.grad = P.grad = Q.grad = None

    times = []
    for i in range(warmup + reps):
        t0 = time.perf_counter()
        step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def cpu_samples_per_s(slice_seconds: float) -> float:
    """Extrapolate the bounded sample to the full training step of one sample: 64/CPU_OUT_SLICE slices per
    layer x 4 layers (pointwise conv / GELU / optimizer excluded, so this flatters the CPU)."""
    per_sample = slice_seconds * (CFG["width"] / CPU_OUT_SLICE) * CFG["n_layers"]
    return 1.0 / per_sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times = cpu_layer_slice_seconds(args.steps, args.warmup)
    per = sum(times) / len(times)
    val = cpu_samples_per_s(per)
    sample = (f"{CPU_OUT_SLICE}/64 output channels of one RationalFourierLayer fwd+bwd on one [1,64,128,128,128] "
              f"sample per step, scaled x{CFG['width'] // CPU_OUT_SLICE} x{CFG['n_layers']} layers; pointwise ops excluded")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, "fp32"),
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int, precision: str) -> dict:
    return {"workload": "RationalFourierOperator3D modes=(32,32,32) width=64 rational_order=(4,4) n_layers=4, "
                        "128^3 synthetic random-spectrum turbulence, batch 8 per GPU, fwd+MSE+bwd+clip_grad_norm(1.0)+AdamW(lr 1e-3 with the trainer's linear warm-up over 1e5 steps -- longer for runs of more than ~15 steps, see bench.py --, wd 1e-4), AMP off",
            "global_batch": CFG["batch_per_gpu"] * n_gpus, "grid": [CFG["grid"]] * 3, "parallelism": f"dp{n_gpus}",
            "transform_precision": precision,
            "init": "reference init (seed 1234) with non-constant Q_coeffs x1e-4 and P_coeffs x1e-4 (activations O(1)); the default init has poles of R(k) inside the retained box and overflows in 4 layers, in the reference too",
            "l2": "per-step activations (4.3 GB per tensor) exceed the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------
# measured DRAM traffic per launch, read from the ncu summaries committed under profiles/ (never a literal)
# ------------------------------------------------------------------------------------------------
# bench kernel name -> prefix of the kernel name in the profiles/*ncu_full*.csv summaries (scripts/ncu_summary.py)
NCU_NAME = {"analysis_zy_tc_kernel": "tc::analysis_zy_tc_kernel", "synthesis_zy_tc_kernel": "tc::synthesis_zy_tc_kernel",
            "analysis_x_tc_kernel": "xtc::analysis_x_tc_kernel", "synthesis_x_tc_kernel": "xtc::synthesis_x_tc_kernel",
            "block_tail_fwd_kernel": "blk::block_tail_fwd_kernel<0", "block_spectral_fwd_kernel": "blk::block_tail_fwd_kernel<1",
            "block_tail_bwd_kernel": "blk::block_tail_bwd_kernel", "block_spectral_bwd_kernel": "blk::block_tail_bwd_kernel",
            "gen_axis_analysis_kernel": "gen::gen_plane_analysis_kernel", "rational_mix_kernel": "rational_mix_kernel",
            "mix_from_R_kernel": "mix_from_R_kernel", "rational_g_kernel": "rational_g_kernel",
            "rational_dpq_kernel": "rational_dpq", "projection_stats_kernel": "projection_stats_kernel",
            "projection_bwd_kernel": "projection_bwd_kernel", "gen_axis_synthesis_kernel": "gen::gen_axis_synthesis_kernel",
            "analysis_zy_tc3_kernel": "tc3::analysis_zy_tc3_kernel", "synthesis_zy_tc3_kernel": "tc3::synthesis_zy_tc3_kernel",
            "pointwise_wgrad_kernel": "pointwise_wgrad_kernel"}


def ncu_traffic_from_profiles():
    """{bench kernel name: (bytes per launch, source file)}: dram__bytes_read.sum + dram__bytes_write.sum of the launch
    with the most traffic of each kernel (the full-size variant) in the newest profiles/*ncu_full*.csv that has it."""
    import csv
    import glob
    out = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_full*.csv"))):      # later rounds override earlier ones
        try:
            rows = list(csv.reader(open(path)))
        except OSError:
            continue
        if not rows:
            continue
        hdr = rows[0]
        try:
            ir = next(i for i, h in enumerate(hdr) if h.startswith("dram__bytes_read.sum"))
            iw = next(i for i, h in enumerate(hdr) if h.startswith("dram__bytes_write.sum"))
        except StopIteration:
            continue
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        ur = scale.get(hdr[ir].split("[")[-1].rstrip("]"), 1e9)
        uw = scale.get(hdr[iw].split("[")[-1].rstrip("]"), 1e9)
        best = {}
        for r in rows[1:]:
            if len(r) <= max(ir, iw):
                continue
            name = r[0].replace("pdephi::", "")
            for bench_name, prefix in NCU_NAME.items():
                if bench_name.startswith("block_") and "bwd" in bench_name:
                    # block_tail_bwd_kernel<GENG, WT, GENH, ZA[, PACKDU]>: the ZA instantiations are the fused backward
                    # (block_spectral_bwd_kernel in the library's profile), the others (or the 3-parameter name of the
                    # round-1 captures) the plain block-tail backward
                    args = name[name.find("<") + 1:name.rfind(">")].replace(" ", "").split(",") if "<" in name else []
                    za = len(args) >= 4 and args[3] in ("1", "true")       # (a fifth argument, PACKDU, since the packed derivative)
                    if za != (bench_name == "block_spectral_bwd_kernel"):
                        continue
                if name.startswith(prefix) or name.startswith(prefix.split("::")[-1]):
                    try:
                        b = float(r[ir]) * ur + float(r[iw]) * uw
                    except ValueError:
                        continue
                    if b > best.get(bench_name, 0.0):
                        best[bench_name] = b
        for k, b in best.items():
            out[k] = (b, os.path.relpath(path, ROOT))
    return out


# ------------------------------------------------------------------------------------------------
# the incumbent on the same GPU: the reference's op sequence in stock PyTorch (cuFFT, cuBLAS, eager elementwise ops)
# ------------------------------------------------------------------------------------------------
def gpu_incumbent(dev, x, y, steps: int = 3):
    """Training step of scripts/incumbent_torch_cufft.py's EagerOperator (the reference restated with plain torch ops in the
    reference's op order; bit-identical to the reference's fixture on the CPU, tests/test_incumbent_cpu.py) on the same
    batch: SURVEY.md section 8d's "number to beat".  One warm-up + `steps` timed steps, CUDA events."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    try:
        import incumbent_torch_cufft as inc
    except Exception as e:       # the script is measurement support, not product: report instead of failing the bench
        return {"unavailable": f"{type(e).__name__}: {e}"}
    try:
        torch.manual_seed(1234)
        ref = inc.EagerOperator(CFG["modes"], CFG["width"], CFG["n_layers"], CFG["rational_order"]).to(dev)
        inc.condition(ref.rational_layers)
        params = list(ref.parameters())
        opt = torch.optim.AdamW(params, lr=1e-8, weight_decay=1e-4)

        def run():
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.mse_loss(ref(x), y)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
        torch.cuda.reset_peak_memory_stats()
        ms = inc.event_times(run, 1, steps)
        med = statistics.median(ms)
        out = {"ms_per_step": round(med, 2), "samples_s": round(x.shape[0] / med * 1e3, 3), "steps": steps,
               "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 1),
               "what": "stock PyTorch eager restatement of the reference operator (torch.fft -> cuFFT, einsum, cuDNN 1x1x1 conv, "
                       "eager elementwise), same batch / weights / optimizer step, same GPU (scripts/incumbent_torch_cufft.py)"}
    except torch.OutOfMemoryError as e:
        out = {"unavailable": f"out of memory: {str(e)[:120]}"}
    finally:
        ref = opt = params = None
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
# N > 1 only: the two multi-GPU parts of the north star that the weak-scaling line does not exercise
# ------------------------------------------------------------------------------------------------
def slab_block(dev, rank: int, world: int, precision: str, steps: int = 5):
    """cfg 5: one RationalFourierLayer [1,64,256^3], modes 64^3, fwd+bwd on x slabs with the NCCL all-to-all transpose
    (strong scaling: one 256^3 sample over all ranks).  Device time, max over ranks; parity of the slab result against the
    single-GPU layer on rank 0's slab (outputs, input gradient) and on the summed coefficient gradients -- asserted."""
    import torch.distributed as dist
    import pde_fluid_phi_b200 as pb
    grid, modes, C = (256, 256, 256), (64, 64, 64), 64
    torch.manual_seed(3)
    layer = pb.RationalFourierLayer(C, C, modes).to(dev)
    with torch.no_grad():
        q0 = layer.Q_coeffs[..., 0, 0, 0].clone()
        layer.Q_coeffs.mul_(1e-6)
        layer.Q_coeffs[..., 0, 0, 0] = q0
    layer.precision = precision
    nxl = grid[0] // world
    g = torch.Generator(device=dev).manual_seed(100)
    xfull = torch.randn(1, C, *grid, device=dev, generator=g)
    gfull = torch.randn(1, C, *grid, device=dev, generator=g)
    x = xfull[:, :, rank * nxl:(rank + 1) * nxl].contiguous().requires_grad_(True)
    go = gfull[:, :, rank * nxl:(rank + 1) * nxl].contiguous()

    def timed(fn, n):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    res = {"workload": "RationalFourierLayer [1,64,256^3] modes 64^3 fwd+bwd, x slabs of 256/N planes, strong scaling",
           "transform_precision": precision, "steps": steps}
    outs = {}
    for exchange in ("alltoall", "allreduce"):
        layer.slab = (None, grid[0], exchange)

        def step():
            x.grad = None
            layer.zero_grad(set_to_none=True)
            out = layer(x)
            out.backward(go)
            return out
        for _ in range(2):
            out = step()
        gP, gQ = layer.P_coeffs.grad.clone(), layer.Q_coeffs.grad.clone()
        dist.all_reduce(gP)     # "alltoall": per-ky-subset partial sums; "allreduce": complete but pre-divided by world
        dist.all_reduce(gQ)
        outs[exchange] = (out.detach().clone(), x.grad.clone(), gP, gQ)
        # a 25-launch, 2-7 ms step is exposed to host jitter (eight ranks share the box's cores): three repetitions of the
        # `steps`-step timing, the median reported and all three kept
        reps = sorted(timed(step, steps) for _ in range(3))
        res["ms" if exchange == "alltoall" else "ms_allreduce_exchange"] = round(reps[1], 3)
        res["ms_reps" if exchange == "alltoall" else "ms_allreduce_exchange_reps"] = [round(v, 3) for v in reps]
        if exchange == "alltoall":
            # where the time goes: the library's kernels (CUDA events around every launch, summed), the collectives of one
            # fwd+bwd run alone on tensors of the same size (4 all-to-all transposes with their local permute copies, the
            # projection's two small all-reduces), and what is left -- launch gaps: at 8 GPUs a kernel lasts 20-60 us,
            # about what the host needs to enqueue the next one
            from pde_fluid_phi_b200 import _cabi, functional as pf
            _cabi.profile(True, True)
            for _ in range(steps):
                step()
            torch.cuda.synchronize()
            rep = _cabi.profile_report(timed=True)
            _cabi.profile(False, True)
            kern = torch.tensor([sum(v[1] for v in rep.values()) / steps], device=dev)
            dist.all_reduce(kern, op=dist.ReduceOp.MAX)
            T = torch.zeros(1, C, nxl, modes[1], modes[2] // 2 + 1, dtype=torch.complex64, device=dev)
            small = torch.zeros(2 * C, device=dev)

            def comm():
                for _ in range(2):
                    pf.exchange_ky_to_x(pf.exchange_x_to_ky(T, dist.group.WORLD, world), dist.group.WORLD, world)
                dist.all_reduce(small)
                dist.all_reduce(small[:C])
            for _ in range(2):
                comm()
            comm_ms = timed(comm, steps)
            res["breakdown_ms"] = {"kernels": round(float(kern), 3), "collectives_alone": round(comm_ms, 3),
                                   "launch_gaps_and_waits": round(res["ms"] - float(kern) - comm_ms, 3),
                                   "kernel_launches_per_step": int(sum(v[0] for v in rep.values()) / steps),
                                   "exchange_MB_per_rank_per_alltoall": round(T.numel() * 8 / 1e6, 1)}
            del T, small
    # the single-GPU layer on the whole volume: every rank runs it (same device time), rank 0's number is reported
    layer.slab = None
    xf = xfull.requires_grad_(True)

    def step1():
        xf.grad = None
        layer.zero_grad(set_to_none=True)
        out = layer(xf)
        out.backward(gfull)
        return out
    for _ in range(2):
        of = step1()
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    sl = slice(rank * nxl, (rank + 1) * nxl)
    a2a = outs["alltoall"]
    errs = torch.tensor([rel(a2a[0], of.detach()[:, :, sl]), rel(a2a[1], xf.grad[:, :, sl]),
                         rel(a2a[2], layer.P_coeffs.grad), rel(a2a[3], layer.Q_coeffs.grad),
                         rel(outs["allreduce"][0], of.detach()[:, :, sl])], device=dev, dtype=torch.float64)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    res["ms_1gpu"] = round(timed(step1, steps), 3)
    res["speedup"] = round(res["ms_1gpu"] / res["ms"], 3)
    res["efficiency"] = round(res["ms_1gpu"] / res["ms"] / world, 4)
    res["relL2_vs_single_gpu_out"], res["relL2_dx"], res["relL2_dP"], res["relL2_dQ"], res["relL2_allreduce_exchange_out"] = (
        float(f"{v:.3e}") for v in errs.tolist())
    res["relL2_grads"] = max(res["relL2_dx"], res["relL2_dP"], res["relL2_dQ"])
    # slab and single-GPU runs use the same kernels on differently shaped x stages: fp32-grade agreement on outputs; the
    # TF32 rounding of differently grouped partial sums on the gradients of the coefficient sums
    gate_out, gate_grad = (2e-3, 2e-3) if precision == "tf32" else (1e-5, 5e-5)
    res["gate"] = {"out": gate_out, "grads": gate_grad}
    ok = res["relL2_vs_single_gpu_out"] <= gate_out and res["relL2_grads"] <= gate_grad and res["relL2_allreduce_exchange_out"] <= gate_out
    res["parity_ok"] = bool(ok)      # False: bench.py prints the line and exits with code 3
    del xfull, gfull, x, go, xf, outs, layer
    torch.cuda.empty_cache()
    return res


def cfg3_dp_block(dev, rank: int, world: int, steps: int = 8):
    """cfg 3: 2-D Kolmogorov FNO (modes (64,64,1), width 64, 4 layers, 256^2, batch 64 per GPU): the data-parallel case with
    real gradient volume (537 MB of complex64 weight gradients per step) through distributed.OverlappedGradientAllReduce.
    Training step with AdamW, device time, max over ranks; efficiency against the same step without any reduction."""
    import math
    import torch.distributed as dist
    import pde_fluid_phi_b200 as pb
    from pde_fluid_phi_b200.distributed import OverlappedGradientAllReduce, allreduce_gradients, broadcast_parameters
    n, B = 256, 64
    torch.manual_seed(1234)
    model = pb.FNO3D(modes=(64, 64, 1), width=64, n_layers=4, in_channels=2, out_channels=2).to(dev).set_precision("tf32")
    broadcast_parameters(model)
    params = list(model.parameters())
    opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=1e-4)
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    yy = torch.linspace(0, 2 * math.pi, n, device=dev)
    base = torch.sin(4 * yy)[None, None, None, :, None].expand(B, 1, n, n, 1)
    x = torch.cat([base + 0.1 * torch.randn(B, 1, n, n, 1, device=dev, generator=g),
                   0.1 * torch.randn(B, 1, n, n, 1, device=dev, generator=g)], 1)
    y = torch.roll(x, 3, dims=3)
    reducer = OverlappedGradientAllReduce(params, world)

    def step(mode):
        opt.zero_grad(set_to_none=True)
        reducer.enabled = mode == "overlapped"
        loss = torch.nn.functional.mse_loss(model(x), y)
        loss.backward()
        if mode == "flat":
            allreduce_gradients(params, world)
        elif mode == "overlapped":
            reducer.finish()
        opt.step()
        return loss

    res = {"workload": "FNO3D modes (64,64,1) width 64, 4 layers, [64,2,256,256,1] per GPU, fwd+MSE+bwd+AdamW, tf32 route",
           "grad_bytes": sum(p.numel() * p.element_size() for p in params), "steps": steps}
    for mode in ("none", "flat", "overlapped"):
        for _ in range(3):
            step(mode)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step(mode)
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[f"ms_per_step_{mode}"] = round(float(t), 3)
    # both reductions give the same gradients from the same weights (element for element)
    grads = {}
    for mode in ("flat", "overlapped"):
        opt.zero_grad(set_to_none=True)
        reducer.enabled = mode == "overlapped"
        torch.nn.functional.mse_loss(model(x), y).backward()
        allreduce_gradients(params, world) if mode == "flat" else reducer.finish()
        grads[mode] = [p.grad.clone() for p in params]
    # Two ranks: one addition per element, the two reductions agree bit for bit.  More ranks: NCCL picks its algorithm (ring /
    # tree / NVLS, i.e. the order of the additions) by message size, and the flat buffer and the per-tensor in-place
    # reductions have different sizes -- same sums up to fp32 rounding, which is what is checked.
    same = all(torch.equal(a, b) for a, b in zip(grads["flat"], grads["overlapped"]))
    def _rel(a, b):
        a, b = torch.view_as_real(a) if a.is_complex() else a, torch.view_as_real(b) if b.is_complex() else b
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))
    worst = max(_rel(a, b) for a, b in zip(grads["overlapped"], grads["flat"]))
    res["overlapped_equals_flat"] = bool(same)
    res["overlapped_vs_flat_max_relL2"] = float(f"{worst:.3e}")
    same = same or worst < 1e-6
    res["samples_s"] = round(B * world / res["ms_per_step_overlapped"] * 1e3, 1)
    res["samples_s_flat"] = round(B * world / res["ms_per_step_flat"] * 1e3, 1)
    res["efficiency"] = round(res["ms_per_step_none"] / res["ms_per_step_overlapped"], 4)
    res["loss_finite"] = bool(torch.isfinite(loss))
    reducer.remove()
    del model, opt, x, y, grads
    torch.cuda.empty_cache()
    res["checks_ok"] = bool(same and res["loss_finite"])      # False: bench.py prints the line and exits with code 3
    return res


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def run_native(args):
    import torch.distributed as dist
    import pde_fluid_phi_b200 as pb
    from pde_fluid_phi_b200 import _cabi
    from pde_fluid_phi_b200.distributed import allreduce_gradients

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched through torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"     # keep stdout to the one JSON line (the image sets NCCL_DEBUG=VERSION)
        dist.init_process_group("nccl", device_id=dev)
    pb.load_library()

    n, B = CFG["grid"], args.batch or CFG["batch_per_gpu"]
    torch.manual_seed(1234)
    model = pb.RationalFourierOperator3D(modes=CFG["modes"], width=CFG["width"], n_layers=CFG["n_layers"],
                                         rational_order=CFG["rational_order"]).to(dev)
    model.set_precision(args.precision)
    # SURVEY.md section 8c weight set (ii): the reference's default init puts poles of R(k) inside the retained
    # box (|R| up to 1e6) and the 4-layer model overflows to NaN on any input, in the reference too.  Shrinking
    # the non-constant Q coefficients keeps the loss finite; the arithmetic per step is identical.
    with torch.no_grad():
        for layer in model.rational_layers:
            q0 = layer.Q_coeffs[..., 0, 0, 0].clone()
            layer.Q_coeffs.mul_(1e-4)
            layer.Q_coeffs[..., 0, 0, 0] = q0
            layer.P_coeffs.mul_(1e-4)
    # Warm-up horizon: 1e5 steps for the default run; stretched for longer runs so that the cumulative AdamW drift of the
    # cubic P / Q coefficients, ~ sum_i lr_i = n^2 / (2 horizon) * 1e-3, stays ~3e-8.  The reference parameterisation is
    # that fragile: a coefficient drift of 1e-6 changes R(k) by 1e-6 k^3 = 3 % at k = 31, and
    # a 60-step run with the fixed 1e5 horizon walked a pole of R(k) back into the retained box (loss 2.7 -> 7e6; on the
    # TF32 route already at a drift of 2e-6, the fp32-grade route survived that one).  Throughput is value-independent;
    # the stretched schedule only keeps the numbers finite for any --steps.
    n_total = 2 * (max(args.warmup, 3) + 2 * args.steps + 8)
    horizon = 1e5 if n_total <= 42 else 2000.0 * n_total * n_total
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3 / horizon, weight_decay=1e-4)   # warm-up value of step 0
    x = turbulence_batch(B, n, 42 + rank, dev)
    y = turbulence_batch(B, n, 1042 + rank, dev)
    hx, hy = x.cpu().pin_memory(), y.cpu().pin_memory()
    h_loss = torch.empty((), dtype=torch.float32).pin_memory()

    # The reference's train step (training/stability_trainer.py:170-210): zero_grad, loss, backward, clip_grad_norm_,
    # AdamW step, linear LR warm-up (warmup_epochs=10 of an assumed 10k-batch epoch -> 1e5 steps: AdamW's sign-like
    # first steps move every Q coefficient by ~lr, and 1e-6 * k^3 (k up to 31) is already enough to put a pole of
    # R(k) inside the retained box, so the reference parameterisation only trains with a long warm-up).
    state = {"n": 0}
    params = [p for p in model.parameters()]

    def step(xd, yd, before_loss=None):
        opt.zero_grad(set_to_none=True)
        pred = model(xd)
        if before_loss is not None:
            before_loss()                                    # e2e leg: the target's copy has to have landed by now
        loss = torch.nn.functional.mse_loss(pred, yd)
        loss.backward()
        if world > 1:
            allreduce_gradients(params, world)
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        state["n"] += 1
        for gparam in opt.param_groups:
            gparam["lr"] = min(1.0, (state["n"] + 1) / horizon) * 1e-3
        return loss

    # End-to-end leg: every step copies its own inputs from pinned host memory (2 x 201 MB) and reads the loss back.
    # The copies run on a side stream into the buffer pair the *next* step will use, as a prefetching data loader
    # does, so they overlap the current step's kernels; the first step's input copy is not overlapped (its target
    # copy runs behind the forward pass, which only needs the inputs) and is inside the timed region like all the others.
    copy_stream = torch.cuda.Stream(device=dev)
    dbuf = [(torch.empty_like(x), torch.empty_like(y)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]           # inputs of the slot have landed
    ready_y = [torch.cuda.Event() for _ in range(2)]         # ... and its targets (needed only at the loss)
    consumed = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"i": 0, "primed": False}

    def _prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])           # the step that last read this pair has finished with it
            dbuf[slot][0].copy_(hx, non_blocking=True)
            ready[slot].record(copy_stream)
            dbuf[slot][1].copy_(hy, non_blocking=True)       # overlaps the forward pass even when nothing was prefetched
            ready_y[slot].record(copy_stream)

    def step_e2e():
        i = e2e_state["i"]
        if not e2e_state["primed"]:
            _prefetch(i & 1)
            e2e_state["primed"] = True
        _prefetch((i + 1) & 1)                               # next step's inputs, overlapped with this step
        torch.cuda.current_stream().wait_event(ready[i & 1])
        loss = step(dbuf[i & 1][0], dbuf[i & 1][1],
                    before_loss=lambda: torch.cuda.current_stream().wait_event(ready_y[i & 1]))
        consumed[i & 1].record(torch.cuda.current_stream())
        h_loss.copy_(loss.detach(), non_blocking=True)
        e2e_state["i"] = i + 1
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    first_loss = None
    for _ in range(max(args.warmup, 3)):
        l0 = step(x, y)
        if first_loss is None:
            first_loss = float(l0.detach())
    # ---- timed region 1: inputs resident in HBM ---------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _cabi.profile(enable=True, reset=True)
    total_ms = timed(lambda: step(x, y), args.steps)
    prof = _cabi.profile_report(timed=True)
    _cabi.profile(enable=False, reset=True)
    clocks = sampler.stop() if rank == 0 else {}
    # ---- timed region 2: end to end, host buffers ----------------------------------------------------
    for ev in consumed:
        ev.record(torch.cuda.current_stream())
    step_e2e()                                               # untimed warm-up of the copy path
    torch.cuda.synchronize()
    e2e_state["primed"] = False                              # the timed region starts with nothing prefetched
    e2e_ms = timed(step_e2e, args.steps)
    final_loss = float(h_loss)

    # the other precision route, same step, same number of timed steps, with its own per-kernel profile
    other = None
    other_prof = None
    if world == 1 and not args.no_other_precision:
        alt = "auto" if args.precision == "tf32" else "tf32"    # "auto" = the modules' default: fp32-grade (split-TF32 tcgen05)
        model.set_precision(alt)
        for _ in range(3):
            step(x, y)
        _cabi.profile(enable=True, reset=True)
        alt_total = timed(lambda: step(x, y), args.steps)
        other_prof = _cabi.profile_report(timed=True)
        _cabi.profile(enable=False, reset=True)
        alt_ms = alt_total / args.steps
        model.set_precision(args.precision)
        other = {"transform_precision": alt, "value": round(B / (alt_ms * 1e-3), 3), "unit": "samples/s",
                 "ms_per_step": round(alt_ms, 3), "steps": args.steps, "total_ms": alt_total}
    fused_ends = bool(getattr(model, "fuse_lift", False) and getattr(model, "fuse_head", False))
    # N > 1: the slab-decomposed cfg-5 layer (all-to-all transpose) and cfg 3's data-parallel step with real gradient volume
    slab = cfg3 = None
    if world > 1 and not args.no_multi_gpu_blocks:
        del x, y, opt, model
        torch.cuda.empty_cache()
        x = y = None
    layer_us = None
    incumbent = None
    if rank == 0 and world == 1:
        if not args.no_incumbent:
            del opt, model
            torch.cuda.empty_cache()
            incumbent = gpu_incumbent(dev, x, y)
        if not args.no_layer_times:
            x = y = None
            torch.cuda.empty_cache()
            layer_us = layer_times(dev, args.precision)
    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = B * world / (ms_per_step * 1e-3)
        e2e_value = B * world / (e2e_ms / args.steps * 1e-3)
        launches = sum(v[0] for v in prof.values())
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak, which = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback")
        traffic = ncu_traffic_from_profiles()
        per_kernel, rooflines, dom = kernel_rooflines(prof, total_ms, B, n, args.precision, fused_ends, peak, traffic)
        roof = None
        if dom:
            roof = dict(rooflines[dom], kernel=dom, peak_source=which,
                        note="largest share of the step; 'rooflines' lists every kernel of the path")
        if other is not None:
            o_kernels, o_roofs, o_dom = kernel_rooflines(other_prof, other.pop("total_ms"), B, n, other["transform_precision"],
                                                        fused_ends, peak, traffic)
            other["roofline"] = dict(o_roofs[o_dom], kernel=o_dom) if o_dom else None
            other["kernels"] = o_kernels
        if incumbent is not None and "ms_per_step" in incumbent:
            incumbent["speedup_of_this_library"] = round(incumbent["ms_per_step"] / ms_per_step, 2)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            t = cpu_layer_slice_seconds(reps=2, warmup=1)
            per = sum(t) / len(t)
            cpu = {"value": cpu_samples_per_s(per), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{CPU_OUT_SLICE}/64 output channels of one RationalFourierLayer fwd+bwd on one "
                             f"[1,64,128,128,128] sample ({per:.2f} s), scaled x{CFG['width'] // CPU_OUT_SLICE} x{CFG['n_layers']} "
                             "layers; pointwise ops excluded"}
        k2 = sum(v["ms_total"] for k, v in per_kernel.items()
                 if k in ("rational_mix_kernel", "mix_from_R_kernel", "rational_g_kernel", "rational_dpq_kernel", "dpq_reduce_kernel",
                          "projection_stats_kernel", "projection_bwd_kernel", "rational_transfer_kernel"))
        line = {"metric": METRIC, "value": round(value, 3), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": {"tf32": "tf32", "fp32": "f32"}.get(args.precision, "tf32x3 (fp32-grade)"),
                "data": "synthetic", "config": workload_config(world, args.precision),
                "e2e": {"value": round(e2e_value, 3), "unit": "samples/s",
                        "h2d_bytes_per_step": hx.numel() * 4 + hy.numel() * 4, "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "rooflines": rooflines,
                "cpu_baseline": cpu, "gpu_incumbent": incumbent, "layer_fwd_bwd_us": layer_us, "other_precision": other,
                "k2_share_of_step": round(k2 / total_ms, 4), "kernels": per_kernel, "first_loss": first_loss,
                "final_loss": final_loss,
                # what is NOT fp32 in memory: stated with the line (include/pdephi_b200.h)
                "storage": {"activations_parameters_gradients_optimizer": "fp32",
                            "block_saved_derivative": {0: "pre-activation u, fp32", 1: "gelu'(u), fp32",
                                                       2: "gelu'(u) as binary16 (the 11 significant bits of the TF32 operand it "
                                                          "becomes), one 32-bit word per channel pair and point"}.get(
                                _cabi.load().pdephi_get_option(_cabi.OPT_BLOCK_SAVES_DERIVATIVE), "?")}}
    else:
        line = None
    # N > 1: the slab-decomposed cfg-5 layer (all-to-all transpose) and cfg 3's data-parallel step with real gradient volume.
    # They run AFTER the headline line is complete and under a watchdog: a hang or an exception in them (they are extra
    # measurements with collectives of their own) must not cost the line -- it is then printed without them, with the reason.
    blocks_ok = True
    if world > 1 and not args.no_multi_gpu_blocks:
        import threading

        def give_up():
            if rank == 0:
                line["multi_gpu_blocks"] = f"timed out after {args.block_timeout} s; the line above is complete without them"
                print(json.dumps(line), flush=True)
            os._exit(0)
        dog = threading.Timer(args.block_timeout, give_up)
        dog.daemon = True
        dog.start()
        try:
            slab = slab_block(dev, rank, world, args.precision if args.precision in ("tf32", "fp32") else "tf32")
            blocks_ok = blocks_ok and slab.get("parity_ok", True)
        except BaseException as e:  # noqa: BLE001 -- SystemExit included: reported in the line, never instead of the line
            slab = {"error": f"{type(e).__name__}: {e}"}
            blocks_ok = False
        try:
            cfg3 = cfg3_dp_block(dev, rank, world)
            blocks_ok = blocks_ok and cfg3.get("checks_ok", True)
        except BaseException as e:  # noqa: BLE001
            cfg3 = {"error": f"{type(e).__name__}: {e}"}
            blocks_ok = False
        dog.cancel()
    if rank == 0:
        if slab is not None:
            line["slab"] = slab
        if cfg3 is not None:
            line["cfg3_dp"] = cfg3
        print(json.dumps(line), flush=True)
        out_dir = os.path.join(ROOT, "gpurun_out")
        if world > 1 and os.path.isdir(out_dir):       # kept beside the run for profiles/ (the JSON line carries them too)
            try:
                if slab is not None:
                    json.dump(slab, open(os.path.join(out_dir, f"slab_n{world}.json"), "w"), indent=1)
                if cfg3 is not None:
                    json.dump(cfg3, open(os.path.join(out_dir, f"dp_cfg3_n{world}.json"), "w"), indent=1)
            except OSError:
                pass
    if world > 1:
        dist.destroy_process_group()
    if not blocks_ok:           # a parity failure of the slab route (or an exception): the line is printed, the exit code says so
        sys.exit(3)


def kernel_rooflines(prof, total_ms, B, n, precision, fused_ends, peak, traffic):
    """({kernel: launches / ms / share}, {kernel: roofline dict}, dominant kernel) from the library's per-launch CUDA events.

    Algorithmic bytes per launch follow DESIGN.md section 5 (one launch = all B*C*Nx planes / all B*S/128 tiles / all modes);
    `traffic` = measured dram bytes per launch from the ncu summaries under profiles/ (None where no capture exists)."""
    C, (mx, my, mz), L = CFG["width"], CFG["modes"], CFG["n_layers"]
    mzh = mz // 2 + 1
    M = mx * my * mzh
    act = 4.0 * B * C * n ** 3                       # one full-size fp32 activation
    mid = 8.0 * B * C * n * my * mzh                 # [BC,Nx,my,mzh] complex intermediate
    modes_b = 8.0 * B * C * M                        # retained modes of one tensor [B,C,M] complex64
    tf = 4.0 * C * C * M                             # one materialised transfer-function tensor [O,I,M] fp32 (R, Qe or G)
    order_p, order_q = CFG["rational_order"]
    tri = lambda o: o * (o + 1) * (o + 2) // 6
    fused_ends = fused_ends and precision == "tf32"
    cin = 3
    # The first block takes the input lift (its forward analysis reads the 3-channel input, its block tail generates
    # h0 and, in the backward, stores no t) and the last block the output projection (its backward generates g'):
    # per-launch algorithmic bytes of those kernels are averages over the L layers' launches of one step.
    a_launch = ((2 * L - 1) + cin / C) / (2 * L) if fused_ends else 1.0          # analysis: 2L launches, one on cin channels
    vbytes = 4.0 * B * C * n * n * ((2 * mzh + 3) // 4 * 4)                      # V [B,Nx,Ny,64,ldz]
    fused_bwd = "block_spectral_bwd_kernel" in prof      # backward analysis of gu: z stage inside the block-tail kernel
    if fused_bwd:
        a_launch = 1.0                                   # analysis_zy_tc: the L - 1 full-size forward analyses only
    # the block kernels' `u` buffer: gelu'(pre-activation) as packed binary16 under PDEPHI_OPT_BLOCK_SAVES_DERIVATIVE = 2
    # (the default), half an activation; a full fp32 activation otherwise
    from pde_fluid_phi_b200 import _cabi as _cabi_opt
    ub = act * (0.5 if _cabi_opt.load().pdephi_get_option(_cabi_opt.OPT_BLOCK_SAVES_DERIVATIVE) == 2 else 1.0)
    alg = {
        "analysis_zy_tc_kernel": (act + mid) * a_launch, "analysis_zy_kernel": act + mid,
        # tf32 route with fused ends: the split kernel runs the 3-channel lift analysis only
        "analysis_zy_tc3_kernel": (act + mid) * (cin / C if fused_ends else 1.0),
        # fp32-grade route, unfused block: forward fuses both skip addends, adjoint one; tf32 route: adjoint launches only
        "synthesis_zy_tc3_kernel": act + mid + (act if precision == "tf32" else 1.5 * act),
        "analysis_x_tc3_kernel": mid + modes_b, "synthesis_x_tc3_kernel": mid + modes_b,
        # fused block route: the forward synthesis lives in block_spectral_fwd_kernel, so every launch of this kernel is
        # an adjoint one with one addend (the unfused forward launches, if any, carry none: counted from the launches)
        "synthesis_zy_tc_kernel": act + mid + act,
        "synthesis_zy_kernel": act + mid + 1.5 * act,   # fp32 route: forward fuses both skip addends, adjoint one
        "block_tail_fwd_kernel": 3 * act + ub,       # read h, spec; write u, h'
        # fused forward: read h and V; write u, h' (first block: h generated from the input, one activation less)
        "block_spectral_fwd_kernel": ((2 * act * (L - 1) + act) / L if fused_ends else 2 * act) + ub + vbytes,
        # read g', u, h; write gu, t -- first block: no h read, no t write (2 activations + u); last block: g' generated (3 + u)
        "block_tail_bwd_kernel": ((4 * act * (L - 2) + 2 * act + 3 * act) / L if fused_ends else 4 * act) + ub,
        # fused backward: gu is not stored; the z-analysed gu (Vg, the size of V) is.  Middle: read g', u, h, write t (3 + u);
        # first block: read g', u (1 + u: no h, no t); last block: read u, h, write t (2 + u: g' generated)
        "block_spectral_bwd_kernel": ((3 * act * (L - 2) + act + 2 * act) / L if fused_ends else 3 * act) + ub + vbytes,
        "gen_axis_analysis_kernel": vbytes + mid,    # y stage of the fused backward: Vg in, [BC,Nx,my,mzh] out
        "analysis_x_tc_kernel": mid + modes_b, "synthesis_x_tc_kernel": mid + modes_b,
        # K2.  SURVEY.md section 8d counts 0.145 GB forward / 0.29 GB backward for an evaluation that
        # never materialises R; this design writes R and Q+eps once per forward (2 x tf) and streams them in the backward.
        "rational_mix_kernel": 2 * modes_b + 2 * tf + 4.0 * C * C * (order_p ** 3 + order_q ** 3),
        "rational_transfer_kernel": 2 * tf + 4.0 * C * C * (order_p ** 3 + order_q ** 3),
        "mix_apply_kernel": 2 * modes_b + tf,
        "mix_from_R_kernel": 2 * modes_b + tf,       # dY in, dX out, R in
        "rational_g_kernel": 2 * modes_b + 2 * tf,   # X, dY, Qe in; G out
        "rational_dpq_kernel": 2 * tf,               # G, R in (the monomial tables stay in L2)
        "projection_stats_kernel": modes_b, "projection_bwd_kernel": 3 * modes_b,
        "pointwise_wgrad_kernel": act + 4.0 * B * cin * n ** 3,
        "gen_axis_synthesis_kernel": None,           # x and y stages of the fused forward share the counter: see below
    }
    survey = {"rational_mix_kernel": 2 * modes_b + 4.0 * C * C * (order_p ** 3 + order_q ** 3)}
    # fp32 instruction floor of the bit-exact P/Q evaluation: one rounded product + one rounded add per monomial and
    # polynomial (packed two channels per FFMA2), +eps, an IEEE division (~8 issue slots), 2B FMAs of the mix per (o,i,k)
    lane_ops = C * C * M * ((tri(order_p) + tri(order_q)) + 0.5 + 8 + 2 * B)
    alu_peak = 148 * 128 * 1.965e9                   # fp32 lanes x SMs x max SM clock: lane-instructions per second
    per_kernel = {k: {"launches": v[0], "ms_total": round(v[1], 3), "share_of_step": round(v[1] / total_ms, 4)}
                  for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    rooflines = {}
    for k, v in prof.items():
        if alg.get(k) is None or v[1] <= 0:
            continue
        avg_ms = v[1] / v[0]
        if k.startswith("block_tail_bwd") or k == "block_spectral_bwd_kernel":
            avg_ms = v[1] / (v[0] / 2)           # the tiny partial-sum reduce kernel shares the counter
        if k == "pointwise_wgrad_kernel":
            avg_ms = v[1] / (v[0] / 2)           # ... and so does this one's
            if fused_ends:                       # ... and, with the lift in the first block, so do the (kernel, reduce) launches
                avg_ms = v[1] / (v[0] / 4)       # of its weight gradient on the MODES (71 MB): one activation-sized pass per step
        a = alg[k]
        if k == "synthesis_zy_tc_kernel" and "block_spectral_fwd_kernel" not in prof:
            a = act + mid + 0.5 * act            # unfused forward: half the launches carry no addend
        ach = a / (avg_ms * 1e-3) / 1e9
        r = {"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
             "algorithmic_bytes_per_launch": a, "avg_launch_ms": round(avg_ms, 4)}
        t = traffic.get(k)
        r["traffic"] = t[0] if t else None
        if t:
            r["traffic_source"] = t[1]
        if k in survey:
            r["survey_8d_bytes_per_launch"] = survey[k]
            r["frac_of_survey_8d_bytes"] = round(survey[k] / (avg_ms * 1e-3) / 1e9 / peak, 4)
        if k in ("rational_mix_kernel", "rational_transfer_kernel"):
            ops = lane_ops if k == "rational_mix_kernel" else lane_ops - C * C * M * 2 * B
            r["alu"] = {"bound": "alu", "achieved": round(ops / (avg_ms * 1e-3) / 1e12, 2), "peak": round(alu_peak / 1e12, 2),
                        "unit": "T fp32 lane-instructions/s", "frac": round(ops / (avg_ms * 1e-3) / alu_peak, 4),
                        "note": "instruction-bound by the bit-exact P(k)/Q(k) evaluation (one separately rounded product and add "
                                "per monomial, packed FFMA2), not by HBM"}
        if fused_ends and k in ("block_tail_bwd_kernel", "block_spectral_fwd_kernel", "block_spectral_bwd_kernel"):
            r["traffic_note"] = ("ncu dram bytes of a middle block's launch; algorithmic_bytes_per_launch is the mean "
                                 "over the step's launches (the first / last block's launches move fewer activations)")
        rooflines[k] = r
    big = [k for k in rooflines if rooflines[k]["algorithmic_bytes_per_launch"] >= act]
    dom = max(big, key=lambda k: prof[k][1]) if big else None
    return per_kernel, rooflines, dom


def layer_times(dev, precision):
    """Second half of the BASELINE metric: fwd+bwd microseconds of one spectral layer (CUDA events, 3 warm + 10 timed)."""
    import pde_fluid_phi_b200 as pb

    def run(layer, x):
        g = torch.randn_like(x)

        def step():
            x.grad = None
            layer.zero_grad(set_to_none=True)
            layer(x).backward(g)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            step()
        e1.record()
        torch.cuda.synchronize()
        return round(e0.elapsed_time(e1) / 10 * 1e3, 1)

    def run_graph(layer, x):
        """The same fwd+bwd captured once into a CUDA graph and replayed: what a launch-latency-bound shape costs without
        the Python / ctypes host path (the library never synchronises or allocates outside torch's allocator)."""
        g = torch.randn_like(x)

        def step():
            x.grad = None
            layer.zero_grad(set_to_none=True)
            layer(x).backward(g)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        return round(e0.elapsed_time(e1) / 50 * 1e3, 1)

    out = {}
    torch.manual_seed(0)
    x2 = torch.randn(8, 64, 128, 128, 128, device=dev, requires_grad=True)
    rfl = pb.RationalFourierLayer(64, 64, (32, 32, 32)).to(dev)
    rfl.precision = precision
    out["RationalFourierLayer [8,64,128^3] modes 32^3 (%s)" % precision] = run(rfl, x2)
    sc = pb.SpectralConv3D(64, 64, (32, 32, 32)).to(dev)
    sc.precision = precision
    out["SpectralConv3D [8,64,128^3] modes 32^3 (%s)" % precision] = run(sc, x2)
    del x2, rfl, sc
    x1 = torch.randn(4, 32, 32, 32, 32, device=dev, requires_grad=True)
    out["SpectralConv3D [4,32,32^3] modes 8^3 (fp32, cfg 1)"] = run(pb.SpectralConv3D(32, 32, (8, 8, 8)).to(dev), x1)
    out["RationalFourierLayer [4,32,32^3] modes 8^3 (fp32, cfg-1 shape)"] = run(pb.RationalFourierLayer(32, 32, (8, 8, 8)).to(dev), x1)
    out["SpectralConv3D [4,32,32^3] modes 8^3 (fp32, cfg 1, CUDA graph replay)"] = run_graph(pb.SpectralConv3D(32, 32, (8, 8, 8)).to(dev), x1)
    out["RationalFourierLayer [4,32,32^3] modes 8^3 (fp32, cfg-1 shape, CUDA graph replay)"] = run_graph(
        pb.RationalFourierLayer(32, 32, (8, 8, 8)).to(dev), x1)
    del x1
    if precision == "tf32":
        # shapes of configs 3-5: the per-axis tensor-core route (transforms_gen.cu)
        def conditioned(layer):           # k^3 reaches 2.6e5 at 64 modes: keep Q(k) away from its poles (timing is value-independent)
            with torch.no_grad():
                q0 = layer.Q_coeffs[..., 0, 0, 0].clone()
                layer.Q_coeffs.mul_(1e-6)
                layer.Q_coeffs[..., 0, 0, 0] = q0
            layer.precision = precision
            return layer
        x3 = torch.randn(64, 64, 256, 256, 1, device=dev, requires_grad=True)
        sc3 = pb.SpectralConv3D(64, 64, (64, 64, 1)).to(dev)
        sc3.precision = precision
        out["SpectralConv3D [64,64,256,256,1] modes (64,64,1) (tf32, cfg 3)"] = run(sc3, x3)
        del x3, sc3
        x4 = torch.randn(1, 64, 128, 128, 128, device=dev, requires_grad=True)
        out["RationalFourierLayer [1,64,128^3] modes 64^3 (tf32, cfg 4 small-scale operator)"] = run(
            conditioned(pb.RationalFourierLayer(64, 64, (64, 64, 64)).to(dev)), x4)
        del x4
        x5 = torch.randn(1, 64, 256, 256, 256, device=dev, requires_grad=True)
        out["RationalFourierLayer [1,64,256^3] modes 64^3 (tf32, cfg 5)"] = run(
            conditioned(pb.RationalFourierLayer(64, 64, (64, 64, 64)).to(dev)), x5)
        del x5
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--precision", default=os.environ.get("PDEPHI_PRECISION", "tf32"), choices=["auto", "fp32", "tf32x3", "tf32"],
                    help="transform arithmetic: tf32 = single-pass tcgen05 (north_star gate 2e-3, the headline); auto = fp32-grade "
                         "(gate 1e-5): split-TF32 tcgen05 (tf32x3) on tensor-core shapes, FFMA (fp32) elsewhere")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (debug only; invalidates the metric)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-layer-times", action="store_true")
    ap.add_argument("--no-other-precision", action="store_true")
    ap.add_argument("--no-incumbent", action="store_true", help="skip the stock-PyTorch (cuFFT) training step on the same GPU")
    ap.add_argument("--no-multi-gpu-blocks", action="store_true", help="N > 1: skip the cfg-5 slab layer and the cfg-3 DP step")
    ap.add_argument("--block-timeout", type=float, default=600.0, help="N > 1: watchdog (s) over the slab / cfg-3 blocks")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
