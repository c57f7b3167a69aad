"""2+ GPU helper (torchrun): cfg 3 -- 2-D Kolmogorov-flow FNO, modes (64,64,1), width 64, 256 x 256 x 1 grid, batch 64 per GPU --
full training step with the gradient averaging done (a) not at all (compute only), (b) by one flat all-reduce after backward,
(c) by the hook-driven overlapped reducer (distributed.OverlappedGradientAllReduce).  537 MB of complex64 spectral-weight
gradients per step.  Device time, max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 scripts/dp_cfg3_bench.py
"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import pde_fluid_phi_b200 as pb
from pde_fluid_phi_b200.distributed import OverlappedGradientAllReduce, allreduce_gradients, broadcast_parameters


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
    n, B = 256, 64
    torch.manual_seed(1234)
    model = pb.FNO3D(modes=(64, 64, 1), width=64, n_layers=4, in_channels=2, out_channels=2).to(dev)
    model.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32")
    broadcast_parameters(model)
    params = list(model.parameters())
    opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=1e-4)
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    yy = torch.linspace(0, 2 * math.pi, n, device=dev)
    base = torch.sin(4 * yy)[None, None, None, :, None].expand(B, 1, n, n, 1)
    x = torch.cat([base + 0.1 * torch.randn(B, 1, n, n, 1, device=dev, generator=g), 0.1 * torch.randn(B, 1, n, n, 1, device=dev, generator=g)], 1)
    y = torch.roll(x, 3, dims=3)
    reducer = OverlappedGradientAllReduce(params, world)
    reducer.enabled = False

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

    res = {"world": world, "grad_bytes": sum(p.numel() * p.element_size() for p in params)}
    grads = {}
    for mode in ("none", "flat", "overlapped"):
        for _ in range(3):
            step(mode)
        # one more backward from identical weights for the cross-check below
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            loss = step(mode)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 8], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[f"ms_per_step_{mode}"] = float(t)
        res[f"samples_per_s_{mode}"] = B * world / float(t) * 1e3
        res[f"loss_{mode}"] = float(loss)
    # the two reductions give the same gradients from the same weights
    for mode in ("flat", "overlapped"):
        opt.zero_grad(set_to_none=True)
        reducer.enabled = mode == "overlapped"
        torch.nn.functional.mse_loss(model(x), y).backward()
        allreduce_gradients(params, world) if mode == "flat" else reducer.finish()
        grads[mode] = [p.grad.clone() for p in params]
    res["overlapped_equals_flat"] = all(torch.equal(a, b) for a, b in zip(grads["flat"], grads["overlapped"]))
    res["max_rel_diff"] = max(float((torch.view_as_real(a) if a.is_complex() else a).sub(torch.view_as_real(b) if b.is_complex() else b).norm()
                                    / (a.abs().norm() + 1e-30)) for a, b in zip(grads["flat"], grads["overlapped"]))
    if rank == 0:
        print(json.dumps(res), flush=True)
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"dp_cfg3_n{world}.json"), "w") as f:
            json.dump(res, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
