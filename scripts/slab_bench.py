"""2+ GPU helper (torchrun): cfg-5-shaped RationalFourierLayer fwd+bwd on x slabs, all-reduce vs all-to-all exchange.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/slab_bench.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import pde_fluid_phi_b200 as pb


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    grid, modes, C = (256, 256, 256), (64, 64, 64), 64
    prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
    torch.manual_seed(3)
    layer = pb.RationalFourierLayer(C, C, modes).cuda()
    with torch.no_grad():
        q0 = layer.Q_coeffs[..., 0, 0, 0].clone()
        layer.Q_coeffs.mul_(1e-6)
        layer.Q_coeffs[..., 0, 0, 0] = q0
    layer.precision = prec
    nxl = grid[0] // world
    g = torch.Generator(device="cuda").manual_seed(100)
    xfull = torch.randn(1, C, *grid, device="cuda", generator=g)
    gfull = torch.randn(1, C, *grid, device="cuda", generator=g)
    x = xfull[:, :, rank * nxl:(rank + 1) * nxl].contiguous().requires_grad_(True)
    go = gfull[:, :, rank * nxl:(rank + 1) * nxl].contiguous()
    res = {}
    outs = {}
    for exchange in ("allreduce", "alltoall"):
        layer.slab = (None, grid[0], exchange)

        def step():
            x.grad = None
            layer.zero_grad(set_to_none=True)
            out = layer(x)
            out.backward(go)
            return out
        for _ in range(2):
            out = step()
        outs[exchange] = (out.detach().clone(), x.grad.clone(), layer.P_coeffs.grad.clone())
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[exchange] = float(t)
    # the two exchanges compute the same layer: outputs and x-gradients agree; P gradients agree after the slab SUM
    a, b = outs["allreduce"], outs["alltoall"]
    gPa, gPb = a[2].clone() * world, b[2].clone()      # allreduce route pre-divides by world (see finish_coeff_grads)
    dist.all_reduce(gPb)
    dist.all_reduce(gPa)
    gPa /= world
    rel = lambda u, v: float((u - v).norm() / v.norm())
    errs = (rel(b[0], a[0]), rel(b[1], a[1]), rel(gPb, gPa))
    if rank == 0:
        # single-GPU layer on the whole volume for reference
        layer.slab = None
        xf = xfull.clone().requires_grad_(True)
        for _ in range(2):
            xf.grad = None
            layer.zero_grad(set_to_none=True)
            of = layer(xf)
            of.backward(gfull)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            xf.grad = None
            layer.zero_grad(set_to_none=True)
            layer(xf).backward(gfull)
        e1.record()
        torch.cuda.synchronize()
        print("cfg-5 layer [1,64,256^3] modes 64^3 %s  world %d: 1 GPU %.2f ms | allreduce %.2f ms | alltoall %.2f ms | "
              "alltoall vs allreduce: out %.1e dx %.1e dP %.1e | alltoall slab vs 1-GPU out %.1e"
              % (prec, world, e0.elapsed_time(e1) / 5, res["allreduce"], res["alltoall"], *errs,
                 rel(b[0], of[:, :, :nxl].detach())), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
