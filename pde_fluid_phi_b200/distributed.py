"""Batch data-parallelism for the spectral models: one process per GPU, gradients averaged with a
single flat all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests).

Mirrors what the reference's DistributedTrainer does after backward (training/distributed.py:221-224:
all_reduce(SUM) then divide by world size), with the complex64 SpectralConv3D weights carried through
the same fp32 bucket via view_as_real.  The whole RFNO-3D cfg-2 gradient set is 8.46 MB, i.e. one
latency-bound collective per step.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


def _real_view(t: torch.Tensor) -> torch.Tensor:
    return torch.view_as_real(t) if t.is_complex() else t


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world_size: int = None, group=None) -> None:
    """Average .grad of every parameter across ranks, in place, with one all-reduce of a flat fp32 bucket.

    Deterministic: the bucket layout is the parameter order, identical on every rank."""
    if world_size is None:
        world_size = dist.get_world_size(group)
    if world_size == 1:
        return
    grads: List[torch.Tensor] = [_real_view(p.grad) for p in params if p.grad is not None]
    if not grads:
        return
    by_dtype = {}
    for g in grads:
        by_dtype.setdefault(g.dtype, []).append(g)
    for same in by_dtype.values():
        flat = torch.cat([g.reshape(-1) for g in same])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world_size)
        off = 0
        for g in same:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s weights and buffers (training/distributed.py:229-230)."""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(_real_view(t.data), src=src, group=group)
