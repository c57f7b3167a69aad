"""Batch data-parallelism for the spectral models: one process per GPU, gradients averaged with a
single flat all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests).

Mirrors what the reference's DistributedTrainer does after backward (training/distributed.py:221-224:
all_reduce(SUM) then divide by world size), with the complex64 SpectralConv3D weights carried through
the same fp32 bucket via view_as_real.  The whole RFNO-3D cfg-2 gradient set is 8.46 MB, i.e. one
latency-bound collective per step.
"""
from __future__ import annotations

import contextlib
from typing import Iterable, List

import torch
import torch.distributed as dist


def _real_view(t: torch.Tensor) -> torch.Tensor:
    return torch.view_as_real(t) if t.is_complex() else t


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world_size: int = None, group=None,
                        average: bool = True) -> None:
    """Combine .grad of every parameter across ranks, in place, with one all-reduce of a flat fp32 bucket:
    the mean for batch data-parallelism (average=True), the sum for slab parallelism (average=False).

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
        if average:
            flat.div_(world_size)
        off = 0
        for g in same:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n


class OverlappedGradientAllReduce:
    """Gradient averaging that overlaps the backward pass (cfg 3: four complex64 `weights` gradients
    of 134 MB each -- too much to leave behind the last backward kernel in one flat bucket).

    A post-accumulate-grad hook on every parameter fires as soon as autograd has finished that parameter's gradient:
    a gradient of at least `bucket_bytes` is all-reduced IN PLACE, asynchronously, while the backward of the layers
    below it is still running (no staging copy at all); smaller ones collect in a flat bucket that is sent when it
    reaches `bucket_bytes`.  `finish()` -- call it after `backward()`, before the optimizer -- sends the last partial
    bucket, waits for every collective and applies the 1/world_size.  The hooks fire in autograd's order, which is the
    same on every rank for replicas of one model, so the collectives pair up without any negotiation; the result equals
    `allreduce_gradients` (the same sums, element by element).  For gradient accumulation wrap the non-final backward
    passes in `no_sync()`, as with DistributedDataParallel.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], world_size: int = None, group=None, average: bool = True,
                 bucket_bytes: int = 32 << 20):
        initialized = dist.is_available() and dist.is_initialized()       # a single process without a group: a no-op
        self.world_size = world_size if world_size is not None else (dist.get_world_size(group) if initialized else 1)
        self.group, self.average, self.bucket_bytes = group, average, int(bucket_bytes)
        self.enabled = True
        # NCCL averages inside the collective (ReduceOp.AVG): no separate pass over a 134 MB gradient for the 1/world_size
        self._fused_avg = bool(average) and initialized and dist.get_backend(group) == "nccl"
        self._op = dist.ReduceOp.AVG if self._fused_avg else dist.ReduceOp.SUM
        self._pending = []            # (work, [gradient views], flat bucket or None)
        self._small, self._small_bytes = [], 0
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params if p.requires_grad]

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if not self.enabled or self.world_size == 1 or p.grad is None:
            return
        g = _real_view(p.grad)
        nbytes = g.numel() * g.element_size()
        if nbytes >= self.bucket_bytes and g.is_contiguous():
            self._pending.append((dist.all_reduce(g, op=self._op, group=self.group, async_op=True), [g], None))
            return
        self._small.append(g)
        self._small_bytes += nbytes
        if self._small_bytes >= self.bucket_bytes:
            self._flush_small()

    def _flush_small(self) -> None:
        by_dtype = {}
        for g in self._small:
            by_dtype.setdefault(g.dtype, []).append(g)
        for same in by_dtype.values():
            flat = torch.cat([g.reshape(-1) for g in same])
            self._pending.append((dist.all_reduce(flat, op=self._op, group=self.group, async_op=True), same, flat))
        self._small, self._small_bytes = [], 0

    def finish(self) -> None:
        """Complete every outstanding collective; afterwards each .grad holds the mean (or the sum) over the ranks."""
        if self._small:
            self._flush_small()
        for work, grads, flat in self._pending:
            work.wait()
            scale = self.average and not self._fused_avg
            if flat is None:
                if scale:
                    grads[0].div_(self.world_size)
                continue
            if scale:
                flat.div_(self.world_size)
            off = 0
            for g in grads:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n
        self._pending = []

    @contextlib.contextmanager
    def no_sync(self):
        """Backward passes inside this context only accumulate locally (no collective is issued)."""
        was, self.enabled = self.enabled, False
        try:
            yield
        finally:
            self.enabled = was

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s weights and buffers (training/distributed.py:229-230)."""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(_real_view(t.data), src=src, group=group)


def enable_slab_parallel(model: torch.nn.Module, nx_total: int, group=None, exchange: str = "allreduce") -> torch.nn.Module:
    """Slab-decomposed transform for grids too large for one GPU (256^3 and above).

    exchange="allreduce" (described below) replicates the retained modes; exchange="alltoall" is the transpose
    formulation: the (z,y) stages run on the local x slab, an NCCL all-to-all re-partitions the small intermediate
    [B,C,Nx,my,mzh] from x slabs to ky subsets (69 MB per rank and direction at 256^3 / modes 64^3 / 2 ranks, 30 MB at
    8 ranks), the x stage, the channel mix and the projection run on this rank's ky subset only (the projection
    energies and its backward's inner product take one tiny all-reduce each), and the inverse mirrors it.  It shards the
    channel mix -- the dominant cost at 64^3 modes -- as well as the transforms; modes[1] must divide by the group size.

    Every rank of `group` holds the x slab [rank*Nx/P, (rank+1)*Nx/P) of each activation.  Inside a spectral
    layer the fused (z,y) stage and the x stage run on the local slab and produce a *partial* pruned spectrum;
    one NCCL all-reduce of that small tensor (69 MB per sample at 256^3 / modes 64^3 / width 64, versus 277 MB
    for an all-to-all of the pre-x-stage intermediate and 4.3 GB for the raw slab) gives every rank the full
    retained modes; the channel mix and projection run replicated; the synthesis x stage then evaluates only
    the local rows, so the inverse needs no communication at all.  All other layers of the operator models
    are pointwise in space and stay local.  After backward, call
    ``allreduce_gradients(model.parameters(), group=group, average=False)``: pointwise-layer gradients are
    slab-local partial sums, spectral-coefficient gradients are pre-divided by the group size ("allreduce") or are
    per-ky-subset partial sums ("alltoall").
    """
    if exchange not in ("allreduce", "alltoall"):
        raise ValueError(f"unknown slab exchange {exchange!r} (expected 'allreduce' or 'alltoall')")
    from .operators import RationalFourierLayer, SpectralConv3D
    for m in model.modules():
        if isinstance(m, (RationalFourierLayer, SpectralConv3D)):
            m.slab = (group, int(nx_total), exchange)
    return model
