"""world_size-2 gloo test of the data-parallel gradient path (host logic; the GPU box runs NCCL)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pde_fluid_phi_b200.distributed import OverlappedGradientAllReduce, allreduce_gradients, broadcast_parameters


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                       # different weights per rank before the broadcast
    lin = torch.nn.Linear(3, 2)
    w = torch.nn.Parameter(torch.randn(2, 2, 3, dtype=torch.cfloat))   # stands in for SpectralConv3D.weights
    mod = torch.nn.Module()
    mod.lin, mod.w = lin, w
    broadcast_parameters(mod, src=0)
    torch.manual_seed(7 + rank)                         # different data per rank
    x = torch.randn(4, 3)
    loss = (lin(x) ** 2).mean() + (w * torch.randn(2, 2, 3, dtype=torch.cfloat)).abs().pow(2).sum()
    loss.backward()
    local = [p.grad.clone() for p in mod.parameters()]
    allreduce_gradients(mod.parameters(), world)
    out[rank] = dict(local=local, reduced=[p.grad.clone() for p in mod.parameters()],
                     weights=[p.detach().clone() for p in mod.parameters()])
    dist.destroy_process_group()


def test_gradient_allreduce_two_ranks():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    for a, b in zip(r0["weights"], r1["weights"]):
        assert torch.equal(a, b)                        # broadcast made the replicas identical
    for l0, l1, g0, g1 in zip(r0["local"], r1["local"], r0["reduced"], r1["reduced"]):
        assert torch.equal(g0, g1)                      # same averaged gradient on both ranks
        assert torch.allclose(g0, (l0 + l1) / 2, rtol=1e-6, atol=1e-7)
        assert g0.dtype == l0.dtype                     # complex64 stays complex64


def _overlap_worker(rank, world, port, out):
    """The hook-driven reducer (large gradients in place, small ones bucketed) against the flat all-reduce."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(3)                                 # replicas: same weights on every rank
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.GELU(), torch.nn.Linear(16, 16), torch.nn.GELU(), torch.nn.Linear(16, 4))
    w = torch.nn.Parameter(torch.randn(4, 4, 5, dtype=torch.cfloat))       # a "large" complex gradient: 640 bytes
    params = list(net.parameters()) + [w]
    torch.manual_seed(50 + rank)                         # different data per rank
    x = torch.randn(8, 6)
    z = torch.randn(4, 4, 5, dtype=torch.cfloat)

    def loss_fn():
        return (net(x) ** 2).mean() + (w * z).abs().pow(2).sum() * net[4].bias.sum().abs()

    loss_fn().backward()
    allreduce_gradients(params, world)
    want = [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    red = OverlappedGradientAllReduce(params, world, bucket_bytes=512)     # w and net[2].weight (1 KB) go in place
    loss_fn().backward()
    assert red._pending, "no collective was issued during backward"
    red.finish()
    got = [p.grad.clone() for p in params]
    # accumulation: first pass local only, second pass reduces the accumulated sum -> mean of the per-rank sums
    for p in params:
        p.grad = None
    with red.no_sync():
        loss_fn().backward()
    assert not red._pending and not red._small
    local2 = [2 * p.grad.clone() for p in params]
    loss_fn().backward()
    red.finish()
    acc = [p.grad.clone() for p in params]
    red.remove()
    out[rank] = dict(want=want, got=got, acc=acc, local2=local2)
    dist.destroy_process_group()


def test_overlapped_gradient_allreduce_two_ranks():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_overlap_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    for w0, g0, g1 in zip(r0["want"], r0["got"], r1["got"]):
        assert g0.dtype == w0.dtype and torch.equal(g0, g1)
        assert torch.equal(g0, w0)                       # the same sums as the flat all-reduce, element by element
    for a0, a1, l0, l1 in zip(r0["acc"], r1["acc"], r0["local2"], r1["local2"]):
        assert torch.equal(a0, a1)
        assert torch.allclose(a0, (l0 + l1) / 2, rtol=1e-5, atol=1e-7)


def test_overlapped_reducer_is_a_no_op_in_a_single_process():
    lin = torch.nn.Linear(3, 2)
    red = OverlappedGradientAllReduce(lin.parameters())          # no process group: world size 1
    lin(torch.randn(4, 3)).sum().backward()
    want = [p.grad.clone() for p in lin.parameters()]
    assert not red._pending and not red._small
    red.finish()
    assert all(torch.equal(p.grad, w) for p, w in zip(lin.parameters(), want))
    red.remove()


def _slab_worker(rank, world, port, out):
    """Slab-decomposed analysis: local partial spectrum (oracle arithmetic) + one all-reduce == full transform."""
    import numpy as np
    from oracle import dense_fp64 as d64
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 3, 8, 6, 10))            # same volume on every rank; each takes its x slab
    nxl = x.shape[2] // world
    mh = (4, 3, 4)
    part = d64.analysis_slab(x[:, :, rank * nxl:(rank + 1) * nxl], mh, x.shape[2], rank * nxl)
    t = torch.view_as_real(torch.from_numpy(part).contiguous())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out[rank] = (torch.view_as_complex(t).numpy(), d64.analysis(x, mh))
    dist.destroy_process_group()


def test_slab_partial_spectra_allreduce_two_ranks():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_slab_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    import numpy as np
    for r in range(world):
        got, want = out[r]
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-12


def _alltoall_worker(rank, world, port, out):
    """Slab-decomposed transform, all-to-all route: (z,y) stages on the x slab, transpose to ky subsets, x stage;
    then the inverse transpose brings a ky-subset tensor back to x slabs.  numpy FFTs stand in for the kernels."""
    import numpy as np
    from pde_fluid_phi_b200.functional import exchange_x_to_ky, exchange_ky_to_x
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    x = rng.standard_normal((2, 3, 8, 6, 10))
    mx, my, mzh = 4, 4, 5
    nxl, myl = x.shape[2] // world, my // world
    T = np.fft.rfft2(x[:, :, rank * nxl:(rank + 1) * nxl], axes=(-2, -1))[..., :my, :mzh]     # [B,C,Nxl,my,mzh]
    T2 = exchange_x_to_ky(torch.from_numpy(np.ascontiguousarray(T)).to(torch.complex64), None, world)
    assert tuple(T2.shape) == (2, 3, 8, myl, mzh)
    X = np.fft.fft(T2.numpy().astype(np.complex128), axis=2)[:, :, :mx]                         # x stage on the ky subset
    back = exchange_ky_to_x(T2, None, world)                                                  # inverse transpose
    out[rank] = (X, back.numpy(), T.astype(np.complex64))
    dist.destroy_process_group()


def test_slab_alltoall_transposes_two_ranks():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_alltoall_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    import numpy as np
    rng = np.random.default_rng(11)
    x = rng.standard_normal((2, 3, 8, 6, 10))
    want = np.fft.rfftn(x, axes=(-3, -2, -1))[:, :, :4, :4, :5]
    for r in range(world):
        X, back, T = out[r]
        ref = want[:, :, :, r * 2:(r + 1) * 2, :]
        assert np.linalg.norm(X - ref) / np.linalg.norm(ref) < 1e-6        # complex64 exchange buffers
        assert np.array_equal(back, T)                                      # the two transposes are exact inverses
