"""CPU, world_size 2, gloo: the N>1 host logic (ray sharding + image gather).  The renderer itself is GPU-only, so the
per-rank render is a stand-in that tags every ray with its global index."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_range_partitions():
    from snerf_b200.parallel import shard_range
    for n in (0, 1, 7, 8, 1440000, 1440001):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from snerf_b200.parallel import render_sharded, shard_range
    rays = torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 11)
    a, b = shard_range(n, rank, world)

    def fake_render(r):
        assert r.shape[0] == b - a and float(r[0, 0]) == a
        return {"rgb_map": r[:, :3] * 2.0, "depth_map": r[:, 0] + 0.5}

    out = render_sharded(rays, fake_render)
    assert torch.equal(out["rgb_map"], rays[:, :3] * 2.0)
    assert torch.equal(out["depth_map"], rays[:, 0] + 0.5)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 1001])
def test_render_sharded_two_ranks(n):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, n), nprocs=2, join=True)


def _grad_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from snerf_b200.parallel import all_reduce_gradients, broadcast_parameters
    torch.manual_seed(100 + rank)                      # ranks start from DIFFERENT parameters
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    ps = list(net.parameters())
    broadcast_parameters(ps, src=0)
    torch.manual_seed(100)
    ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    assert all(torch.equal(a, b) for a, b in zip(ps, ref.parameters()))
    # rank-dependent gradients; parameter 1 has a gradient on rank 1 ONLY (a head one rank's batch left unused): the flat
    # buffer must still cover every parameter on every rank, zeros standing in for the missing gradient
    for i, p in enumerate(ps):
        p.grad = None if (i == 1 and rank == 0) else torch.full_like(p, float(rank + 1) * (i + 1))
    all_reduce_gradients(ps, average=True)
    for i, p in enumerate(ps):
        want = (0.0 + 2.0) / 2 * (i + 1) if i == 1 else 1.5 * (i + 1)      # mean of (1, 2) x (i+1); (0, 2) for parameter 1
        assert p.grad is not None and torch.allclose(p.grad, torch.full_like(p, want)), (i, p.grad)
    all_reduce_gradients(ps, average=False)
    assert torch.allclose(ps[0].grad, torch.full_like(ps[0], 3.0))
    # broadcast_parameters must move every parameter's version counter (NeRF.packed() keys its cache on it)
    v0 = [p._version for p in ps]
    broadcast_parameters(ps, src=1)
    assert all(p._version > v for p, v in zip(ps, v0))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_all_reduce_two_ranks():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_grad_worker, args=(2, port), nprocs=2, join=True)


def _flat_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from snerf_b200 import NeRF
    from snerf_b200.parallel import FlatGradients, broadcast_parameters
    torch.manual_seed(7 + rank)
    nets = [NeRF(D=2, W=64, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True) for _ in range(2)]
    broadcast_parameters([p for n in nets for p in n.parameters()], src=0)
    fg = FlatGradients(nets)                       # every p.grad is a view of ONE buffer; the all-reduce runs on it in place
    n_total = sum(p.numel() for n in nets for p in n.parameters())
    assert fg.flat.numel() == n_total
    for n in nets:
        for _, _, p in n._slots():
            assert p.grad.data_ptr() >= fg.flat.data_ptr() and p.grad.shape == p.shape
    fg.zero()
    for i, p in enumerate(p for n in nets for _, _, p in n._slots()):
        p.grad.add_(float(rank + 1) * (i + 1))     # "the backward kernels accumulate straight into the views"
    before = fg.flat.data_ptr()
    fg.all_reduce(average=True)
    assert fg.flat.data_ptr() == before
    for i, p in enumerate(p for n in nets for _, _, p in n._slots()):
        assert torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1)))       # mean of (1, 2) x (i + 1)
    fg.all_reduce(average=False)
    assert torch.allclose(fg.flat[:4], torch.full((4,), 3.0))
    fg.release()
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradients_in_place_all_reduce_two_ranks():
    """the training step's ONE collective -- the flat 4.77 MB-style gradient buffer summed in place."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_flat_worker, args=(2, port), nprocs=2, join=True)
