"""N>1 host logic on CPU: world_size-2 gloo processes exercise the batch sharding / gathering that
bench.py and multi-GPU reconstruction use (no GPU, no CUDA library calls)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from caspr_b200.sharding import shard_range, gather_batch, reconstruct_sharded, max_over_ranks


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


class _StubModel(object):
    """Stands in for CaSPR.reconstruct: a deterministic per-sequence function, so sharded == unsharded."""

    def reconstruct(self, x, num_points=4, constant_in_time=False, timestamps=None, y=None, e=None):
        B, T = x.shape[0], x.shape[1]
        Tq = T if timestamps is None else timestamps.numel()
        s = x.sum(dim=(1, 2, 3)).view(B, 1, 1, 1)
        yy = y.view(B, -1, num_points, 3) if y is not None else torch.zeros(B, Tq, num_points, 3)
        if constant_in_time:
            yy = yy.expand(B, Tq, num_points, 3)
        ee = e.view(B, Tq, num_points, 3)
        xr = yy * 2.0 + s + ee
        return yy.contiguous(), yy.sum(-1), xr, x.clone()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, B, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        T, N, P = 3, 5, 4
        x = torch.randn(B, T, N, 4, generator=g)
        y = torch.randn(B * T, P, 3, generator=g)
        e = torch.randn(B * T, P, 3, generator=g)
        model = _StubModel()
        full = model.reconstruct(x, num_points=P, y=y, e=e)
        got = reconstruct_sharded(model, x, num_points=P, y=y, e=e)
        ok = all(torch.equal(a, b) for a, b in zip(full, got))
        # interpolated reconstruction: shared base cloud per sequence, 6 query times
        ts = torch.linspace(0, 1, 6)
        y2 = torch.randn(B, P, 3, generator=g)
        e2 = torch.randn(B * 6, P, 3, generator=g)
        full2 = model.reconstruct(x, num_points=P, constant_in_time=True, timestamps=ts, y=y2, e=e2)
        got2 = reconstruct_sharded(model, x, num_points=P, constant_in_time=True, timestamps=ts, y=y2, e=e2)
        ok = ok and all(torch.equal(a, b) for a, b in zip(full2, got2))
        local = torch.full((shard_range(B, rank, world)[1] - shard_range(B, rank, world)[0], 2), float(rank))
        gathered = gather_batch(local, B)
        ok = ok and gathered.shape == (B, 2)
        ok = ok and max_over_ranks(10.0 + rank, 'cpu') == 10.0 + world - 1
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('B', [5, 2, 1])
def test_world_size_2_gloo(B):
    """B=5: uneven shards; B=1: one rank idles with an empty shard."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
