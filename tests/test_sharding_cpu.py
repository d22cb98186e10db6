"""N>1 host logic on CPU: world_size-2 gloo processes exercise the batch sharding / gathering that
bench.py and multi-GPU reconstruction use (no GPU, no CUDA library calls)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from caspr_b200.sharding import (shard_range, gather_batch, reconstruct_sharded, max_over_ranks,
                                 allreduce_gradients, train_step_sharded, gather_rows, lockstep, FlatGradients,
                                 broadcast_moving_batchnorm, moving_batchnorm_buffers)


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


class _TinyTrainModel(torch.nn.Module):
    """Stands in for CaSPR.forward in training mode: per-sequence losses from a linear map."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.lin = torch.nn.Linear(4, 3)
        self.unused = torch.nn.Parameter(torch.zeros(2))         # never receives a gradient

    def forward(self, x, sample_points, e=None):
        y = self.lin(x)                                           # (B,T,N,3)
        nll = (y ** 2).sum(-1)
        if e is not None:
            nll = nll + (y * e.view(x.shape[0], x.shape[1], -1, 3)).sum(-1)
        return nll, (y - sample_points[..., :3]).abs()


def _loss(nll, tl1):
    return 0.01 * nll.sum(2).mean() + 100.0 * tl1.mean()


def _train_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1)
        B, T, N = 4, 2, 5
        x = torch.randn(B, T, N, 4, generator=g)
        nocs = torch.randn(B, T, N, 4, generator=g)
        e = torch.randn(B * T, N, 3, generator=g)
        ref = _TinyTrainModel()
        opt_ref = torch.optim.SGD(ref.parameters(), lr=0.1)
        opt_ref.zero_grad()
        _loss(*ref(x, nocs, e=e)).backward()
        opt_ref.step()
        model = _TinyTrainModel()
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        loss = train_step_sharded(model, opt, x, nocs, _loss, e=e)
        ok = all(torch.allclose(a, b, atol=1e-6) for a, b in zip(model.parameters(), ref.parameters()))
        ok = ok and loss == loss
        # gradient layout with a missing gradient on one rank only
        m2 = _TinyTrainModel()
        if rank == 0:
            m2.lin.weight.grad = torch.ones_like(m2.lin.weight)
        n = allreduce_gradients(m2.parameters(), average=False)
        ok = ok and n == sum(p.numel() for p in m2.parameters())
        ok = ok and torch.equal(m2.lin.weight.grad, torch.ones_like(m2.lin.weight))
        ok = ok and torch.equal(m2.unused.grad, torch.zeros(2))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_world_size_2_gloo():
    """Sharded training step == single-process step on the full batch (equal shards, mean loss)."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def _rows_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # lock-step latent solve: every rank gathers the z0 rows of all ranks (uneven shards) and keeps its own slice
        n_local = 3 if rank == 0 else 1
        local = torch.arange(n_local * 4, dtype=torch.float32).view(n_local, 4) + 100 * rank
        full, lo, hi = gather_rows(local)
        ok = full.shape == (4, 4) and (lo, hi) == ((0, 3) if rank == 0 else (3, 4)) and torch.equal(full[lo:hi], local)
        ok = ok and torch.equal(full[3], torch.arange(4, dtype=torch.float32) + 100)

        class _M(object):
            lockstep_group = False

            def __init__(self):
                self.point_cnf = type('F', (), {'lockstep_group': False})()
        m = _M()
        with lockstep(m, group=None):
            ok = ok and m.lockstep_group is None and m.point_cnf.lockstep_group is None
        ok = ok and m.lockstep_group is False and m.point_cnf.lockstep_group is False
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_rows_and_lockstep_switch_world_size_2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rows_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


class _TinyFlowModel(torch.nn.Module):
    """A trainable map followed by the product's MovingBatchNorm1d (pure torch in training mode): its running statistics
    are updated from the local shard and are the ones USED by the next forward (normalization.py:60-64)."""

    def __init__(self):
        super().__init__()
        from caspr_b200.models.cnf import MovingBatchNorm1d
        torch.manual_seed(0)
        self.lin = torch.nn.Linear(4, 3)
        self.mbn = MovingBatchNorm1d(3)

    def forward(self, x, sample_points, e=None):
        B, T, N, _ = x.shape
        y = self.mbn(self.lin(x).view(B * T, N, 3)).view(B, T, N, 3)
        return (y ** 2).sum(-1), (y - sample_points[..., :3]).abs()


def _mbn_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(2)
        B, T, N = 4, 2, 6
        model = _TinyFlowModel().train()
        flat = FlatGradients(model.parameters())
        opt = torch.optim.SGD(model.parameters(), lr=0.05)
        ok = len(moving_batchnorm_buffers(model)) == 3
        for step in range(3):
            x = torch.randn(B, T, N, 4, generator=g) + 0.5 * step
            nocs = torch.randn(B, T, N, 4, generator=g)
            train_step_sharded(model, opt, x, nocs, _loss, flat_grads=flat)
            # gradients are still views of the one flat buffer after backward + optimizer step
            ok = ok and all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(flat.params, flat.views))
            state = torch.cat([b.reshape(-1) for b in moving_batchnorm_buffers(model)] +
                              [p.detach().reshape(-1) for p in model.parameters()])
            both = [torch.empty_like(state) for _ in range(world)]
            dist.all_gather(both, state)
            ok = ok and torch.equal(both[0], both[1])          # buffers AND parameters identical on every rank
            ok = ok and float(model.mbn.step) == step + 1
        # DataParallel semantics: the statistics are those of rank 0's shard (sequences [0, 2))
        ref = _TinyFlowModel().train()
        g = torch.Generator().manual_seed(2)
        x = torch.randn(B, T, N, 4, generator=g)
        ref(x[:2], None if False else torch.zeros(2, T, N, 4))
        first = _TinyFlowModel().train()
        g = torch.Generator().manual_seed(2)
        x = torch.randn(B, T, N, 4, generator=g)
        nocs = torch.randn(B, T, N, 4, generator=g)
        train_step_sharded(first, torch.optim.SGD(first.parameters(), lr=0.0), x, nocs, _loss)
        ok = ok and torch.allclose(first.mbn.running_mean, ref.mbn.running_mean, atol=1e-7)
        ok = ok and torch.allclose(first.mbn.running_var, ref.mbn.running_var, atol=1e-7)
        # without the broadcast the ranks would have drifted apart
        drift = _TinyFlowModel().train()
        lo = 2 * rank
        drift(x[lo:lo + 2], torch.zeros(2, T, N, 4))
        state = drift.mbn.running_mean.clone()
        both = [torch.empty_like(state) for _ in range(world)]
        dist.all_gather(both, state)
        ok = ok and not torch.equal(both[0], both[1])
        n = broadcast_moving_batchnorm(drift)
        ok = ok and n == 7
        state = drift.mbn.running_mean.clone()
        dist.all_gather(both, state)
        ok = ok and torch.equal(both[0], both[1])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_moving_batchnorm_buffers_stay_identical_across_ranks_gloo():
    """Three sharded steps: after every step all ranks hold identical MovingBatchNorm statistics (rank 0's, as
    nn.DataParallel's replica 0 does in the reference) and identical parameters; gradients live in one flat buffer."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mbn_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
