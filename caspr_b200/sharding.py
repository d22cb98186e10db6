"""Batch sharding of the reconstruction path across the GPUs of one box.

The path shards by SEQUENCE (batch dim): every cross-element reduction of the model is within a
sequence or a cloud (GroupNorm per sample, max-pools per sequence, frozen MovingBatchNorm statistics in
eval), so ranks need no data-path collective.  The reference's only multi-GPU
mechanism is single-process ``nn.DataParallel`` (caspr/train.py:131-132), whose replicas each run their
own adaptive step controller — the same "independent" semantics as here: one process per GPU, rank r
owns sequences [lo, hi) and solves with its own step sequence.

NCCL (or gloo on CPU) is used only to gather results / timings, never inside the reconstruction hot path.
Training (BASELINE config 5) adds the one real exchange step of the reference's multi-GPU path: the gradient
reduction DataParallel performs when it scatters the batch and gathers the replicas' gradients
(``train.py:131-132``).  Here it is ONE sum all-reduce of a flat fp32 gradient buffer (16.26 M floats, 65 MB)
over NCCL / NVLink per step, followed by the identical optimizer step on every rank.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced partition of range(n_items): the first (n_items % world) ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError('bad rank/world: %r/%r' % (rank, world))
    base, extra = divmod(int(n_items), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_batch(local, n_items, group=None):
    """All-gather per-rank result shards (dim 0 = this rank's sequences, possibly empty) back into the
    unsharded order.  Works for uneven shards by padding to the largest one."""
    world = dist.get_world_size(group)
    sizes = [shard_range(n_items, r, world) for r in range(world)]
    max_n = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


def gather_rows(local, group=None):
    """All-gather row blocks of possibly different sizes: returns (all rows in rank order, lo, hi) with
    full[lo:hi] == local."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    pad = torch.zeros((max(counts),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    lo = sum(counts[:rank])
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0), lo, lo + counts[rank]


class lockstep(object):
    """Context manager: inside it the adaptive solvers of `model` take their step decisions over ALL ranks of `group`
    (one all-reduce of two doubles per attempted CNF step, full-batch latent solve), so a sharded batch follows the
    step sequence of the unsharded batch ("lock-step").  Every rank must own at least one sequence
    and make the same sequence of model calls."""

    def __init__(self, model, group=None):
        self.model, self.group = model, group

    def __enter__(self):
        self.model.lockstep_group = self.group
        self.model.point_cnf.lockstep_group = self.group
        return self

    def __exit__(self, *exc):
        self.model.lockstep_group = False
        self.model.point_cnf.lockstep_group = False
        return False


def reconstruct_sharded(model, x, gather=True, group=None, **kwargs):
    """``model.reconstruct`` over this rank's slice of the batch.  Per-sequence keyword tensors ``y`` (base
    samples) and ``e`` (Hutchinson noise) given for the FULL batch are sliced consistently.

    Returns the local 4-tuple, or — with ``gather`` — the tuple gathered over ranks in batch order."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    B, T = x.shape[0], x.shape[1]
    lo, hi = shard_range(B, rank, world)
    kw = dict(kwargs)
    ts = kw.get('timestamps')
    Tq = T if ts is None else int(ts.numel())
    per_seq_y = 1 if kw.get('constant_in_time', False) else Tq
    if kw.get('y') is not None:
        kw['y'] = kw['y'].reshape(B * per_seq_y, *kw['y'].shape[-2:])[lo * per_seq_y:hi * per_seq_y]
    if kw.get('e') is not None:
        kw['e'] = kw['e'].reshape(B * Tq, *kw['e'].shape[-2:])[lo * Tq:hi * Tq]
    if hi > lo:
        out = model.reconstruct(x[lo:hi], **kw)
    else:                                   # more ranks than sequences: this rank idles
        P = kw.get('num_points', 1024)
        z = x.new_zeros((0, Tq, P, 3))
        out = (z, x.new_zeros((0, Tq, P)), z, x.new_zeros((0, T, x.shape[2], 4)))
    if not gather:
        return out
    return tuple(None if o is None else gather_batch(o, B, group) for o in out)


def max_over_ranks(value, device, group=None):
    """Max of a python float over ranks (multi-GPU timings are reported as the slowest rank's)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def allreduce_gradients(parameters, group=None, average=True):
    """Sum (or average) the ``.grad`` of ``parameters`` over ranks with ONE all-reduce of a flat buffer.

    Parameters that received no gradient on this rank (``grad is None``) contribute zeros, so every rank reduces the
    same layout even when a rank's shard was empty.  Returns the number of elements reduced."""
    params = [p for p in parameters if p.requires_grad]
    if not params:
        return 0
    world = dist.get_world_size(group)
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=params[0].device)
    off = 0
    for p in params:
        if p.grad is not None:
            flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
        off += p.numel()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= world
    off = 0
    for p in params:
        g = flat[off:off + p.numel()].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += p.numel()
    return flat.numel()


class FlatGradients(object):
    """Gradients of a model kept as views into ONE persistent flat fp32 buffer, so the per-step all-reduce needs no
    pack / unpack copies (round 1 issued ~440 small copy kernels per step for them).

    ``p.grad`` of every trainable parameter is a view of ``self.flat``; autograd accumulates into the views in place.
    Use ``zero()`` instead of ``optimizer.zero_grad()`` (whose default ``set_to_none=True`` would drop the views —
    ``attach()`` re-establishes them if that happened)."""

    def __init__(self, parameters):
        self.params = [p for p in parameters if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else 'cpu'
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.attach()

    def attach(self):
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                    v.copy_(p.grad)
                p.grad = v

    def zero(self):
        self.attach()
        self.flat.zero_()

    def allreduce(self, group=None, average=True):
        """ONE sum all-reduce of the flat buffer (16.26 M floats = 65 MB for CaSPR).  Returns the element count."""
        self.attach()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat /= dist.get_world_size(group)
        return self.flat.numel()


def moving_batchnorm_buffers(model):
    """The running statistics that feed the NEXT forward pass (normalization.py:60-64): ``running_mean``,
    ``running_var`` and ``step`` of every MovingBatchNorm layer, in module order."""
    bufs = []
    for m in model.modules():
        if hasattr(m, 'running_mean') and hasattr(m, 'running_var') and hasattr(m, 'step'):
            bufs += [m.running_mean, m.running_var, m.step]
    return bufs


def broadcast_moving_batchnorm(model, src=0, group=None):
    """DataParallel semantics for the MovingBatchNorm statistics: the reference's ``nn.DataParallel`` (train.py:131-132)
    re-broadcasts replica 0's buffers before every forward, so all replicas normalise with ONE set of statistics
    (those updated from replica 0's shard, normalization.py:43-64).  Here every rank updates its own copy during the
    step; this sends rank ``src``'s copy to everyone (one broadcast of 14 floats for CaSPR) so the next forward and
    any rank's checkpoint see identical buffers.  Returns the number of elements sent."""
    bufs = moving_batchnorm_buffers(model)
    if not bufs:
        return 0
    flat = torch.cat([b.detach().reshape(-1).to(torch.float32) for b in bufs])
    dist.broadcast(flat, src=src if group is None else dist.get_global_rank(group, src), group=group)
    off = 0
    with torch.no_grad():
        for b in bufs:
            b.copy_(flat[off:off + b.numel()].view_as(b))
            off += b.numel()
    return off


def train_step_sharded(model, optimizer, x, sample_points, loss_fn, group=None, flat_grads=None, **forward_kwargs):
    """One data-parallel training step (train_utils.py:118-175 under ``--parallel``): rank r runs forward + backward on
    its slice of the batch, gradients are averaged with one all-reduce, every rank applies the same optimizer step,
    and rank 0's MovingBatchNorm statistics are broadcast (``broadcast_moving_batchnorm``).

    ``loss_fn(nll, tnocs_l1) -> scalar`` must be a MEAN over the sequences it is given; shards of equal size then
    reproduce the single-process gradient.  ``flat_grads``: a ``FlatGradients`` over ``model.parameters()`` (kept by
    the caller across steps) makes the reduction copy-free; without it the gradients are packed per step.
    Returns this rank's loss (python float; nan for an empty shard)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(x.shape[0], rank, world)
    if flat_grads is not None:
        flat_grads.zero()
    else:
        optimizer.zero_grad()
    loss_value = float('nan')
    if hi > lo:
        kw = dict(forward_kwargs)
        if kw.get('e') is not None:
            T = x.shape[1]
            kw['e'] = kw['e'].reshape(x.shape[0] * T, *kw['e'].shape[-2:])[lo * T:hi * T]
        losses = model(x[lo:hi], sample_points[lo:hi], **kw)
        loss = loss_fn(*losses)
        loss.backward()
        loss_value = float(loss.detach())
    if flat_grads is not None:
        flat_grads.allreduce(group=group, average=True)
    else:
        allreduce_gradients(model.parameters(), group=group, average=True)
    optimizer.step()
    broadcast_moving_batchnorm(model, src=0, group=group)
    return loss_value
