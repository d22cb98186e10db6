"""torchrun --nproc-per-node N tools/sharded_train_parity.py : one data-parallel training step on N GPUs
(train_step_sharded: forward/backward on the rank's sequences, ONE NCCL all-reduce of the flat gradient buffer, same
optimizer step everywhere) against the same step on the full batch on every rank.  Each rank runs its own adaptive step
controllers (the semantics of the reference's DataParallel replicas), so the gradients agree to the solver tolerance,
not bit for bit; all ranks must end with IDENTICAL parameters."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caspr_b200.models import CaSPR                                        # noqa: E402
from caspr_b200.sharding import train_step_sharded                        # noqa: E402
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences    # noqa: E402


def loss_fn(nll, tl1):
    return 0.01 * nll.sum(2).mean() + 100.0 * tl1[:, :, :, :4].mean()      # train_utils.py:148-166


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    B, T, N = 2 * world, 3, 1024
    sd = synthetic_state_dict(0, cnf_init='vigorous')
    x, nocs = synthetic_sequences(B, T, N, seed=21)
    x, nocs = x.to(dev), nocs.to(dev)
    e = torch.randn(B * T, N, 3, generator=torch.Generator().manual_seed(3)).to(dev)
    # reference: the full batch on this rank
    ref = CaSPR().to(dev).train()
    ref.load_state_dict(sd)
    opt_ref = torch.optim.SGD(ref.parameters(), lr=1e-3)
    opt_ref.zero_grad()
    loss_full = loss_fn(*ref(x, nocs, e=e))
    loss_full.backward()
    g_ref = {k: p.grad.clone() for k, p in ref.named_parameters()}
    # sharded step
    model = CaSPR().to(dev).train()
    model.load_state_dict(sd)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3)
    loss_local = train_step_sharded(model, opt, x, nocs, loss_fn, e=e)
    # conv biases in front of a one-channel-per-group GroupNorm have a true gradient of zero (pure rounding noise)
    noise = ('set_abstractions.0.pointnet_modules.0.conv_layers.0.bias', 'set_abstractions.0.pointnet_modules.0.conv_layers.1.bias')
    errs = []
    for k, p in model.named_parameters():
        if k.endswith(noise):
            continue
        a, b = p.grad.double().flatten(), g_ref[k].double().flatten()
        errs.append((float((a - b).norm() / b.norm().clamp_min(1e-30)), k))
    errs.sort(reverse=True)
    worst, worst_name = errs[0]
    solver_side = max(e for e, k in errs if not k.startswith('encoder.'))
    # every rank must hold the same parameters after the step
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(lo, hi))
    if rank == 0:
        print('world %d: full-batch loss %.6f, rank-0 shard loss %.6f; averaged sharded gradient vs full-batch gradient: '
              'worst relative L2 %.3e (%s), median %.3e, CNF / latent parameters %.3e; parameters identical on all '
              'ranks: %s' % (world, float(loss_full.detach()), loss_local, worst, worst_name, errs[len(errs) // 2][0],
                             solver_side, same))
        assert same and worst < 0.3 and solver_side < 2e-2
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
