"""torchrun --nproc-per-node N tools/sharded_parity.py : sharded vs unsharded reconstruction on N GPUs.
Every rank also reconstructs the full batch; the gathered sharded result must agree within the solver
tolerance (independent per-rank step control) and FPS-driven encoder outputs must be identical."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caspr_b200.models import CaSPR                          # noqa: E402
from caspr_b200.sharding import reconstruct_sharded, lockstep          # noqa: E402
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences   # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    B, T, N, P = 2 * world, 4, 1024, 512
    model = CaSPR().to(dev).eval()
    model.load_state_dict(synthetic_state_dict(0, cnf_init='vigorous'))
    x, _ = synthetic_sequences(B, T, N, seed=7)
    g = torch.Generator().manual_seed(1)
    y = torch.randn(B * T, P, 3, generator=g).to(dev)
    e = torch.randn(B * T, P, 3, generator=g).to(dev)
    x = x.to(dev)
    full = model.reconstruct(x, num_points=P, y=y, e=e)
    nfe_full = model.get_nfe()
    shard = reconstruct_sharded(model, x, num_points=P, y=y, e=e)
    nfe_shard = model.get_nfe()
    err_x = float((full[2] - shard[2]).abs().max() / full[2].abs().max())
    same_tnocs = bool(torch.equal(full[3], shard[3]))
    if rank == 0:
        print('world %d: sharded vs unsharded rec_x max rel %.3e, tnocs identical %s, nfe full %s shard(rank0) %s'
              % (world, err_x, same_tnocs, nfe_full, nfe_shard))
        assert err_x < 1e-4 and same_tnocs
    # lock-step mode: shared step decisions -> the step sequence (NFE) of the unsharded batch on every rank
    with lockstep(model):
        shard_ls = reconstruct_sharded(model, x, num_points=P, y=y, e=e)
    nfe_ls = model.get_nfe()
    err_ls = float((full[2] - shard_ls[2]).abs().max() / full[2].abs().max())
    nfe_all = [None] * world
    dist.all_gather_object(nfe_all, [int(v) for v in nfe_ls])
    if rank == 0:
        print('lock-step: sharded vs unsharded rec_x max rel %.3e, nfe per rank %s (unsharded %s)'
              % (err_ls, nfe_all, [int(v) for v in nfe_full]))
        assert all(v == [int(q) for q in nfe_full] for v in nfe_all)
        assert err_ls < 2e-5
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
