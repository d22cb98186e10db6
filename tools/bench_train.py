#!/usr/bin/env python
"""Training-step benchmark (BASELINE config 5: airplanes-style step, T=5 frames x 1024 points, batch 8 per GPU).

    python tools/bench_train.py [--steps K] [--warmup W]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py

One step = zero_grad -> CaSPR.forward (encoder, latent ODE, CNF forward flow) -> loss (train_utils.py:148-166) ->
loss.backward() (hand-written encoder backward, CUDA adjoint solves) -> flat-buffer gradient all-reduce over NCCL
(N > 1) -> Adam step (train.py:135).  Weights are re-loaded before every step so that the adaptive solvers see
the same dynamics each time.  Prints one JSON line on rank 0; not the headline metric (that is bench.py)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch                      # noqa: E402
from bench import cpu_training_step          # noqa: E402  (the CPU-baseline leg lives in bench.py)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--frames', type=int, default=5)
    ap.add_argument('--points', type=int, default=1024)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        if rank == 0:
            print(json.dumps(dict(cpu_training_step(args.frames, args.points), impl='reference', n_gpus=world)), flush=True)
        return
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from caspr_b200 import _lib
    from caspr_b200.models import CaSPR
    from caspr_b200.sharding import allreduce_gradients
    from caspr_b200.synth import synthetic_state_dict, synthetic_sequences
    B, T, N = args.batch, args.frames, args.points
    sd = synthetic_state_dict(0, cnf_init='vigorous')
    x, nocs = synthetic_sequences(B, T, N, seed=200 + rank)
    x, nocs = x.to(dev), nocs.to(dev)
    e = torch.randn(B * T, N, 3, generator=torch.Generator().manual_seed(rank)).to(dev)
    model = CaSPR().to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)

    def loss_fn(nll, tl1):
        return 0.01 * nll.sum(2).mean() + 100.0 * tl1[:, :, :, :4].mean()

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    phase = [0.0] * 4

    def step(timed):
        model.load_state_dict(sd)
        torch.cuda.synchronize()
        ev[0].record()
        opt.zero_grad()
        loss = loss_fn(*model(x, nocs, e=e))
        ev[1].record()
        loss.backward()
        ev[2].record()
        if world > 1:
            allreduce_gradients(model.parameters(), average=True)
        ev[3].record()
        opt.step()
        ev[4].record()
        torch.cuda.synchronize()
        if timed:
            for i in range(4):
                phase[i] += ev[i].elapsed_time(ev[i + 1])
        return float(loss.detach())

    for _ in range(max(args.warmup, 3)):
        step(False)
    if world > 1:
        dist.barrier()
    n0 = _lib.lib.caspr_launch_count()
    losses = [step(True) for _ in range(args.steps)]
    launches = _lib.lib.caspr_launch_count() - n0
    ms = sum(phase)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    if rank == 0:
        cnf = model.point_cnf.chain[1]
        line = {'metric': 'trained_points_per_sec', 'value': world * B * T * N * args.steps / (ms * 1e-3), 'unit': 'points/s',
                'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
                'higher_is_better': True, 'scaling': 'weak', 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': 'train_T%d_N%d_B%d' % (T, N, B), 'global_batch': B * world,
                           'phase_ms_forward_backward_allreduce_adam': [round(p / args.steps, 3) for p in phase],
                           'nfe_latent_cnf_forward': [int(v) for v in model.get_nfe()],
                           'cnf_adjoint_info': cnf.last_adjoint_info[:4],
                           'latent_adjoint_info': model.latent_ode.solver.last_adjoint_info[:4],
                           'loss': losses[-1], 'gradient_elements': sum(p.numel() for p in model.parameters())},
                'gpu_launches': int(launches)}
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_training_step(T, N)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
