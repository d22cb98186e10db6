"""Run one phase of the hot path a few times (for ncu): python tools/profile_phase.py encode|decode|all [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caspr_b200.models import CaSPR                          # noqa: E402
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences   # noqa: E402


def main():
    phase = sys.argv[1] if len(sys.argv) > 1 else 'all'
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = 'cuda:0'
    B, T, N, P = 8, 10, 1024, 2048
    model = CaSPR().to(dev).eval()
    model.load_state_dict(synthetic_state_dict(0, cnf_init='vigorous'))
    model.encoder.use_cuda_graph = False
    x, _ = synthetic_sequences(B, T, N, seed=100)
    x = x.to(dev)
    g = torch.Generator().manual_seed(1000)
    y = torch.randn(B * T, P, 3, generator=g).to(dev)
    e = torch.randn(B * T, P, 3, generator=g).to(dev)
    z0, _ = model.encode(x)
    z = model.aggregate_and_solve_latent(z0, x[:, :, 0, 3] / 5.0)
    torch.cuda.synchronize()
    for _ in range(reps):
        if phase in ('encode', 'all'):
            torch.cuda.nvtx.range_push('encode')
            model.encode(x)
            torch.cuda.nvtx.range_pop()
        if phase in ('decode', 'all'):
            torch.cuda.nvtx.range_push('decode')
            model.decode(z, P, y=y, e=e)
            torch.cuda.nvtx.range_pop()
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()
