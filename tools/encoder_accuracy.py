"""Encoder accuracy study (GPU box): z0 / T-NOCS / reconstruction error against the reference fixture
with the dense layers on the exact-fp32 SIMT kernel vs the tcgen05 fp16x3 GEMM."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caspr_b200 import ops                                   # noqa: E402
from caspr_b200.models import CaSPR                          # noqa: E402
from caspr_b200.models.cnf import SequentialFlow             # noqa: E402
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences   # noqa: E402


def rel(a, b):
    a = a.detach().cpu().double().numpy()
    return np.abs(a - b).max() / np.abs(b).max()


def main():
    dev = 'cuda:0'
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for tag, init in (('vig', 'vigorous'), ('def', 'default')):
        gold = dict(np.load(os.path.join(here, 'tests', 'golden', 'caspr_%s.npz' % tag)))
        sd = synthetic_state_dict(0, cnf_init=init)
        model = CaSPR().to(dev).eval()
        model.load_state_dict(sd)
        model.encoder.use_cuda_graph = False
        x, _ = synthetic_sequences(1, 3, 1024, seed=1)
        y = torch.from_numpy(gold['rec_y']).reshape(3, 256, 3)
        e = torch.from_numpy(gold['rec_e']).to(dev)
        for lin in ('simt', 'auto'):
            ops.LINEAR_ENGINE = lin
            for cnf_eng, cname in ((ops.CNF_SIMT_FP32, 'simt'), (ops.CNF_TC_FP16X3, 'tc')):
                SequentialFlow.engine = cnf_eng
                z0, tn = model.encode(x.to(dev))
                _, _, xr, _ = model.reconstruct(x.to(dev), num_points=256, y=y, e=e)
                print('%s linear=%-4s cnf=%-4s z0 %.2e tnocs %.2e rec_x %.2e nfe %s' % (
                    tag, lin, cname, rel(z0, gold['z0']), rel(tn, gold['tnocs']), rel(xr, gold['rec_x']),
                    model.get_nfe()))


if __name__ == '__main__':
    main()
