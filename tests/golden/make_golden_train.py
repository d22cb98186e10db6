"""Freeze a training-step fixture from the UNMODIFIED reference modules (dev container only).

    python tests/golden/make_golden_train.py

Reference ``CaSPR`` in ``.train()`` mode over the oracle shims (oracle/reference_loader.py; torchdiffeq 0.0.1's
``odeint_adjoint`` restated in oracle/odeint001.py), seeded synthetic weights, one forward + ``loss.backward()``
exactly as ``train_utils.py:125-173`` does it.  The full gradient set is 65 MB, so the fixture keeps per parameter:
the L2 norm, the first 64 entries and the dot product with a fixed pseudo-random vector (seed = index of the key).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.reference_loader import build_reference_caspr          # noqa: E402
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences   # noqa: E402

CNF_LOSS_WEIGHT, TNOCS_LOSS_WEIGHT = 0.01, 100.0        # config_utils.py:42-43


def grad_summary(index, grad):
    g = grad.detach().double().flatten()
    probe = torch.randn(g.numel(), generator=torch.Generator().manual_seed(1000 + index), dtype=torch.float64)
    head = torch.zeros(64, dtype=torch.float64)
    head[:min(64, g.numel())] = g[:64]
    return np.concatenate([[float(g.norm()), float(torch.dot(g, probe))], head.numpy()])


def main():
    torch.set_num_threads(8)
    sd = synthetic_state_dict(0, cnf_init='vigorous')
    model = build_reference_caspr()
    model.load_state_dict(sd)
    model.train()
    x, nocs = synthetic_sequences(1, 2, 1024, seed=5)
    torch.manual_seed(7)
    nll, tl1 = model(x, nocs)
    nfe = np.asarray(model.get_nfe())
    loss = CNF_LOSS_WEIGHT * nll.sum(2).mean() + TNOCS_LOSS_WEIGHT * tl1[:, :, :, :4].mean()    # train_utils.py:148-166
    loss.backward()
    out = {'loss': np.asarray(loss.item()), 'nll': nll.detach().numpy(), 'tnocs_l1_mean': np.asarray(tl1.mean().item()),
           'nfe': nfe}
    torch.manual_seed(7)
    out['e'] = torch.randn(2, 1024, 3).numpy()
    names, rows = [], []
    for i, (k, p) in enumerate(model.named_parameters()):
        names.append(k)
        rows.append(grad_summary(i, p.grad))
    out['grad_names'] = np.asarray(names)
    out['grad_summary'] = np.stack(rows)
    for i in (0, 2):        # MovingBatchNorm statistics after the training-mode update (normalization.py:43-51)
        out['mbn%d_running_mean' % i] = model.point_cnf.chain[i].running_mean.numpy()
        out['mbn%d_running_var' % i] = model.point_cnf.chain[i].running_var.numpy()
    path = os.path.join(HERE, 'caspr_train.npz')
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, 'KiB', 'loss', out['loss'], 'nfe', nfe, len(names), 'parameters')


if __name__ == '__main__':
    main()
