"""Pin the training oracle (oracle/train_oracle.py, CPU) against the training-step fixture frozen from the UNMODIFIED
reference modules in .train() mode (tests/golden/make_golden_train.py): loss, per-point NLL, NFE and a summary of
every parameter's gradient (norm, pseudo-random projection, first 64 entries)."""
import os

import numpy as np
import pytest
import torch

from caspr_b200.synth import synthetic_state_dict, synthetic_sequences
from oracle.train_oracle import TrainOracle
from oracle.grad_check import check_gradients


@pytest.fixture(scope='module')
def gold(golden_dir):
    return dict(np.load(os.path.join(golden_dir, 'caspr_train.npz')))


def test_training_step_matches_reference_fixture(gold):
    torch.set_num_threads(8)
    sd = synthetic_state_dict(0, cnf_init='vigorous')
    x, nocs = synthetic_sequences(1, 2, 1024, seed=5)
    orc = TrainOracle(sd)
    nll, tl1 = orc.forward_train(x, nocs, torch.from_numpy(gold['e']))
    loss = TrainOracle.loss(nll, tl1)
    loss.backward()
    assert list(orc.get_nfe()) == [int(v) for v in gold['nfe']]
    assert abs(float(loss) - float(gold['loss'])) / abs(float(gold['loss'])) < 1e-4
    assert np.abs(nll.detach().numpy() - gold['nll']).max() / np.abs(gold['nll']).max() < 1e-4
    grads = {k: v.grad for k, v in orc.parameters().items()}
    grads.update({k.replace('latent_ode.ode_func.', 'latent_ode.solver.ode_func.'): v for k, v in grads.items()
                  if k.startswith('latent_ode.ode_func.')})
    worst = check_gradients(gold, grads)
    print('worst deviation: solver-side %.3g, encoder %.3g' % worst)
