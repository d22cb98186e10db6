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


def _replay_fixed_steps(func, y0, log, t_end):
    """Differentiable dopri5 over the ACCEPTED steps of a solver log with constant step sizes + dense output at t_end:
    discretise-then-differentiate WITHOUT gradient paths through the step-size controller."""
    from oracle.odeint001 import DP_ALPHA, DP_BETA, DP_C_MID, Dopri5, _scaled_dot, _State
    y = y0
    f = func(torch.tensor(log[0][0], dtype=torch.float32), y)
    last = None
    for (t0, dt, accept, _) in log:
        if not accept:
            continue
        t0_s, dt_s = torch.tensor(t0, dtype=torch.float32), torch.tensor(dt, dtype=torch.float32)
        k = tuple([f_] for f_ in f)
        yi = y
        for alpha_i, beta_i in zip(DP_ALPHA, DP_BETA):
            yi = tuple(y_ + _scaled_dot(dt_s, beta_i, k_) for y_, k_ in zip(y, k))
            for k_, f_ in zip(k, func(t0_s + alpha_i * dt_s, yi)):
                k_.append(f_)
        f1 = tuple(k_[-1] for k_ in k)
        y_mid = tuple(y_ + _scaled_dot(dt_s, DP_C_MID, k_) for y_, k_ in zip(y, k))
        last = (y, yi, y_mid, f, f1, dt_s, t0, t0 + dt)
        y, f = yi, f1
    y0_, y1_, ym_, f0_, f1_, dt_s, ta, tb = last
    st = _State(y1_, f1_, torch.tensor(ta, dtype=torch.float64), torch.tensor(tb, dtype=torch.float64), None,
                Dopri5._interp_fit(y0_, y1_, ym_, f0_, f1_, dt_s))
    return Dopri5._interp_eval(st, torch.tensor(t_end, dtype=torch.float64))


@pytest.mark.parametrize('tol,bound', [(1e-5, 3e-3), (1e-7, 1e-4)])
def test_adjoint_restatement_converges_to_backprop_through_the_steps(tol, bound):
    """The restated OdeintAdjointMethod (oracle/odeint001.py, the thing every training-parity claim rests on) against
    plain autograd through the same accepted dopri5 steps.  The two are different discretisations of the same gradient:
    they must agree to the solver tolerance and CONVERGE as it is tightened (measured 9.8e-4 at 1e-5, 1.0e-5 at 1e-7).
    (Backprop through torchdiffeq's own adaptive loop is not a usable reference: it also differentiates the step-size
    controller and the NaN-producing initial-step probe.)  Small CNF: 2 frames x 24 points."""
    from oracle import odeint001
    from oracle.train_oracle import _CnfFunc
    torch.set_num_threads(8)
    sd = synthetic_state_dict(0, cnf_init='vigorous')
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 24, 3, generator=g) * 0.3
    e = torch.randn(2, 24, 3, generator=g)
    ctx = torch.randn(2, 1600, generator=g) * 0.5
    gy = torch.randn(2, 24, 3, generator=g)
    gl = torch.randn(2, 24, 1, generator=g)
    orc = TrainOracle(sd)
    func = _CnfFunc(orc.sd, e)
    xo, co = x.clone().requires_grad_(True), ctx.clone().requires_grad_(True)
    s = orc.sd['point_cnf.chain.1.sqrt_end_time']
    times = torch.stack([torch.tensor(0.0), s * s])
    out = odeint001.odeint_adjoint(func, (xo, torch.zeros(2, 24, 1), co), times, atol=[tol] * 3, rtol=[tol] * 3,
                                   method='dopri5')
    fwd_log = list(odeint001.LAST_SOLVER[0].log)                  # before backward() runs the adjoint solve
    la = (out[0][1] * gy).sum() + (out[1][1] * gl).sum()
    la.backward()
    ga = [xo.grad, co.grad] + [orc.sd[n].grad for n in func.names]
    orc2 = TrainOracle(sd)
    func2 = _CnfFunc(orc2.sd, e)
    x2, c2 = x.clone().requires_grad_(True), ctx.clone().requires_grad_(True)
    s2 = orc2.sd['point_cnf.chain.1.sqrt_end_time']
    yT = _replay_fixed_steps(func2, (x2, torch.zeros(2, 24, 1), c2), fwd_log, float((s2 * s2).detach()))
    lb = (yT[0] * gy).sum() + (yT[1] * gl).sum()
    lb.backward()
    gb = [x2.grad, c2.grad] + [orc2.sd[n].grad for n in func2.names]
    assert abs(float(la.detach()) - float(lb.detach())) <= 1e-6 * abs(float(lb.detach()))
    worst = max(float((a - b).abs().max()) / (float(b.abs().max()) + 1e-12) for a, b in zip(ga, gb))
    assert worst < bound, worst


def test_decision_replay_reproduces_the_oracles_own_gradients():
    """The frozen-decision machinery of the oracle encoder (`record` / `decisions`, used by the GPU test of the composed
    encoder backward): replaying the oracle's OWN ReLU sign patterns and max-pool winners must change nothing - same z0,
    same T-NOCS, same parameter gradients, zero differing decisions."""
    torch.set_num_threads(8)
    sd = synthetic_state_dict(0)
    x, _ = synthetic_sequences(2, 2, 256, seed=6)
    g = torch.Generator().manual_seed(8)
    gz, gt = torch.randn(2, 1600, generator=g), torch.randn(2, 2, 256, 4, generator=g)
    a = TrainOracle(sd)
    a.record = {}
    z0a, tna = a.encode(x)
    ((z0a * gz).sum() + (tna * gt).sum()).backward()
    assert len(a.record) == 3 + 5 * 2 * 3 + 5 * 2 + 1 + 3
    b = TrainOracle(sd)
    b.decisions = a.record
    z0b, tnb = b.encode(x)
    ((z0b * gz).sum() + (tnb * gt).sum()).backward()
    assert b.trace['flips']['relu'] == 0 and b.trace['flips']['max'] == 0
    assert torch.equal(z0a, z0b) and torch.equal(tna, tnb)
    pa, pb = a.parameters(), b.parameters()
    for k in pa:
        if k.startswith('encoder.'):
            assert torch.allclose(pa[k].grad, pb[k].grad, rtol=1e-5, atol=1e-7 * float(pa[k].grad.abs().max())), k
