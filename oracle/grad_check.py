"""Gradient-fixture helpers shared by the CPU and GPU training tests.  TEST INFRASTRUCTURE ONLY."""
import numpy as np
import torch

# Gradients of the encoder are only reproducible to the flip-noise floor of fp32 (ReLU masks / max-pool winners /
# per-ball GroupNorm on padded balls decided by 1e-6-level forward differences): the reference's own fp32 and fp64
# gradients differ by 3.5 % (median relative L2, measured in DESIGN.md section 2b).  Solver-side gradients are smooth.
SMOOTH_TOL, ENCODER_TOL = 2e-3, 0.25
# Conv biases in front of a GroupNorm with one channel per group: the true gradient is zero, the value is noise
ZERO_GRADIENT = ('set_abstractions.0.pointnet_modules.0.conv_layers.0.bias',
                 'set_abstractions.0.pointnet_modules.0.conv_layers.1.bias')


def grad_summary(index, grad):
    g = grad.detach().double().flatten().cpu()
    probe = torch.randn(g.numel(), generator=torch.Generator().manual_seed(1000 + index), dtype=torch.float64)
    head = torch.zeros(64, dtype=torch.float64)
    head[:min(64, g.numel())] = g[:64]
    return np.concatenate([[float(g.norm()), float(torch.dot(g, probe))], head.numpy()])


def check_gradients(gold, named_grads):
    """named_grads: {reference parameter name: gradient tensor}.  Returns the worst deviations for reporting."""
    names = [str(n) for n in gold['grad_names']]
    worst_smooth, worst_enc = 0.0, 0.0
    for i, name in enumerate(names):
        ref = gold['grad_summary'][i]
        assert name in named_grads, 'no gradient for %s' % name
        mine = grad_summary(i, named_grads[name])
        if name.endswith(ZERO_GRADIENT):
            continue
        norm = max(ref[0], 1e-30)
        dev_norm = abs(mine[0] - ref[0]) / norm
        dev_head = np.abs(mine[2:] - ref[2:]).max() / max(np.abs(ref[2:]).max(), 1e-30)
        # the projection is a sum of numel terms of size ~norm/sqrt(numel): compare on the scale of the norm
        dev_dot = abs(mine[1] - ref[1]) / norm
        dev = max(dev_norm, dev_dot, dev_head if not name.startswith('encoder.') else 0.0)
        if name.startswith('encoder.'):
            worst_enc = max(worst_enc, dev)
            assert dev < ENCODER_TOL, (name, dev_norm, dev_dot, dev_head)
        else:
            worst_smooth = max(worst_smooth, dev)
            assert dev < SMOOTH_TOL, (name, dev_norm, dev_dot, dev_head)
    return worst_smooth, worst_enc
