"""Accuracy of the two CNF engines against an fp64 evaluation of the same dynamics (GPU box).
Prints max / rms relative errors of dy and the divergence, and the NFE / output deviation of a
full reverse solve.  Used to decide operand precision (CNF engine section)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caspr_b200 import ops                                   # noqa: E402
from caspr_b200.models import CaSPR                          # noqa: E402
from caspr_b200.models.cnf import SequentialFlow             # noqa: E402
from caspr_b200.synth import synthetic_state_dict            # noqa: E402


def f64_dynamics(sd, t, y, e, ctx):
    p = 'point_cnf.chain.1.odefunc.diffeq.layers.%d.'
    y = y.double().requires_grad_(True)
    tc = torch.cat([torch.full((y.shape[0], 1), t, dtype=torch.float64), ctx.double()], 1)
    dx = y
    for l in range(4):
        W = sd[p % l + '_layer.weight'].double(); b = sd[p % l + '_layer.bias'].double()
        gate = torch.sigmoid(tc @ sd[p % l + '_hyper_gate.weight'].double().t() + sd[p % l + '_hyper_gate.bias'].double())
        bias = tc @ sd[p % l + '_hyper_bias.weight'].double().t()
        dx = (dx @ W.t() + b) * gate.unsqueeze(1) + bias.unsqueeze(1)
        if l < 3:
            dx = torch.nn.functional.softplus(dx)
    ed = torch.autograd.grad(dx, y, e.double())[0]
    return dx.detach(), -(ed * e.double()).sum(-1)


def main():
    dev = 'cuda:0'
    for init in ('vigorous', 'default'):
        sd = synthetic_state_dict(0, cnf_init=init)
        model = CaSPR().to(dev).eval()
        model.load_state_dict(sd)
        g = torch.Generator().manual_seed(3)
        F, P = 4, 2048
        y = torch.randn(F, P, 3, generator=g); e = torch.randn(F, P, 3, generator=g)
        ctx = 0.5 * torch.randn(F, 1600, generator=g)
        dy64, nd64 = f64_dynamics(sd, 0.37, y, e, ctx)
        pack = model.point_cnf.chain[1].weight_pack()
        for name, eng in (('simt_fp32', ops.CNF_SIMT_FP32), ('tc_fp16x3', ops.CNF_TC_FP16X3)):
            dy, nd = ops.cnf_feval(y.to(dev), e.to(dev), ctx.to(dev), pack, 0.37, engine=eng)
            ed = (dy.cpu().double() - dy64); en = (nd.cpu().double() - nd64)
            print('%s %-10s dy: max %.3e rms %.3e (|dy| rms %.3e) | div: max %.3e rms %.3e (|div| rms %.3e)' % (
                init, name, ed.abs().max(), ed.pow(2).mean().sqrt(), dy64.pow(2).mean().sqrt(),
                en.abs().max(), en.pow(2).mean().sqrt(), nd64.pow(2).mean().sqrt()))
        outs = {}
        for name, eng in (('simt_fp32', ops.CNF_SIMT_FP32), ('tc_fp16x3', ops.CNF_TC_FP16X3)):
            SequentialFlow.engine = eng
            x = model.point_cnf(y.to(dev), ctx.to(dev), reverse=True, e=e.to(dev))
            outs[name] = x.cpu()
            print('%s %-10s reverse solve info %s' % (init, name, model.point_cnf.last_info))
        d = (outs['simt_fp32'] - outs['tc_fp16x3']).abs().max() / outs['simt_fp32'].abs().max()
        print('%s solve deviation tc vs simt: %.3e' % (init, d))


if __name__ == '__main__':
    main()
