"""Differentiable CPU restatement of the CaSPR training step (BASELINE config 5).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Same math as ``CasprOracle`` but with the
trainable state_dict entries turned into leaf tensors, the CNF and the latent ODE solved through
``odeint001.odeint_adjoint`` exactly as the reference does in training mode
(``cnf.py:101-111``, ``latent_ode_model.py:98``), and the loss of ``train_utils.py:148-166``.
Pinned against the unmodified reference modules by ``tests/test_train_oracle.py`` (CPU) and the
fixtures of ``tests/golden/make_golden_train.py``.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import odeint001
from .caspr_oracle import CasprOracle

NOT_TRAINABLE = ('running_mean', 'running_var', 'step', '_num_evals')


class _CnfFunc(nn.Module):
    """ODEfunc.forward (odefunc.py:119-142) in training mode; parameters() order = the reference's."""

    def __init__(self, sd, e):
        super().__init__()
        self.names = []
        p = 'point_cnf.chain.1.odefunc.diffeq.layers.%d.'
        for l in range(4):
            for suffix in ('_layer.weight', '_layer.bias', '_hyper_bias.weight', '_hyper_gate.weight',
                           '_hyper_gate.bias'):
                key = (p % l) + suffix
                self.register_parameter('p%d' % len(self.names), sd[key])
                self.names.append(key)
        self.sd = sd
        self.e = e
        self.nfe = 0

    def _layer(self, l, tc, x):
        p = 'point_cnf.chain.1.odefunc.diffeq.layers.%d.' % l
        sd = self.sd
        gate = torch.sigmoid(F.linear(tc, sd[p + '_hyper_gate.weight'], sd[p + '_hyper_gate.bias'])).unsqueeze(1)
        bias = F.linear(tc, sd[p + '_hyper_bias.weight']).unsqueeze(1)
        return F.linear(x, sd[p + '_layer.weight'], sd[p + '_layer.bias']) * gate + bias   # diffeq_layers.py:83-90

    def forward(self, t, states):
        y, _, c = states
        self.nfe += 1
        tt = torch.ones(y.size(0), 1).to(y) * t.clone().detach().requires_grad_(True).type_as(y)   # :121
        for s in states:
            s.requires_grad_(True)                                                    # :123-124
        with torch.set_grad_enabled(True):
            tc = torch.cat([tt, c.view(y.size(0), -1)], dim=1)                        # :133
            dx = y
            for l in range(4):                                                        # odefunc.py:98-105
                dx = self._layer(l, tc, dx)
                if l < 3:
                    dx = F.softplus(dx)
            e_dzdx = torch.autograd.grad(dx, y, self.e, create_graph=True)[0]         # :14
            div = (e_dzdx * self.e).sum(dim=-1).unsqueeze(-1)                         # :15,:26,:135
            return dx, -div, torch.zeros_like(c).requires_grad_(True)                 # :136


class _LatentFunc(nn.Module):
    """DynamicsNet.forward (latent_ode_model.py:139-147)."""

    def __init__(self, sd):
        super().__init__()
        self.sd = sd
        p = 'latent_ode.ode_func.dynamics_net.'
        for i, l in enumerate((0, 2, 4, 6)):
            self.register_parameter('w%d' % i, sd[p + '%d.weight' % l])
            self.register_parameter('b%d' % i, sd[p + '%d.bias' % l])
        self.nfe = 0

    def forward(self, t, z):
        self.nfe += 1
        p = 'latent_ode.ode_func.dynamics_net.'
        sd = self.sd
        h = torch.tanh(F.linear(z, sd[p + '0.weight'], sd[p + '0.bias']))
        h = torch.tanh(F.linear(h, sd[p + '2.weight'], sd[p + '2.bias']))
        h = torch.tanh(F.linear(h, sd[p + '4.weight'], sd[p + '4.bias']))
        return F.linear(h, sd[p + '6.weight'], sd[p + '6.bias'])


class TrainOracle(CasprOracle):
    def __init__(self, state_dict, **kw):
        super().__init__(state_dict, **kw)
        # the latent dynamics are registered twice in the reference; one tensor serves both
        for k, v in list(self.sd.items()):
            if k.startswith('latent_ode.solver.ode_func.'):
                continue
            if v.is_floating_point() and not k.endswith(NOT_TRAINABLE):
                self.sd[k] = nn.Parameter(v.clone())
        for k in list(self.sd):
            if k.startswith('latent_ode.solver.ode_func.'):
                self.sd[k] = self.sd[k.replace('latent_ode.solver.ode_func.', 'latent_ode.ode_func.')]

    def parameters(self):
        return {k: v for k, v in self.sd.items()
                if isinstance(v, nn.Parameter) and not k.startswith('latent_ode.solver.ode_func.')}

    def zero_grad(self):
        for v in self.parameters().values():
            v.grad = None

    # ----------------------------------------------------------------- differentiable pieces
    def latent_ode(self, z0, t):
        """latent_ode_model.py:45-70 through odeint_adjoint, rtol = atol = 1e-3."""
        func = _LatentFunc(self.sd)
        rel_t = t - t[0]
        pred = odeint001.odeint_adjoint(func, z0, rel_t, rtol=1e-3, atol=1e-3, method='dopri5')
        self.nfe[0] = func.nfe
        return pred.permute(1, 0, 2)

    def cnf_train(self, x, context, logpx, e):
        """cnf.py:70-128 in training mode (forward direction): odeint_adjoint, list tolerances."""
        s = self.sd['point_cnf.chain.1.sqrt_end_time']
        times = torch.stack([torch.tensor(0.0).to(x), s * s]).to(x)
        func = _CnfFunc(self.sd, e)
        out = odeint001.odeint_adjoint(func, (x, logpx, context), times, atol=[1e-5] * 3, rtol=[1e-5] * 3,
                                       method='dopri5', options={})
        self.nfe[1] = func.nfe
        self._cnf_func = func
        return out[0][1], out[1][1]

    def point_cnf_train(self, x, context, logpx, e):
        """cnf.py:33-48, forward order [MBN, CNF, MBN]; running statistics are used as they are
        (normalization.py:60-61 takes the PRE-update values), their refresh is not modelled here."""
        x, logpx = self._mbn(0, x, logpx, False)
        x, logpx = self.cnf_train(x, context, logpx, e)
        return self._mbn(2, x, logpx, False)

    def forward_train(self, x, sample_points, e):
        """caspr.py:76-146 -> (nll (B,T,N), tnocs_l1 (B,T,N,4)), differentiable."""
        z0, tnocs = self.encode(x)
        B, T, N, _ = sample_points.shape
        tnocs_loss = (tnocs[..., :4] - sample_points[..., :4]).abs() if self.regress_tnocs else None
        z = self.aggregate_and_solve_latent(z0, sample_points[:, :, 0, 3])
        pts = sample_points.reshape(B * T, N, 4)[:, :, :3].clone()
        yy, dlogp = self.point_cnf_train(pts, z.reshape(B * T, -1), torch.zeros(B * T, N, 1), e)
        log_py = self.standard_normal_logprob(yy).sum(2)
        nll = -(log_py - dlogp.view(B * T, N))
        return nll.view(B, T, -1), tnocs_loss

    @staticmethod
    def loss(nll, tnocs_l1, cnf_loss_weight=0.01, tnocs_loss_weight=100.0):
        """train_utils.py:148-166 with the default weights of config_utils.py:42-43."""
        total = cnf_loss_weight * nll.sum(2).mean()
        if tnocs_l1 is not None:
            total = total + tnocs_loss_weight * tnocs_l1[:, :, :, :4].mean()
        return total
