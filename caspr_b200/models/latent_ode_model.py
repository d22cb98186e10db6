"""Latent ODE (mirror of reference caspr/models/latent_ode_model.py:11-156).

``LatentODE`` / ``ODESolver`` / ``DynamicsNet`` keep the reference's constructor arguments and
state_dict keys (``ode_func.dynamics_net.{0,2,4,6}``, the aliased ``solver.ode_func.*`` and the
``_num_evals`` buffers).  The whole adaptive dopri5 solve (torchdiffeq 0.0.1 semantics) runs in
one persistent CUDA kernel: ``caspr_latent_ode_solve``.
"""
import torch
import torch.nn as nn

from .. import ops
from .._lib import CasprError


class LatentODE(nn.Module):
    def __init__(self, input_size=1024, hidden_size=1024, num_layers=2, nonlinearity=nn.Tanh, augment_size=0):
        super(LatentODE, self).__init__()
        if nonlinearity is not nn.Tanh or num_layers != 2:
            raise NotImplementedError('the fused latent-ODE kernel implements the reference configuration: '
                                      '2 hidden layers, Tanh (caspr.py:59-62)')
        self.input_size = input_size
        self.augment_size = augment_size
        self.output_size = input_size + augment_size
        self.ode_func = DynamicsNet(input_size=self.output_size, hidden_size=hidden_size,
                                    num_layers=num_layers, nonlinearity=nonlinearity)
        self.solver = ODESolver(self.ode_func, method='dopri5', rtol=1e-3, atol=1e-4)
        init_network_weights(self.ode_func)

    def get_output_size(self):
        return self.output_size

    def forward(self, z0, t):
        """z0 (B,H), t (T,) increasing -> (B,T,H)   (latent_ode_model.py:45-70)."""
        self.ode_func._num_evals.fill_(0)
        rel_t = t - t[0]
        aug_z0 = z0
        if self.augment_size > 0:
            aug_z0 = torch.cat([z0, torch.zeros(z0.shape[0], self.augment_size, dtype=z0.dtype, device=z0.device)], 1)
        pred_z = self.solver(aug_z0, rel_t)
        return pred_z.permute(1, 0, 2)

    def num_evals(self):
        return self.ode_func._num_evals.item()


class ODESolver(nn.Module):
    def __init__(self, ode_func, method='dopri5', rtol=1e-4, atol=1e-5):
        super(ODESolver, self).__init__()
        if method != 'dopri5':
            raise NotImplementedError('only dopri5 is implemented')
        self.method = method
        self.ode_func = ode_func
        self.rtol = rtol
        self.atol = rtol        # reference quirk kept on purpose: latent_ode_model.py:83 `self.atol = rtol`

    def forward(self, z0, t):
        """-> (T,B,H), as torchdiffeq.odeint_adjoint(ode_func, z0, t, rtol, atol, 'dopri5') (:98)."""
        net = self.ode_func.dynamics_net
        lin = [net[0], net[2], net[4], net[6]]
        times = t.detach().to(torch.float32).cpu().tolist()        # float32 grid, widened to float64 in the solver
        if self.training and torch.is_grad_enabled():
            # training: same forward solve, adjoint backward in CUDA (caspr_latent_ode_adjoint)
            params = [p for l in lin for p in (l.weight, l.bias)]
            return _LatentSolveFunction.apply(z0.to(torch.float32), self, times, *params)
        out, info, rc = ops.latent_ode_solve(z0.to(torch.float32), [l.weight for l in lin], [l.bias for l in lin],
                                             times, self.rtol, self.atol)
        self.ode_func._num_evals += float(info[1])
        if rc != 0:
            raise CasprError(rc, 'caspr_latent_ode_solve')
        return out


class _LatentSolveFunction(torch.autograd.Function):
    """odeint_adjoint for the latent ODE (latent_ode_model.py:98): forward = caspr_latent_ode_solve, backward =
    caspr_latent_ode_adjoint (torchdiffeq 0.0.1's OdeintAdjointMethod.backward)."""

    @staticmethod
    def forward(ctx, z0, solver, times, *params):
        out, info, rc = ops.latent_ode_solve(z0, [p.detach() for p in params[0::2]],
                                             [p.detach() for p in params[1::2]], times, solver.rtol, solver.atol)
        solver.ode_func._num_evals += float(info[1])
        if rc != 0:
            raise CasprError(rc, 'caspr_latent_ode_solve')
        ctx.solver, ctx.times, ctx.params = solver, times, params
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g_out):
        out, = ctx.saved_tensors
        params, solver = ctx.params, ctx.solver
        gz0, gpar, info, rc = ops.latent_ode_adjoint(out, g_out.contiguous(), [p.detach() for p in params[0::2]],
                                                     [p.detach() for p in params[1::2]], ctx.times, solver.rtol,
                                                     solver.atol)
        solver.last_adjoint_info = info
        if rc != 0:
            raise CasprError(rc, 'caspr_latent_ode_adjoint')
        grads, off = [], 0
        for p in params:
            grads.append(gpar[off:off + p.numel()].view_as(p))
            off += p.numel()
        return (gz0, None, None) + tuple(grads)


class DynamicsNet(nn.Module):
    def __init__(self, input_size=1024, hidden_size=1024, num_layers=2, nonlinearity=nn.Tanh):
        super(DynamicsNet, self).__init__()
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.num_layers = num_layers
        self.nonlinearity = nonlinearity
        self.register_buffer('_num_evals', torch.tensor(0.))
        layers = [nn.Linear(input_size, hidden_size), nonlinearity()]
        for _ in range(num_layers):
            layers += [nn.Linear(hidden_size, hidden_size), nonlinearity()]
        layers.append(nn.Linear(hidden_size, input_size))
        self.dynamics_net = nn.Sequential(*layers)

    def forward(self, t, z):
        raise RuntimeError('DynamicsNet is evaluated inside caspr_latent_ode_solve; there is no '
                           'PyTorch fallback path')


def init_network_weights(net, std=0.1):
    """latent_ode_model.py:152-156."""
    for m in net.modules():
        if isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, mean=0, std=std)
            nn.init.constant_(m.bias, val=0)
