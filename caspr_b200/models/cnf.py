"""Conditional continuous normalizing flow (PointFlow-style) on the fused CUDA solver.

Mirrors the module tree of reference caspr/models/cnf.py (``SequentialFlow`` :20-48, ``CNF``
:50-131), odefunc.py (``ODEnet`` :62-105, ``ODEfunc`` :108-142), diffeq_layers.py
(``ConcatSquashLinear`` :76-90) and normalization.py (``MovingBatchNorm1d`` :12-127) so the
state_dict keys ``chain.{0,2}.{weight,bias,step,running_mean,running_var}``,
``chain.1.sqrt_end_time``, ``chain.1.odefunc._num_evals`` and
``chain.1.odefunc.diffeq.layers.{l}.{_layer,_hyper_bias,_hyper_gate}.*`` load unchanged.

The modules are parameter containers: ``SequentialFlow.forward`` hands the whole chain
[MovingBatchNorm, CNF, MovingBatchNorm] to ``caspr_cnf_flow`` — one device-resident dopri5
solve that evaluates the dynamics MLP, its Hutchinson divergence and the step controller in
CUDA (the host only polls the 8-word solver state once per attempted step; no autograd VJP).
"""
import torch
import torch.nn as nn

from .. import ops
from .._lib import CasprError

__all__ = ['CNF', 'SequentialFlow', 'ODEnet', 'ODEfunc', 'ConcatSquashLinear', 'MovingBatchNorm1d']


class ConcatSquashLinear(nn.Module):
    """diffeq_layers.py:76-90: (W x + b) * sigmoid(Wg [t,c] + bg) + Wb [t,c]."""

    def __init__(self, dim_in, dim_out, dim_c):
        super(ConcatSquashLinear, self).__init__()
        self._layer = nn.Linear(dim_in, dim_out)
        self._hyper_bias = nn.Linear(1 + dim_c, dim_out, bias=False)
        self._hyper_gate = nn.Linear(1 + dim_c, dim_out)


class ODEnet(nn.Module):
    def __init__(self, hidden_dims, input_shape, context_dim, layer_type='concatsquash', nonlinearity='softplus'):
        super(ODEnet, self).__init__()
        if layer_type != 'concatsquash' or nonlinearity != 'softplus':
            raise NotImplementedError('the fused CNF kernels implement the reference configuration '
                                      '(concatsquash + softplus, flow.py:86-100)')
        dims = [input_shape[0]] + list(hidden_dims) + [input_shape[0]]
        self.layers = nn.ModuleList([ConcatSquashLinear(dims[i], dims[i + 1], context_dim)
                                     for i in range(len(dims) - 1)])
        self.activation_fns = nn.ModuleList([nn.Softplus() for _ in range(len(dims) - 2)])


class ODEfunc(nn.Module):
    def __init__(self, diffeq):
        super(ODEfunc, self).__init__()
        self.diffeq = diffeq
        self.register_buffer('_num_evals', torch.tensor(0.))
        self._e = None

    def before_odeint(self, e=None):
        """odefunc.py:115-117.  Unlike the reference, a Hutchinson noise tensor passed here is used."""
        self._e = e
        self._num_evals.fill_(0)


class MovingBatchNorm1d(nn.Module):
    """normalization.py:12-127 (eps 1e-4, decay 0.1, zero-initialised affine)."""

    def __init__(self, num_features, eps=1e-4, decay=0.1, bn_lag=0., affine=True):
        super(MovingBatchNorm1d, self).__init__()
        self.num_features = num_features
        self.affine = affine
        self.eps = eps
        self.decay = decay
        self.bn_lag = bn_lag
        self.register_buffer('step', torch.zeros(1))
        if affine:
            self.weight = nn.Parameter(torch.zeros(num_features))
            self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer('running_mean', torch.zeros(num_features))
        self.register_buffer('running_var', torch.ones(num_features))

    def update_running_mean(self, x):
        """normalization.py:43-51, including its transpose/reshape quirk."""
        nc = x.size(-1)
        x_t = x.transpose(0, 1).reshape(nc, -1)
        self.running_mean -= self.decay * (self.running_mean - torch.mean(x_t, dim=1).data)
        self.running_var -= self.decay * (self.running_var - torch.var(x_t, dim=1).data)
        self.step += 1

    def forward(self, x, c=None, logpx=None, reverse=False):
        """Stand-alone (unfused) evaluation, used only for the training-mode tail of the chain where the
        running statistics are refreshed from the layer input (normalization.py:59-101).  In eval mode
        the layer is folded into caspr_cnf_flow."""
        used_mean = self.running_mean.clone().detach()
        used_var = self.running_var.clone().detach()
        if self.training and not reverse:
            self.update_running_mean(x)
        w = self.weight if self.affine else torch.zeros_like(used_mean)
        b = self.bias if self.affine else torch.zeros_like(used_mean)
        logdet = (-0.5 * torch.log(used_var + self.eps) + w).sum()
        if not reverse:
            y = (x - used_mean) * torch.exp(-0.5 * torch.log(used_var + self.eps))
            y = y * torch.exp(w) + b
            return y if logpx is None else (y, logpx - logdet)
        y = (x - b) * torch.exp(-w)
        y = y * torch.exp(0.5 * torch.log(used_var + self.eps)) + used_mean
        return y if logpx is None else (y, logpx + logdet)

    def params(self):
        z = torch.zeros_like(self.running_mean)
        return {'weight': self.weight.detach() if self.affine else z,
                'bias': self.bias.detach() if self.affine else z,
                'running_mean': self.running_mean.clone(), 'running_var': self.running_var.clone()}


class CNF(nn.Module):
    def __init__(self, odefunc, conditional=True, T=1.0, train_T=False, solver='dopri5', atol=1e-5, rtol=1e-5,
                 use_adjoint=True):
        super(CNF, self).__init__()
        if solver != 'dopri5' or not conditional:
            raise NotImplementedError('caspr_cnf_flow implements the conditional dopri5 flow')
        self.train_T = train_T
        self.T = T
        if train_T:
            self.register_parameter('sqrt_end_time', nn.Parameter(torch.sqrt(torch.tensor(T))))
        self.use_adjoint = use_adjoint
        self.odefunc = odefunc
        self.solver = solver
        self.atol = atol
        self.rtol = rtol
        self.test_solver = solver
        self.test_atol = atol
        self.test_rtol = rtol
        self.solver_options = {}
        self.conditional = conditional
        self._pack = None
        self._pack_key = None

    def end_time(self):
        """cnf.py:89-91: sqrt_end_time^2 formed in fp32."""
        if self.train_T:
            s = self.sqrt_end_time.detach().to(torch.float32)
            return float((s * s).item())
        return float(torch.tensor(self.T, dtype=torch.float32).item())

    def weight_pack(self):
        """caspr_cnf_weights view of the ODEnet parameters (rebuilt when they move or change storage)."""
        layers = self.odefunc.diffeq.layers
        if len(layers) != 4:
            raise NotImplementedError('fused CNF kernels expect 3 hidden layers (dims "512-512-512")')
        tensors = []
        for l in layers:
            tensors.append({'W': l._layer.weight.detach(), 'b': l._layer.bias.detach(),
                            'Wgate': l._hyper_gate.weight.detach(), 'bgate': l._hyper_gate.bias.detach(),
                            'Wbias': l._hyper_bias.weight.detach()})
        key = tuple(t.data_ptr() for d in tensors for t in d.values())
        if key != self._pack_key:
            hidden = layers[0]._layer.weight.shape[0]
            ctx_dim = layers[0]._hyper_gate.weight.shape[1] - 1
            for d in tensors:
                for k in d:
                    # the pack holds POINTERS to the parameters: a converted copy would go stale after optimizer steps
                    if d[k].dtype != torch.float32 or not d[k].is_contiguous():
                        raise TypeError('the CNF kernels read the ODEnet parameters in place: they must be contiguous '
                                        'float32 tensors (got %s, contiguous=%s)' % (d[k].dtype, d[k].is_contiguous()))
            self._pack = ops.CnfWeightPack(tensors, hidden, ctx_dim)
            self._pack_key = key
        return self._pack

    def num_evals(self):
        return self.odefunc._num_evals.item()


class _CnfBlockFunction(torch.autograd.Function):
    """odeint_adjoint for one CNF block in the forward direction (cnf.py:101-111): forward = caspr_cnf_flow without
    the MovingBatchNorm layers, backward = caspr_cnf_adjoint (torchdiffeq 0.0.1's OdeintAdjointMethod.backward with the
    VJP through ODEfunc, odefunc.py:119-142)."""

    @staticmethod
    def forward(ctx, x, logp, context, e, cnf, engine, sqrt_end_time, *params):
        pack = cnf.weight_pack()
        end_time = cnf.end_time()
        x1, logp1, info, rc = ops.cnf_flow(x, logp, e, context, pack, None, None, end_time, False, cnf.rtol, cnf.atol,
                                           engine)
        cnf.odefunc._num_evals += float(info[1])
        cnf.last_info = info
        if rc != 0:
            raise CasprError(rc, 'caspr_cnf_flow')
        ctx.cnf, ctx.pack, ctx.end_time, ctx.params, ctx.engine = cnf, pack, end_time, params, engine
        ctx.save_for_backward(x1, logp1, e, context, sqrt_end_time)
        return x1, logp1

    @staticmethod
    def backward(ctx, gx1, glogp1):
        x1, logp1, e, context, sqrt_end_time = ctx.saved_tensors
        cnf = ctx.cnf
        gx0, glogp0, gctx, gpar, gtimes, info, rc = ops.cnf_adjoint(
            x1, logp1, gx1.contiguous(), glogp1.contiguous(), e, context, ctx.pack, ctx.end_time, cnf.rtol, cnf.atol,
            engine=ctx.engine)
        cnf.last_adjoint_info = info
        if rc != 0:
            raise CasprError(rc, 'caspr_cnf_adjoint')
        grads, off = [], 0
        for p in ctx.params:
            grads.append(gpar[off:off + p.numel()].view_as(p))
            off += p.numel()
        # integration_times = [0, sqrt_end_time^2] (cnf.py:89-91)
        g_sqrt = (2.0 * sqrt_end_time.detach() * gtimes[1]).reshape(sqrt_end_time.shape) if cnf.train_T else None
        return (gx0, glogp0, gctx, None, None, None, g_sqrt) + tuple(grads)


class SequentialFlow(nn.Module):
    """chain = [MovingBatchNorm1d, CNF x num_blocks, MovingBatchNorm1d] (flow.py:67-74)."""

    engine = ops.CNF_TC_FP16X3      # tcgen05 fp16x3 engine; ops.CNF_SIMT_FP32 is the exact-fp32 SIMT engine
    lockstep_group = False          # set by sharding.lockstep(): process group whose ranks share one step sequence

    def __init__(self, layer_list, use_bn=True):
        super(SequentialFlow, self).__init__()
        self.chain = nn.ModuleList(layer_list)
        self.use_bn = use_bn

    def forward(self, x, context, logpx=None, reverse=False, inds=None, integration_times=None, e=None):
        """x (F,P,3), context (F,ctx), logpx (F,P,1) or None -> x' or (x', logpx') as cnf.py:33-48.

        ``e``: optional Hutchinson noise (F,P,3); the reference draws it with torch.randn_like on the
        device once per solve (odefunc.py:127-128)."""
        if inds is not None or integration_times is not None:
            raise NotImplementedError('custom chain indices / integration times are not used by CaSPR')
        if not x.is_cuda:
            raise RuntimeError('caspr_b200 runs on CUDA only (no CPU fallback)')
        F, P, _ = x.shape
        x = x.to(torch.float32)
        context = context.reshape(F, -1).to(torch.float32)
        logp = None if logpx is None else logpx.reshape(F, P).to(torch.float32)
        mods = list(self.chain)
        cnfs = [m for m in mods if isinstance(m, CNF)]
        first_bn = mods[0] if isinstance(mods[0], MovingBatchNorm1d) else None
        last_bn = mods[-1] if isinstance(mods[-1], MovingBatchNorm1d) else None
        train_fwd = self.training and not reverse
        if train_fwd and torch.is_grad_enabled():
            return self._forward_train(x, context, logp, mods, e, logpx is not None)
        order = list(reversed(cnfs)) if reverse else cnfs
        for i, cnf in enumerate(order):
            is_first, is_last = i == 0, i == len(order) - 1
            if reverse:
                pre_bn, post_bn = (last_bn if is_first else None), (first_bn if is_last else None)
            else:
                pre_bn, post_bn = (first_bn if is_first else None), (last_bn if is_last else None)
            # statistics used by this pass are the PRE-update ones (normalization.py:60-64)
            p_pre = pre_bn.params() if pre_bn is not None else None
            if train_fwd and pre_bn is not None:
                pre_bn.update_running_mean(x)
            fuse_post = post_bn is not None and not train_fwd
            p_post = post_bn.params() if fuse_post else None
            mbn0, mbn2 = (p_post, p_pre) if reverse else (p_pre, p_post)
            cnf.odefunc.before_odeint(e)
            noise = e if e is not None else torch.randn_like(x)
            rtol, atol = (cnf.rtol, cnf.atol) if self.training else (cnf.test_rtol, cnf.test_atol)
            sync = None
            if self.lockstep_group is not False:
                import torch.distributed as dist
                n_all = torch.tensor([F * P], dtype=torch.int64, device=x.device)
                dist.all_reduce(n_all, group=self.lockstep_group)
                sync = ops.LockstepSync(int(n_all.item()), x.device, self.lockstep_group)
            x_out, logp_out, info, rc = ops.cnf_flow(x, logp, noise, context, cnf.weight_pack(), mbn0, mbn2,
                                                     cnf.end_time(), reverse, rtol, atol, self.engine, sync=sync)
            cnf.odefunc._num_evals += float(info[1])
            self.last_info = info
            if rc != 0:
                raise CasprError(rc, 'caspr_cnf_flow')
            x, logp = x_out, logp_out
            if post_bn is not None and not fuse_post:
                # training: the trailing layer refreshes its statistics from the CNF output
                if logp is None:
                    x = post_bn(x)
                else:
                    x, lp = post_bn(x, None, logp.unsqueeze(-1))
                    logp = lp.squeeze(-1)
        if logpx is None:
            return x
        return x, logp.view(F, P, 1)

    def _forward_train(self, x, context, logp, mods, e, want_logp):
        """Differentiable forward pass of the chain (training, cnf.py:33-48 in chain order): the MovingBatchNorm layers
        are elementwise torch code, every CNF block is a ``_CnfBlockFunction`` (CUDA solve + CUDA adjoint)."""
        F, P, _ = x.shape
        if logp is None:
            logp = torch.zeros(F, P, dtype=torch.float32, device=x.device)            # cnf.py:71-74
        for m in mods:
            if isinstance(m, MovingBatchNorm1d):
                x, lp = m(x, None, logp.unsqueeze(-1))
                logp = lp.squeeze(-1)
                continue
            m.odefunc.before_odeint(e)
            noise = e if e is not None else torch.randn_like(x)
            params = []
            for l in m.odefunc.diffeq.layers:       # ODEfunc.parameters() order (diffeq_layers.py:79-81)
                params += [l._layer.weight, l._layer.bias, l._hyper_bias.weight, l._hyper_gate.weight,
                           l._hyper_gate.bias]
            s = m.sqrt_end_time if m.train_T else torch.zeros((), device=x.device)
            x, logp = _CnfBlockFunction.apply(x.contiguous(), logp.contiguous(), context, noise.detach(), m,
                                              self.engine, s, *params)
        if not want_logp:
            return x
        return x, logp.view(F, P, 1)
