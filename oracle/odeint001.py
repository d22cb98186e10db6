"""Oracle restatement of torchdiffeq == 0.0.1 (`odeint`, `odeint_adjoint`, dopri5).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

torchdiffeq is NOT under ``/root/reference``; the reference pins it with
"torchdiffeq==0.0.1 (note the version is important)" (README.md:22) because it
passes list-valued tolerances (``cnf.py:80-81,106-107``) that later versions
reject.  Reference call sites: ``latent_ode_model.py:98`` (tensor state,
rtol=atol=1e-3, increasing t) and ``cnf.py:102-119`` (3-tuple state
``(x, logp, context)``, list tolerances ``[1e-5]*3``, two time points, possibly
decreasing).  What is restated here is the published algorithm of that release:

* tensor y0 is wrapped into a 1-tuple; strictly decreasing ``t`` integrates
  ``-f(-t, y)`` over ``-t``;
* time grid, ``t0``, ``t1`` and ``dt`` bookkeeping are float64; every quantity
  that meets the state (stage times, ``dt`` inside the RK combination, the
  interpolation abscissa) is first cast to the state dtype;
* initial step from the Hairer/Norsett/Wanner heuristic of order 4 using
  ``rtol[0]``/``atol[0]`` for EVERY state tensor — 2 function evaluations
  before the first step (``f(t0,y0)`` and ``f(t0+h0, y0+h0 f0)``);
* Dormand-Prince 5(4) with FSAL: 6 new evaluations per attempted step;
* error ratio per state tensor = mean over ALL its elements of
  ``(err / (atol + rtol*max(|y0|,|y1|)))**2``; accept iff every ratio <= 1;
  tolerance lists are zipped with the state tuple, so surplus state tensors
  take no part in step control;
* ``dt_next = dt / max(1/ifactor, min(ratio**(1/10)/safety, 1/dfactor))`` with
  safety .9, ifactor 10, dfactor .2 (dfactor := 1 when ratio < 1; ``dt*10``
  when ratio == 0); rejected steps retry from the same (y0, f0, t0);
* steps are never clipped to the output times: the loop runs
  ``while t_out > t1`` and the output is the quartic dense-output polynomial of
  the step that overshoots;
* ``odeint_adjoint``: forward under ``no_grad``; backward integrates
  ``(y, adj_y, adj_t, adj_params)`` from ``t[i]`` to ``t[i-1]`` with the same
  solver and the SAME (possibly shorter) tolerance lists.

Quirk kept on purpose: with a state tensor whose derivative is identically
zero (the CNF's context, ``odefunc.py:136``) the heuristic's
``h0 = 0.01*max(d0/d1)`` is +inf, the probe evaluation ``f(t0+h0, y0+h0*f0)``
is garbage (NaN), Python's ``max`` skips the NaNs, and the chosen first step is
``(0.01/max(d1))**(1/5)``.  The evaluation still counts towards NFE.
"""
import math

import torch
import torch.nn as nn

# Dormand-Prince 5(4) tableau (Shampine's dense-output variant)
DP_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1., 1.]
DP_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
DP_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
DP_C_ERROR = [
    35 / 384 - 1951 / 21600,
    0,
    500 / 1113 - 22642 / 50085,
    125 / 192 - 451 / 720,
    -2187 / 6784 - -12231 / 42400,
    11 / 84 - 649 / 6300,
    -1. / 60.,
]
DP_C_MID = [
    6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
    187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2
]


def _is_iterable(x):
    try:
        iter(x)
        return True
    except TypeError:
        return False


def _rms(x):
    return x.norm() / (x.numel() ** 0.5)


def _scaled_dot(scale, coeffs, tensors):
    """sum_i (scale*c_i)*k_i, left to right starting from integer 0, every term kept."""
    return sum([(scale * c) * k for c, k in zip(coeffs, tensors)])


def _dot(coeffs, tensors):
    return sum([c * k for c, k in zip(coeffs, tensors)])


def _select_initial_step(fun, t0, y0, order, rtol, atol, f0):
    t0 = t0.to(y0[0])
    rtol = rtol if _is_iterable(rtol) else [rtol] * len(y0)
    atol = atol if _is_iterable(atol) else [atol] * len(y0)
    scale = tuple(a + torch.abs(y) * r for y, a, r in zip(y0, atol, rtol))
    d0 = tuple(_rms(y / s) for y, s in zip(y0, scale))
    d1 = tuple(_rms(f / s) for f, s in zip(f0, scale))
    if max(d0).item() < 1e-5 or max(d1).item() < 1e-5:
        h0 = torch.tensor(1e-6).to(t0)
    else:
        h0 = 0.01 * max(a / b for a, b in zip(d0, d1))
    y1 = tuple(y + h0 * f for y, f in zip(y0, f0))
    f1 = fun(t0 + h0, y1)
    d2 = tuple(_rms((b - a) / s) / h0 for b, a, s in zip(f1, f0, scale))
    if max(d1).item() <= 1e-15 and max(d2).item() <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6).to(h0), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1 + d2)) ** (1. / float(order + 1))
    return torch.min(100 * h0, h1)


def _error_ratio(err, rtol, atol, y0, y1):
    tol = tuple(a + r * torch.max(torch.abs(p), torch.abs(q))
                for a, r, p, q in zip(atol, rtol, y0, y1))
    ratio = tuple(e / t for e, t in zip(err, tol))
    return tuple(torch.mean(r * r) for r in ratio)


def _optimal_step_size(last_step, mean_error_ratio, safety, ifactor, dfactor, order=5):
    mean_error_ratio = max(mean_error_ratio)
    if mean_error_ratio == 0:
        return last_step * ifactor
    if mean_error_ratio < 1:
        dfactor = torch.tensor(1., dtype=torch.float64)
    error_ratio = torch.sqrt(mean_error_ratio).to(last_step)
    exponent = torch.tensor(1 / order).to(last_step)
    factor = torch.max(1 / ifactor, torch.min(error_ratio ** exponent / safety, 1 / dfactor))
    return last_step / factor


class _State(object):
    __slots__ = ('y1', 'f1', 't0', 't1', 'dt', 'interp')

    def __init__(self, y1, f1, t0, t1, dt, interp):
        self.y1, self.f1, self.t0, self.t1, self.dt, self.interp = y1, f1, t0, t1, dt, interp


class Dopri5(object):
    """Adaptive dopri5 exactly as torchdiffeq 0.0.1 steps it; records a step log."""

    def __init__(self, func, y0, rtol, atol, safety=0.9, ifactor=10.0, dfactor=0.2,
                 max_num_steps=2 ** 31 - 1):
        self.func = func
        self.y0 = y0
        self.rtol = rtol if _is_iterable(rtol) else [rtol] * len(y0)
        self.atol = atol if _is_iterable(atol) else [atol] * len(y0)
        self.safety = torch.tensor(safety, dtype=torch.float64)
        self.ifactor = torch.tensor(ifactor, dtype=torch.float64)
        self.dfactor = torch.tensor(dfactor, dtype=torch.float64)
        self.max_num_steps = max_num_steps
        self.log = []          # (t0, dt, accepted, max_ratio) per attempted step

    def integrate(self, t):
        assert (t[1:] > t[:-1]).all(), 't must be strictly increasing or decreasing'
        solution = [self.y0]
        t = t.to(torch.float64)
        f0 = self.func(t[0].type_as(self.y0[0]), self.y0)
        first = _select_initial_step(self.func, t[0], self.y0, 4, self.rtol[0], self.atol[0], f0=f0).to(t)
        self.state = _State(self.y0, f0, t[0], t[0], first, [self.y0] * 5)
        self.first_step = float(first)
        for i in range(1, len(t)):
            n = 0
            while t[i] > self.state.t1:
                assert n < self.max_num_steps, 'max_num_steps exceeded'
                self.state = self._step(self.state)
                n += 1
            solution.append(self._interp_eval(self.state, t[i]))
        return tuple(map(torch.stack, tuple(zip(*solution))))

    def _step(self, st):
        y0, f0, t0, dt = st.y1, st.f1, st.t1, st.dt
        assert t0 + dt > t0, 'underflow in dt {}'.format(dt.item())
        for y in y0:
            assert torch.isfinite(y).all(), 'non-finite values in state `y`: {}'.format(y)
        dtype = y0[0].dtype
        t0_s = t0.to(dtype)
        dt_s = dt.to(dtype)
        k = tuple([f] for f in f0)
        yi = y0
        for alpha_i, beta_i in zip(DP_ALPHA, DP_BETA):
            ti = t0_s + alpha_i * dt_s
            yi = tuple(y + _scaled_dot(dt_s, beta_i, k_) for y, k_ in zip(y0, k))
            for k_, f_ in zip(k, self.func(ti, yi)):
                k_.append(f_)
        y1 = yi                                   # c_sol == beta[-1] (+0): saved combination
        f1 = tuple(k_[-1] for k_ in k)
        err = tuple(_scaled_dot(dt_s, DP_C_ERROR, k_) for k_ in k)
        ratio = _error_ratio(err, self.rtol, self.atol, y0, y1)
        accept = bool((torch.tensor(ratio) <= 1).all())
        self.log.append((float(t0), float(dt), accept, float(max(ratio))))
        if accept:
            y_mid = tuple(y + _scaled_dot(dt_s, DP_C_MID, k_) for y, k_ in zip(y0, k))
            interp = self._interp_fit(y0, y1, y_mid, f0, f1, dt_s)
            y_next, f_next, t_next = y1, f1, t0 + dt
        else:
            interp = st.interp
            y_next, f_next, t_next = y0, f0, t0
        dt_next = _optimal_step_size(dt, ratio, self.safety, self.ifactor, self.dfactor)
        return _State(y_next, f_next, t0, t_next, dt_next, interp)

    @staticmethod
    def _interp_fit(y0, y1, y_mid, f0, f1, dt):
        a = tuple(_dot([-2 * dt, 2 * dt, -8, -8, 16], [f0_, f1_, y0_, y1_, ym_])
                  for f0_, f1_, y0_, y1_, ym_ in zip(f0, f1, y0, y1, y_mid))
        b = tuple(_dot([5 * dt, -3 * dt, 18, 14, -32], [f0_, f1_, y0_, y1_, ym_])
                  for f0_, f1_, y0_, y1_, ym_ in zip(f0, f1, y0, y1, y_mid))
        c = tuple(_dot([-4 * dt, dt, -11, -5, 16], [f0_, f1_, y0_, y1_, ym_])
                  for f0_, f1_, y0_, y1_, ym_ in zip(f0, f1, y0, y1, y_mid))
        d = tuple(dt * f0_ for f0_ in f0)
        e = y0
        return [a, b, c, d, e]

    @staticmethod
    def _interp_eval(st, t):
        dtype = st.interp[0][0].dtype
        t0, t1, t = st.t0.to(dtype), st.t1.to(dtype), t.to(dtype)
        assert (t0 <= t) & (t <= t1), 'invalid interpolation, fails `t0 <= t <= t1`'
        x = ((t - t0) / (t1 - t0)).to(dtype)
        xs = [torch.tensor(1).to(dtype), x]
        for _ in range(2, len(st.interp)):
            xs.append(xs[-1] * x)
        return tuple(_dot(coeffs, reversed(xs)) for coeffs in zip(*st.interp))


LAST_SOLVER = [None]     # test hook: the most recent Dopri5 object (its step log)


def odeint(func, y0, t, rtol=1e-7, atol=1e-9, method=None, options=None):
    assert method in (None, 'dopri5'), 'oracle restates dopri5 only'
    tensor_input = False
    if torch.is_tensor(y0):
        tensor_input = True
        y0 = (y0,)
        _base = func
        func = lambda t, y: (_base(t, y[0]),)
    assert isinstance(y0, tuple)
    if len(t) > 1 and bool((t[1:] < t[:-1]).all()):
        t = -t
        _fwd = func
        func = lambda t, y: tuple(-f_ for f_ in _fwd(-t, y))
    for y in y0:
        assert torch.is_floating_point(y)
    solver = Dopri5(func, y0, rtol=rtol, atol=atol, **(options or {}))
    solution = solver.integrate(t)
    LAST_SOLVER[0] = solver
    if tensor_input:
        solution = solution[0]
    return solution


def _flatten(seq):
    flat = [p.contiguous().view(-1) for p in seq]
    return torch.cat(flat) if len(flat) > 0 else torch.tensor([])


def _flatten_none_to_zeros(seq, like):
    flat = [p.contiguous().view(-1) if p is not None else torch.zeros_like(q).view(-1)
            for p, q in zip(seq, like)]
    return torch.cat(flat) if len(flat) > 0 else torch.tensor([])


class _AdjointMethod(torch.autograd.Function):

    @staticmethod
    def forward(ctx, *args):
        y0, func, t, flat_params, rtol, atol, method, options = \
            args[:-7], args[-7], args[-6], args[-5], args[-4], args[-3], args[-2], args[-1]
        ctx.func, ctx.rtol, ctx.atol, ctx.method, ctx.options = func, rtol, atol, method, options
        with torch.no_grad():
            ans = odeint(func, y0, t, rtol=rtol, atol=atol, method=method, options=options)
        ctx.save_for_backward(t, flat_params, *ans)
        return ans

    @staticmethod
    def backward(ctx, *grad_output):
        t, flat_params, *ans = ctx.saved_tensors
        ans = tuple(ans)
        func, rtol, atol, method, options = ctx.func, ctx.rtol, ctx.atol, ctx.method, ctx.options
        n = len(ans)
        f_params = tuple(func.parameters())

        def augmented_dynamics(t, y_aug):
            y, adj_y = y_aug[:n], y_aug[n:2 * n]
            with torch.set_grad_enabled(True):
                t = t.to(y[0].device).detach().requires_grad_(True)
                y = tuple(y_.detach().requires_grad_(True) for y_ in y)
                func_eval = func(t, y)
                vjp_t, *vjp_y_and_params = torch.autograd.grad(
                    func_eval, (t,) + y + f_params, tuple(-a for a in adj_y),
                    allow_unused=True, retain_graph=True)
            vjp_y = vjp_y_and_params[:n]
            vjp_params = vjp_y_and_params[n:]
            vjp_t = torch.zeros_like(t) if vjp_t is None else vjp_t
            vjp_y = tuple(torch.zeros_like(y_) if v is None else v for v, y_ in zip(vjp_y, y))
            vjp_params = _flatten_none_to_zeros(vjp_params, f_params)
            if len(f_params) == 0:
                vjp_params = torch.tensor(0.).to(vjp_y[0])
            return (*func_eval, *vjp_y, vjp_t, vjp_params)

        T = ans[0].shape[0]
        with torch.no_grad():
            adj_y = tuple(g[-1] for g in grad_output)
            adj_params = torch.zeros_like(flat_params)
            adj_time = torch.tensor(0.).to(t)
            time_vjps = []
            for i in range(T - 1, 0, -1):
                ans_i = tuple(a[i] for a in ans)
                grad_i = tuple(g[i] for g in grad_output)
                func_i = func(t[i], ans_i)
                dLd_cur_t = sum(torch.dot(f_.reshape(-1), g_.reshape(-1)).reshape(1)
                                for f_, g_ in zip(func_i, grad_i))
                adj_time = adj_time - dLd_cur_t
                time_vjps.append(dLd_cur_t)
                if adj_params.numel() == 0:
                    adj_params = torch.tensor(0.).to(adj_y[0])
                aug_y0 = (*ans_i, *adj_y, adj_time, adj_params)
                aug_ans = odeint(augmented_dynamics, aug_y0, torch.tensor([t[i], t[i - 1]]),
                                 rtol=rtol, atol=atol, method=method, options=options)
                adj_y = aug_ans[n:2 * n]
                adj_time = aug_ans[2 * n]
                adj_params = aug_ans[2 * n + 1]
                adj_y = tuple(a[1] if len(a) > 0 else a for a in adj_y)
                if len(adj_time) > 0:
                    adj_time = adj_time[1]
                if len(adj_params) > 0:
                    adj_params = adj_params[1]
                adj_y = tuple(a + g[i - 1] for a, g in zip(adj_y, grad_output))
            time_vjps.append(adj_time)
            time_vjps = torch.cat(time_vjps[::-1])
            return (*adj_y, None, time_vjps, adj_params, None, None, None, None, None)


class _TupleFunc(nn.Module):
    def __init__(self, base_func):
        super().__init__()
        self.base_func = base_func

    def forward(self, t, y):
        return (self.base_func(t, y[0]),)


def odeint_adjoint(func, y0, t, rtol=1e-6, atol=1e-12, method=None, options=None):
    if not isinstance(func, nn.Module):
        raise ValueError('func is required to be an instance of nn.Module.')
    tensor_input = False
    if torch.is_tensor(y0):
        tensor_input = True
        y0 = (y0,)
        func = _TupleFunc(func)
    flat_params = _flatten(func.parameters())
    ys = _AdjointMethod.apply(*y0, func, t, flat_params, rtol, atol, method, options)
    if tensor_input:
        ys = ys[0]
    return ys
