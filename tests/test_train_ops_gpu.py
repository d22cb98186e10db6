"""GPU parity of the training operators (caspr_b200/train_ops.py) against torch fp32/fp64 autograd of the same
operator (floating-point kernels: tolerance written per test)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.fixture(scope='module')
def tops():
    from caspr_b200 import train_ops
    return train_ops


def _gn_ref(x, samples, rps, gamma, beta, relu):
    C = x.shape[1]
    y = F.group_norm(x.view(samples, rps, C).transpose(1, 2), 16, gamma, beta, eps=1e-5)
    y = F.relu(y) if relu else y
    return y.transpose(1, 2).reshape(samples * rps, C)


@pytest.mark.parametrize('samples,rps,C', [(3, 700, 64), (257, 16, 16), (130, 32, 96), (2, 2048, 1600), (64, 32, 512)])
@pytest.mark.parametrize('relu', [False, True])
def test_groupnorm_forward_backward(tops, samples, rps, C, relu):
    g = torch.Generator().manual_seed(samples + C)
    x = (torch.randn(samples * rps, C + 5, generator=g) * 2 + 0.5).cuda()[:, 2:2 + C]       # strided view
    if rps <= 32:
        # balls padded with copies of one point (near-zero variance): the statistics must not cancel
        xv = x.view(samples, rps, C)
        xv[::2, 1:, :] = xv[::2, :1, :] + 10.0
        xv[::2, 0, :] += 10.0
    gamma = (torch.rand(C, generator=g) + 0.5).cuda()
    beta = torch.randn(C, generator=g).cuda()
    d_out = torch.randn(samples * rps, C, generator=g).cuda()
    d_max = torch.randn(samples, C, generator=g).cuda()
    mr = tops.gn_moments(x, samples, rps, 16)
    y = tops.gn_apply(x, mr, samples, rps, 16, gamma, beta, relu)
    xd = x.double().contiguous().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    y_ref = _gn_ref(xd, samples, rps, gd, bd, relu)
    assert _rel(y, y_ref.detach()) < 2e-5
    mx, arg = tops.rowmax(y, samples, rps)
    mx_ref, arg_ref = y.view(samples, rps, C).max(1)
    assert torch.equal(mx, mx_ref)
    # ties (ReLU zeros) may resolve differently: compare the values the indices point at
    assert torch.equal(torch.gather(y.view(samples, rps, C), 1, arg.long().unsqueeze(1)).squeeze(1), mx_ref)
    # backward with both a dense and a max-pool cotangent
    pooled = torch.gather(y_ref.view(samples, rps, C), 1, arg.long().unsqueeze(1)).squeeze(1)
    ((y_ref * d_out.double()).sum() + (pooled * d_max.double()).sum()).backward()
    dx, dgamma, dbeta = tops.gn_backward(x, mr, samples, rps, 16, gamma, beta, relu, d_out=d_out, d_max=d_max, argmax=arg)
    assert _rel(dx, xd.grad) < 2e-4
    assert _rel(dgamma, gd.grad) < 2e-4
    assert _rel(dbeta, bd.grad) < 2e-4


@pytest.mark.parametrize('rows,cin,cout', [(5000, 9, 16), (4096, 99, 32), (777, 515, 256), (20480, 1600, 1600), (3000, 64, 4), (81920, 512, 512)])
def test_linear_wgrad(tops, rows, cin, cout):
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, cin + 3, generator=g).cuda()[:, 1:1 + cin]
    dy = torch.randn(rows, cout, generator=g).cuda()
    ref = dy.double().t() @ x.double()
    for engine, tol in (('simt', 1e-5), ('tc', 3e-5)):
        dW, db = tops.linear_wgrad(dy, x, engine=engine)
        assert _rel(dW, ref) < tol, engine
        assert _rel(db, dy.double().sum(0)) < 1e-5
        dW2, _ = tops.linear_wgrad(dy, x, relu_x=True, engine=engine)
        assert _rel(dW2, dy.double().t() @ x.double().clamp_min(0)) < tol, engine


def test_gather_backwards(tops):
    from caspr_b200 import ops
    g = torch.Generator().manual_seed(0)
    B, N, M, C, ns = 3, 200, 50, 37, 16
    xyz = torch.rand(B, N, 3, generator=g).cuda()
    idx_f, new_xyz = ops.fps(xyz, M)
    bq, _ = ops.ball_query2(xyz, new_xyz, 0.3, ns, 0.5, 32)
    feat = torch.randn(B, N, C, generator=g).cuda()
    rows = ops.group_points(xyz, new_xyz, feat, bq)
    d_rows = torch.randn_like(rows)
    d_feat = torch.zeros(B, N, C, device='cuda')
    tops.group_points_bwd(d_rows, bq, N, C, d_feat)
    fr = feat.double().requires_grad_(True)
    gathered = torch.gather(fr.unsqueeze(1).expand(B, M, N, C), 2, bq.long().unsqueeze(-1).expand(B, M, ns, C))
    (gathered.reshape(-1, C) * d_rows[:, 3:].double()).sum().backward()
    assert _rel(d_feat, fr.grad) < 1e-5
    # three_interpolate
    n, m, Cp = 120, 40, 29
    unknown, known = torch.rand(B, n, 3, generator=g).cuda(), torch.rand(B, m, 3, generator=g).cuda()
    dist, idx = ops.three_nn(unknown, known)
    prev = torch.randn(B, m, Cp, generator=g).cuda()
    out = ops.three_interp_concat(prev, idx, dist, None)
    d_out = torch.randn_like(out)
    d_prev = torch.zeros(B, m, Cp, device='cuda')
    tops.three_interp_bwd(d_out, idx, dist, m, Cp, d_prev)
    pr = prev.double().requires_grad_(True)
    inv = 1.0 / (dist.double() + 1e-8)
    w = inv / inv.sum(2, keepdim=True)
    nb = torch.gather(pr.unsqueeze(1).expand(B, n, m, Cp), 2, idx.long().unsqueeze(-1).expand(B, n, 3, Cp))
    interp = (nb * w.unsqueeze(-1)).sum(2)
    assert _rel(out, interp.detach().reshape(-1, Cp)) < 1e-5
    (interp.reshape(-1, Cp) * d_out.double()).sum().backward()
    assert _rel(d_prev, pr.grad) < 1e-5


def test_small_helpers(tops):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(5000, 70, generator=g).cuda()
    out = torch.zeros(64, device='cuda')
    tops.colsum(x[:, 3:67], out)
    assert _rel(out, x[:, 3:67].double().sum(0)) < 1e-5
    tops.colsum(x[:, 3:67], out, accumulate=True)
    assert _rel(out, 2 * x[:, 3:67].double().sum(0)) < 1e-5
    w = torch.randn(37, 91, generator=g).cuda()
    assert torch.equal(tops.transpose(w), w.t().contiguous())
    a, ref = torch.randn(100, 33, generator=g).cuda(), torch.randn(100, 33, generator=g).cuda()
    dst = torch.ones(100, 40, device='cuda')
    tops.rows_update(a, dst[:, 2:35], accumulate=True, relu_ref=ref)
    assert torch.equal(dst[:, 2:35], 1 + a * (ref > 0))
