"""Development check: caspr_cnf_adjoint against the CPU training oracle on a small problem."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from caspr_b200 import ops
from caspr_b200.models import CaSPR
from caspr_b200.synth import synthetic_state_dict
from oracle.train_oracle import TrainOracle

F_, P_ = int(os.environ.get('F', 3)), int(os.environ.get('P', 70))
sd = synthetic_state_dict(0, cnf_init='vigorous')
g = torch.Generator().manual_seed(1)
x = torch.randn(F_, P_, 3, generator=g) * 0.3
e = torch.randn(F_, P_, 3, generator=g)
ctx = torch.randn(F_, 1600, generator=g) * 0.5
gx1 = torch.randn(F_, P_, 3, generator=g)
gl1 = torch.randn(F_, P_, generator=g)

orc = TrainOracle(sd)
xo = x.clone().requires_grad_(True)
co = ctx.clone().requires_grad_(True)
lo = torch.zeros(F_, P_, 1, requires_grad=True)
t0 = time.time()
y, lp = orc.cnf_train(xo, co, lo, e)
(y * gx1).sum().add((lp.squeeze(-1) * gl1).sum()).backward()
print('oracle fwd+bwd %.1fs nfe %s' % (time.time() - t0, orc.nfe), flush=True)
log = list(__import__('oracle.odeint001', fromlist=['x']).LAST_SOLVER[0].log)
print('oracle adjoint steps', len(log), 'first dt', log[0][1])

model = CaSPR().cuda().eval()
model.load_state_dict(sd)
cnf = [m for m in model.point_cnf.chain if hasattr(m, 'odefunc')][0]
pack = cnf.weight_pack()
T = cnf.end_time()
xg, eg, cg = x.cuda(), e.cuda(), ctx.cuda()
x1, lp1, info, rc = ops.cnf_flow(xg, torch.zeros(F_, P_, device='cuda'), eg, cg, pack, None, None, T, False,
                                 1e-5, 1e-5, ops.CNF_SIMT_FP32)
print('fwd rc', rc, info, 'x1 err', float((x1.cpu() - y.detach()).abs().max()), 'lp err',
      float((lp1.cpu() - lp.detach().squeeze(-1)).abs().max()))
torch.cuda.synchronize(); t0 = time.time()
gx0, gl0, gctx, gpar, gt, info, rc = ops.cnf_adjoint(x1, lp1, gx1.cuda(), gl1.cuda(), eg, cg, pack, T)
torch.cuda.synchronize()
print('adjoint rc', rc, info, 'first dt', torch.tensor(info[6], dtype=torch.int32).view(torch.float32).item(),
      '%.1f ms' % ((time.time() - t0) * 1e3))


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


print('gx0   rel', rel(gx0.cpu(), xo.grad))
print('glogp rel', rel(gl0.cpu(), lo.grad.squeeze(-1)))
print('gctx  rel', rel(gctx.cpu(), co.grad))
off = 0
gp = gpar.cpu()
for name in orc._cnf_func.names:
    ref = orc.sd[name].grad
    mine = gp[off:off + ref.numel()].view_as(ref)
    off += ref.numel()
    print('%-60s rel %.3g  |ref| %.3g' % (name, rel(mine, ref), float(ref.abs().max())))
s = orc.sd['point_cnf.chain.1.sqrt_end_time']
print('gtimes', gt.cpu().tolist(), 'sqrt_end_time grad mine', float(gt[1].cpu() * 2 * s.detach()), 'ref', float(s.grad))
