"""Development check: caspr_latent_ode_adjoint against the CPU training oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from caspr_b200 import ops
from caspr_b200.synth import synthetic_state_dict
from oracle.train_oracle import TrainOracle
from oracle import odeint001

B, T = int(os.environ.get('B', 5)), int(os.environ.get('T', 5))
sd = synthetic_state_dict(0, cnf_init='vigorous')
g = torch.Generator().manual_seed(3)
z0 = torch.randn(B, 64, generator=g)
gz = torch.randn(B, T, 64, generator=g)
times = torch.linspace(0, 1, T)
orc = TrainOracle(sd)
zo = z0.clone().requires_grad_(True)
pred = orc.latent_ode(zo, times)            # (B,T,64)
(pred * gz).sum().backward()
print('oracle nfe', orc.nfe)
p = 'latent_ode.ode_func.dynamics_net.'
Ws = [sd[p + '%d.weight' % l].cuda() for l in (0, 2, 4, 6)]
bs = [sd[p + '%d.bias' % l].cuda() for l in (0, 2, 4, 6)]
out, info, rc = ops.latent_ode_solve(z0.cuda(), Ws, bs, times.tolist(), 1e-3, 1e-3)
print('fwd', rc, info, 'err', float((out.permute(1, 0, 2).cpu() - pred.detach()).abs().max()))
torch.cuda.synchronize(); t0 = time.time()
gz0, gpar, info, rc = ops.latent_ode_adjoint(out, gz.permute(1, 0, 2).contiguous().cuda(), Ws, bs, times.tolist(), 1e-3, 1e-3)
torch.cuda.synchronize()
print('adjoint', rc, info, '%.1f ms' % ((time.time() - t0) * 1e3))
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
print('gz0 rel', rel(gz0.cpu(), zo.grad))
off = 0
for l in (0, 2, 4, 6):
    for suf in ('weight', 'bias'):
        ref = orc.sd[p + '%d.%s' % (l, suf)].grad
        mine = gpar[off:off + ref.numel()].cpu().view_as(ref); off += ref.numel()
        print(l, suf, 'rel %.3g' % rel(mine, ref), '|ref| %.3g' % float(ref.abs().max()))
