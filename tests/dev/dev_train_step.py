"""Development check: one full training step (forward + backward) against the CPU training oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from caspr_b200.models import CaSPR
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences
from oracle.train_oracle import TrainOracle

B, T, N = int(os.environ.get('B', 1)), int(os.environ.get('T', 2)), int(os.environ.get('N', 1024))
sd = synthetic_state_dict(0, cnf_init='vigorous')
x, nocs = synthetic_sequences(B, T, N, seed=5)
g = torch.Generator().manual_seed(7)
e = torch.randn(B * T, N, 3, generator=g)

model = CaSPR().cuda().train()
model.load_state_dict(sd)
torch.cuda.synchronize(); t0 = time.time()
nll, tl1 = model(x.cuda(), nocs.cuda(), e=e.cuda())
loss = TrainOracle.loss(nll, tl1)
loss.backward()
torch.cuda.synchronize()
print('gpu step %.1f ms, loss %.6f, nfe %s' % ((time.time() - t0) * 1e3, float(loss), model.get_nfe()), flush=True)

t0 = time.time()
orc = TrainOracle(sd)
nll_o, tl1_o = orc.forward_train(x, nocs, e)
loss_o = TrainOracle.loss(nll_o, tl1_o)
loss_o.backward()
print('oracle step %.1f s, loss %.6f, nfe %s' % (time.time() - t0, float(loss_o), orc.nfe), flush=True)
print('nll rel', float((nll.detach().cpu() - nll_o.detach()).abs().max() / nll_o.detach().abs().max()))
ref = orc.parameters()
worst = []
for k, p in model.named_parameters():
    if k.startswith('latent_ode.solver.'):
        continue
    r = ref[k].grad
    if p.grad is None:
        print('MISSING grad', k); continue
    if r is None:
        print('oracle has no grad for', k); continue
    a, b = p.grad.cpu().double().flatten(), r.double().flatten()
    err = float((a - b).norm() / b.norm().clamp_min(1e-30))
    worst.append((err, k, float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))))
worst.sort(reverse=True)
for err, k, m in worst[:25]:
    print('%-75s l2rel %.3g cos %.5f' % (k, err, m))
print('params compared', len(worst), 'median l2rel', worst[len(worst) // 2][0])
for err, k, m in worst:
    if not k.startswith('encoder.'):
        print('%-75s l2rel %.3g cos %.5f' % (k, err, m))
