"""Development check: training-mode encoder forward/backward against the CPU oracle's autograd."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from caspr_b200.models import CaSPR
from caspr_b200.models.encoder_train import EncoderTrainer
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences
from oracle.train_oracle import TrainOracle

B, T, N = int(os.environ.get('B', 1)), int(os.environ.get('T', 2)), int(os.environ.get('N', 1024))
sd = synthetic_state_dict(0, cnf_init='vigorous')
x, nocs = synthetic_sequences(B, T, N, seed=5)
g = torch.Generator().manual_seed(7)
gz = torch.randn(B, 1600, generator=g)
gt = torch.randn(B, T, N, 4, generator=g)
model = CaSPR().cuda().train()
model.load_state_dict(sd)
tr = EncoderTrainer(model.encoder)
with torch.no_grad():
    z0, tn = tr.forward(x.cuda())
    model.eval()
    z0e, tne = model.encoder(x.cuda())
orc = TrainOracle(sd)
z0o, tno = orc.encode(x)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))
print('sa_feat_4', rel(tr.sa_feats[5].cpu(), orc.trace['sa_feat_4'].detach().transpose(1, 2)))
print('global_max', rel(tr.p3.out.view(B, T * N, -1).max(1)[0].cpu(), orc.trace['global_max'].detach()))
print('local_feat', rel(tr.h1.x[:, :512].cpu().view(B, T * N, 512), orc.trace['local_feat'].detach().transpose(1, 2)))
print('z0 train vs oracle', rel(z0.cpu(), z0o.detach()), 'eval vs oracle', rel(z0e.cpu(), z0o.detach()))
print('tnocs train vs oracle', rel(tn.cpu(), tno.detach()), 'eval vs oracle', rel(tne.cpu(), tno.detach()))
for k in list(orc.trace):
    if k.startswith(('sa_out_', 'fp_out_', 'local_feat')):
        orc.trace[k].retain_grad()
((z0o * gz).sum() + (tno * gt).sum()).backward()
tr.debug = {}
with torch.no_grad():
    grads = tr.backward(gz.cuda(), gt.cuda())
def l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm())
dbg = tr.debug
for i, d in enumerate(dbg['d_fp']):          # d_fp[0] = grad of FP4 output, d_fp[1] = FP3 output ...
    r = orc.trace['fp_out_%d' % (4 - i)].grad.transpose(1, 2).reshape(-1, d.shape[1])
    print('d fp_out_%d l2rel %.3g' % (4 - i, l2(d.cpu(), r)))
for lvl in range(1, 6):
    r = orc.trace['sa_out_%d' % lvl].grad.transpose(1, 2)
    print('d sa_out_%d l2rel %.3g' % (lvl, l2(dbg['d_sa'][lvl].cpu(), r)))
ref = orc.parameters()
rows = []
for k, p in model.named_parameters():
    if not k.startswith('encoder.'):
        continue
    r = ref[k].grad
    mine = grads.get(p)
    if mine is None:
        print('MISSING', k); continue
    mc = mine.cpu().double().flatten(); rd = r.double().flatten()
    rows.append((rel(mine.cpu(), r), k, float((mc - rd).norm() / rd.norm()), float(torch.dot(mc, rd) / (mc.norm() * rd.norm()))))
for err, k, m, c in rows:
    print('%-80s rel %.3g l2rel %.3g cos %.6f' % (k, err, m, c))
