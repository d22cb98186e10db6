import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caspr_b200.models import CaSPR
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences
B, T, N = int(os.environ.get('B', 2)), int(os.environ.get('T', 2)), 1024
sd = synthetic_state_dict(0, cnf_init='vigorous')
x, nocs = synthetic_sequences(B, T, N, seed=9)
x, nocs = x.cuda(), nocs.cuda()
model = CaSPR().cuda().train(); model.load_state_dict(sd)
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
e = torch.randn(B * T, N, 3, generator=torch.Generator().manual_seed(0)).cuda()
for it in range(int(os.environ.get('ITERS', 4))):
    torch.cuda.synchronize(); t0 = time.time()
    opt.zero_grad()
    out = model(x, nocs, e=e)
    torch.cuda.synchronize(); t1 = time.time()
    loss = 0.01 * out[0].sum(2).mean() + 100.0 * out[1][:, :, :, :4].mean()       # train_utils.py:148-166
    loss.backward()
    torch.cuda.synchronize(); t2 = time.time()
    bad = [k for k, p in model.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
    opt.step()
    torch.cuda.synchronize(); t3 = time.time()
    cnf = model.point_cnf.chain[1]
    print('it %d loss %.5f fwd %.0f ms bwd %.0f ms opt %.0f ms nfe %s cnf fwd info %s adj info %s latent adj %s nonfinite grads %d' % (
        it, float(loss.detach()), (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, model.get_nfe(), cnf.last_info[:4],
        cnf.last_adjoint_info[:4], model.latent_ode.solver.last_adjoint_info[:4], len(bad)), flush=True)
