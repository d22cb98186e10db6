"""Wall / device time of the three backward pieces of a training step at the config-5 per-GPU shape."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caspr_b200 import ops, _lib
from caspr_b200.models import CaSPR
from caspr_b200.models.encoder_train import EncoderTrainer
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences
B, T, N = 8, 5, 1024
sd = synthetic_state_dict(0, cnf_init='vigorous')
model = CaSPR().cuda().train(); model.load_state_dict(sd)
x, nocs = synthetic_sequences(B, T, N, seed=200)
x = x.cuda()
g = torch.Generator().manual_seed(0)
def timed(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    n0 = _lib.lib.caspr_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    print('%-28s device %.2f ms  wall %.2f ms  launches %d' % (name, e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) * 1e3 / reps,
          (_lib.lib.caspr_launch_count() - n0) // reps), flush=True)
    return out
with torch.no_grad():
    tr = EncoderTrainer(model.encoder)
    z0, tn = timed('encoder forward (train)', lambda: tr.forward(x))
    gz, gt = torch.randn_like(z0), torch.randn_like(tn)
    timed('encoder backward', lambda: tr.backward(gz, gt))
    cnf = model.point_cnf.chain[1]; pack, Tend = cnf.weight_pack(), cnf.end_time()
    pts = (torch.randn(B * T, N, 3, generator=g) * 0.3).cuda(); e = torch.randn(B * T, N, 3, generator=g).cuda()
    ctx = (torch.randn(B * T, 1600, generator=g) * 0.5).cuda()
    lp0 = torch.zeros(B * T, N, device='cuda')
    for eng, nm in ((ops.CNF_TC_FP16X3, 'tc'), (ops.CNF_SIMT_FP32, 'simt')):
        x1, lp1, info, rc = timed('cnf forward flow (%s)' % nm, lambda: ops.cnf_flow(pts, lp0, e, ctx, pack, None, None, Tend, False, 1e-5, 1e-5, eng))
        gx, gl = torch.randn_like(x1), torch.randn_like(lp1)
        out = timed('cnf adjoint (%s)' % nm, lambda: ops.cnf_adjoint(x1, lp1, gx, gl, e, ctx, pack, Tend, engine=eng))
        print('   adjoint info', out[5][:4])
    p = 'latent_ode.ode_func.dynamics_net.'
    Ws = [sd[p + '%d.weight' % l].cuda() for l in (0, 2, 4, 6)]; bs = [sd[p + '%d.bias' % l].cuda() for l in (0, 2, 4, 6)]
    zl = torch.randn(B, 64, generator=g).cuda(); times = torch.linspace(0, 1, T).tolist()
    out, info, rc = timed('latent forward', lambda: ops.latent_ode_solve(zl, Ws, bs, times, 1e-3, 1e-3))
    go = torch.randn_like(out)
    r = timed('latent adjoint', lambda: ops.latent_ode_adjoint(out, go, Ws, bs, times, 1e-3, 1e-3))
    print('   latent adjoint info', r[2][:4])
