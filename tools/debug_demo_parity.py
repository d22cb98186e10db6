"""Where does the encoder deviate from the oracle on the real demo sequences?  (GPU box; fixture inputs.)"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caspr_b200.synth import synthetic_state_dict
from caspr_b200.models import CaSPR
from caspr_b200 import ops
from oracle.caspr_oracle import CasprOracle

g = dict(np.load(os.path.join(ROOT, 'tests/golden/caspr_r2.npz')))
x = torch.from_numpy(g['demo_x'])
sd = synthetic_state_dict(0, 'vigorous')
o = CasprOracle(sd)
z_ref, _ = o.encode(x)
for eng in ('auto', 'simt'):
    ops.LINEAR_ENGINE = eng
    m = CaSPR().cuda().eval()
    m.load_state_dict(sd)
    m.encoder.trace = {}
    z0, tn = m.encode(x.cuda())
    tr = m.encoder.trace
    print('engine', eng, 'z0 rel err', float((z0.cpu() - z_ref).abs().max() / z_ref.abs().max()))
    z64 = torch.from_numpy(g['demo_z0_f64'])
    print('   vs f64: gpu %.3g   ref32 %.3g' % (float((z0.cpu() - z64).abs().max() / z64.abs().max()),
                                               float((z_ref - z64).abs().max() / z64.abs().max())))
    if eng == 'auto':
        torch.manual_seed(17); y = torch.randn(10, 256, 3); e = torch.randn(10, 256, 3)
        _, _, xr, tn2 = m.reconstruct(x.cuda(), num_points=256, y=y, e=e.cuda())
        x64 = torch.from_numpy(g['demo_x_rec_f64']); x32 = torch.from_numpy(g['demo_x_rec'])
        print('   x_rec vs f64: gpu %.3g  ref32 %.3g   gpu vs ref32 %.3g  nfe %s' % (
            float((xr.cpu() - x64).abs().max() / x64.abs().max()), float((x32 - x64).abs().max() / x64.abs().max()),
            float((xr.cpu() - x32).abs().max() / x32.abs().max()), m.get_nfe()))
        t64 = torch.from_numpy(g['demo_tnocs_f64']); t32 = torch.from_numpy(g['demo_tnocs'])
        print('   tnocs vs f64: gpu %.3g  ref32 %.3g' % (float((tn2[:, :, ::4].cpu() - t64).abs().max()), float((t32 - t64).abs().max())))
    for l in range(5):
        fe = torch.equal(tr['fps_idx'][l].cpu(), o.trace['fps_idx_%d' % l])
        be = [torch.equal(tr['ball_idx'][l][s].cpu(), o.trace['ball_idx_%d_%d' % (l, s)]) for s in range(2)]
        ref = o.trace['sa_out_%d' % (l + 1)].transpose(1, 2)          # (B',M,C)
        got = tr['sa_out'][l].cpu()
        d = (got - ref).abs()
        worst = d.view(d.shape[0], -1).max(1)[0]
        print(' level', l, 'fps', fe, 'ball', be, 'sa_out rel err %.3g' % float(d.max() / ref.abs().max()),
              'per-cloud max', [round(float(v), 6) for v in worst])
        if not fe:
            a, b = tr['fps_idx'][l].cpu(), o.trace['fps_idx_%d' % l]
            for c in range(a.shape[0]):
                ne = (a[c] != b[c]).nonzero()
                if len(ne):
                    j = int(ne[0])
                    print('   cloud', c, 'first FPS mismatch at pick', j, 'gpu', int(a[c, j]), 'oracle', int(b[c, j]))
