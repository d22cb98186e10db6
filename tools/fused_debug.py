"""Where do the producer / MMA threads of the fused CNF kernel wait?  CASPR_CNF_FUSED_DEBUG=1 python tools/fused_debug.py"""
import ctypes, os, sys
os.environ['CASPR_CNF_FUSED_DEBUG'] = '1'
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caspr_b200.models import CaSPR
from caspr_b200._lib import lib
from caspr_b200.synth import synthetic_state_dict

dev = 'cuda:0'
B, T, P = 8, 10, 2048
model = CaSPR().to(dev).eval()
model.load_state_dict(synthetic_state_dict(0, cnf_init='vigorous'))
g = torch.Generator().manual_seed(1000)
y = torch.randn(B * T, P, 3, generator=g).to(dev)
e = torch.randn(B * T, P, 3, generator=g).to(dev)
z = (0.5 * torch.randn(B, T, 1600, generator=g)).to(dev)
for _ in range(2):
    model.decode(z, P, y=y, e=e)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (148 * 8))()
assert lib.caspr_cnf_fused_debug_read(buf, 148 * 8) == 0
a = np.array(buf[:]).reshape(148, 8).astype(np.float64)
names = ['prod wait empty', 'prod wait sa_full', 'prod total', 'prod wait sb_full', 'mma wait tempty', 'mma wait full', 'mma total', 'L0 wait sa_free']
for r in (0, 1):
    sel = a[r::2]
    print('rank', r, {n: '%.0f (%.0f%%)' % (sel[:, i].mean(), 100 * sel[:, i].mean() / max(sel[:, 2].mean(), 1)) for i, n in enumerate(names) if n != '-'})
