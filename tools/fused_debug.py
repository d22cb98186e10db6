"""Where do the producer / MMA threads of the fused CNF kernel wait?  CASPR_CNF_FUSED_DEBUG=1 python tools/fused_debug.py"""
import ctypes, os, sys
os.environ.setdefault('CASPR_CNF_FUSED_DEBUG', '1')
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
from caspr_b200 import ops
pack = model.point_cnf.chain[1].weight_pack()
for _ in range(2):                   # ONE dynamics evaluation of all points (experiment modes produce garbage values)
    ops.cnf_feval(y, e, z.reshape(B * T, 1600), pack, 0.37, engine=ops.CNF_TC_FP16X3)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (148 * 24))()
assert lib.caspr_cnf_fused_debug_read(buf, 148 * 24) == 0
a = np.array(buf[:]).reshape(148, 24).astype(np.float64)
names = ['prod wait empty', 'prod wait sa_full', 'prod total', 'prod wait sb_full', 'mma wait tempty', 'mma wait full', 'mma total', 'L0 wait sa_free']
for r in (0, 1):
    sel = a[r::2]
    print('rank', r, {n: '%.0f (%.0f%%)' % (sel[:, i].mean(), 100 * sel[:, i].mean() / max(sel[:, 2].mean(), 1)) for i, n in enumerate(names) if n != '-'})

tiles = (B * T * P // 64) / 148.0
for g_ in (0, 1):
    w = a[:, 8 + 8 * g_: 12 + 8 * g_].mean(0) / tiles
    c = a[:, 12 + 8 * g_: 16 + 8 * g_].mean(0) / tiles
    print('epilogue group', g_, 'per tile: wait tfull [L1.n0 L1.n1 L2.n0 L2.n1] =', np.round(w).astype(int), ' work =', np.round(c).astype(int))
print('tile period', round(a[:, 6].max() / tiles), 'cycles; MMA floor 4 x 12288 = 49152')
