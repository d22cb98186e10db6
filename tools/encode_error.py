"""z0 / T-NOCS error of CaSPR.encode against the two-sequence reference fixture: python tools/encode_error.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caspr_b200.models import CaSPR
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences

gold = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'caspr_r2.npz'))
model = CaSPR().to('cuda:0').eval()
model.load_state_dict(synthetic_state_dict(0, cnf_init='vigorous'))
x, _ = synthetic_sequences(2, 10, 1024, seed=31)
z0, tn = model.encode(x.to('cuda:0'))
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
print('z0 rel %.3e   tnocs rel %.3e' % (rel(z0.cpu().numpy(), gold['b2_z0']),
                                        rel(tn[:, ::3, ::8].cpu().numpy(), gold['b2_tnocs_frame'])))
