"""Time CaSPR.encode (CUDA graph replay) at config 2: python tools/time_encode.py [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caspr_b200.models import CaSPR
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B, T, N = (int(v) for v in os.environ.get('SHAPE', '8,10,1024').split(','))
dev = 'cuda:0'
model = CaSPR().to(dev).eval()
model.load_state_dict(synthetic_state_dict(0, cnf_init='vigorous'))
x, _ = synthetic_sequences(B, T, N, seed=100)
x = x.to(dev)
for _ in range(5):
    model.encode(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    model.encode(x)
e1.record()
torch.cuda.synchronize()
print('encode %.3f ms (B=%d T=%d N=%d, %d reps)' % (e0.elapsed_time(e1) / reps, B, T, N, reps))
