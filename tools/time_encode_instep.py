"""Why is CaSPR.encode slower inside the step than back to back?  Times encode (CUDA graph replay, config 2)
(a) back to back, (b) after an L2 flush (256 MB memset), (c) after a full decode.  python tools/time_encode_instep.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caspr_b200.models import CaSPR
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences

dev = 'cuda:0'
B, T, N, P = 8, 10, 1024, 2048
model = CaSPR().to(dev).eval()
model.load_state_dict(synthetic_state_dict(0, cnf_init='vigorous'))
x, _ = synthetic_sequences(B, T, N, seed=100)
x = x.to(dev)
g = torch.Generator().manual_seed(1000)
y = torch.randn(B * T, P, 3, generator=g).to(dev)
e = torch.randn(B * T, P, 3, generator=g).to(dev)
z0, _ = model.encode(x)
z = model.aggregate_and_solve_latent(z0, x[:, :, 0, 3] / 5.0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(pre, reps=10):
    tot = 0.0
    for _ in range(reps):
        pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model.encode(x)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


for _ in range(5):
    model.encode(x)
    model.decode(z, P, y=y, e=e)
torch.cuda.synchronize()
print('back to back      %.3f ms' % timed(lambda: None))
print('after L2 flush    %.3f ms' % timed(lambda: flush.zero_()))
print('after a decode    %.3f ms' % timed(lambda: model.decode(z, P, y=y, e=e)))
print('after decode+flush %.3f ms' % timed(lambda: (model.decode(z, P, y=y, e=e), flush.zero_())))
print('back to back      %.3f ms' % timed(lambda: None))
