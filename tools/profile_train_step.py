"""One training step (config-5 per-GPU shape) for `ncu --metrics gpu__time_duration.sum`: two untimed steps, then one
step between cudaProfilerStart/Stop (run ncu with --profile-from-start off)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caspr_b200.models import CaSPR
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences
B, T, N = 8, 5, 1024
sd = synthetic_state_dict(0, cnf_init='vigorous')
x, nocs = synthetic_sequences(B, T, N, seed=200)
x, nocs = x.cuda(), nocs.cuda()
e = torch.randn(B * T, N, 3, generator=torch.Generator().manual_seed(0)).cuda()
model = CaSPR().cuda().train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
for it in range(3):
    model.load_state_dict(sd)
    torch.cuda.synchronize()
    if it == 2:
        torch.cuda.cudart().cudaProfilerStart()
    opt.zero_grad()
    nll, tl1 = model(x, nocs, e=e)
    (0.01 * nll.sum(2).mean() + 100.0 * tl1.mean()).backward()
    opt.step()
    torch.cuda.synchronize()
    if it == 2:
        torch.cuda.cudart().cudaProfilerStop()
