"""Kernel timeline of one CaSPR.decode via torch.profiler: busy time vs span, largest gaps, per-kernel totals."""
import os
import sys

import torch
from torch.profiler import profile, ProfilerActivity

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caspr_b200.models import CaSPR                          # noqa: E402
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences   # noqa: E402


def main():
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
    for _ in range(2):
        model.decode(z, P, y=y, e=e)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        model.decode(z, P, y=y, e=e)
        torch.cuda.synchronize()
    evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda ev: ev.time_range.start)
    span = evs[-1].time_range.end - evs[0].time_range.start
    busy = sum(ev.time_range.end - ev.time_range.start for ev in evs)
    gaps = []
    for a, b in zip(evs, evs[1:]):
        gp = b.time_range.start - a.time_range.end
        if gp > 0:
            gaps.append((gp, a.name[:50], b.name[:50]))
    gaps.sort(reverse=True)
    print('decode: kernels %d span %.2f ms busy %.2f ms total gap %.2f ms' % (
        len(evs), span / 1e3, busy / 1e3, sum(gp for gp, _, _ in gaps) / 1e3))
    for gp, a, b in gaps[:14]:
        print('   gap %7.1f us after %-50s before %s' % (gp, a, b))
    agg = {}
    for ev in evs:
        k = ev.name[:70]
        c, t = agg.get(k, (0, 0))
        agg[k] = (c + 1, t + ev.time_range.end - ev.time_range.start)
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        print('   %-70s %4d %8.2f ms' % (k, c, v / 1e3))


if __name__ == '__main__':
    main()
