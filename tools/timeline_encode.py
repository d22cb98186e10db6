"""Kernel timeline of one CaSPR.encode (graph and eager) via torch.profiler: busy time vs span, largest gaps."""
import os
import sys

import torch
from torch.profiler import profile, ProfilerActivity

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caspr_b200.models import CaSPR                          # noqa: E402
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences   # noqa: E402


def main():
    dev = 'cuda:0'
    model = CaSPR().to(dev).eval()
    model.load_state_dict(synthetic_state_dict(0, cnf_init='vigorous'))
    x, _ = synthetic_sequences(8, 10, 1024, seed=100)
    x = x.to(dev)
    for use_graph in (True, False):
        model.encoder.use_cuda_graph = use_graph
        for _ in range(3):
            model.encode(x)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            model.encode(x)
            torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        if not evs:
            print('no CUDA events recorded')
            continue
        span = evs[-1].time_range.end - evs[0].time_range.start
        busy = sum(e.time_range.end - e.time_range.start for e in evs)
        gaps = []
        for a, b in zip(evs, evs[1:]):
            g = b.time_range.start - a.time_range.end
            if g > 0:
                gaps.append((g, a.name[:50], b.name[:50]))
        gaps.sort(reverse=True)
        print('graph=%s kernels %d span %.2f ms busy %.2f ms total gap %.2f ms' % (
            use_graph, len(evs), span / 1e3, busy / 1e3, sum(g for g, _, _ in gaps) / 1e3))
        for g, a, b in gaps[:12]:
            print('   gap %7.1f us after %-50s before %s' % (g, a, b))
        agg = {}
        for e in evs:
            k = e.name[:60]
            agg[k] = agg.get(k, 0) + (e.time_range.end - e.time_range.start)
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:12]:
            print('   %-60s %8.2f ms' % (k, v / 1e3))


if __name__ == '__main__':
    main()
