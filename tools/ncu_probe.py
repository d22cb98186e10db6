import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
step = sys.argv[1]
import torch
if step == 'avail':
    print(torch.cuda.is_available()); torch.cuda.set_device(0); print(torch.zeros(4, device='cuda').sum().item())
elif step == 'pin':
    x = torch.randn(1000).pin_memory(); print(x.to('cuda', non_blocking=True).sum().item())
elif step == 'lib':
    from caspr_b200.build import build_library; build_library()
    from caspr_b200 import _lib, ops
    print(_lib.lib.caspr_launch_count())
elif step == 'inputs':
    import bench
    sd, x, y, e, kw, dims = bench.make_inputs(bench.DEFAULT_WORKLOAD, 0, 'vigorous'); print(dims)
elif step == 'model':
    from caspr_b200.models import CaSPR
    m = CaSPR().to('cuda:0').eval(); print('model ok')
