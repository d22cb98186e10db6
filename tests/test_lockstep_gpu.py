"""Lock-step step control (caspr_cnf_flow_lockstep) on one GPU: with a world of one rank the
all-reduce hook is the identity, so the result must equal the plain solve bit for bit, the hook must run once per
attempted step plus once for the initial-step heuristic, and a failing hook must surface as an error.
The two-rank check against an unsharded batch is tools/sharded_parity.py (needs two GPUs)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def single_rank_group():
    import torch.distributed as dist
    if dist.is_initialized():
        yield None
        return
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device('cuda', 0))
    yield None
    dist.destroy_process_group()


def test_lockstep_world_of_one_is_identity(single_rank_group):
    from caspr_b200 import ops
    from caspr_b200.models import CaSPR
    from caspr_b200.sharding import lockstep, reconstruct_sharded
    from caspr_b200.synth import synthetic_state_dict, synthetic_sequences
    model = CaSPR().cuda().eval()
    model.load_state_dict(synthetic_state_dict(0, cnf_init='vigorous'))
    x, _ = synthetic_sequences(2, 3, 1024, seed=3)
    g = torch.Generator().manual_seed(2)
    y = torch.randn(6, 256, 3, generator=g).cuda()
    e = torch.randn(6, 256, 3, generator=g).cuda()
    ref = model.reconstruct(x.cuda(), num_points=256, y=y, e=e)
    nfe = [int(v) for v in model.get_nfe()]
    with lockstep(model):
        got = reconstruct_sharded(model, x.cuda(), num_points=256, y=y, e=e)
    assert [int(v) for v in model.get_nfe()] == nfe
    assert torch.equal(ref[2], got[2]) and torch.equal(ref[3], got[3])
    # hook bookkeeping through the raw operator
    cnf = model.point_cnf.chain[1]
    ctx = torch.randn(6, 1600, generator=g).cuda() * 0.5
    sync = ops.LockstepSync(6 * 256, torch.device('cuda', 0))
    x1, _, info, rc = ops.cnf_flow(y, None, e, ctx, cnf.weight_pack(), None, None, cnf.end_time(), True, 1e-5, 1e-5,
                                   ops.CNF_SIMT_FP32, sync=sync)
    assert rc == 0 and sync.calls == info[2] + info[3] + 1
    x2, _, info2, rc2 = ops.cnf_flow(y, None, e, ctx, cnf.weight_pack(), None, None, cnf.end_time(), True, 1e-5, 1e-5,
                                     ops.CNF_SIMT_FP32)
    assert rc2 == 0 and info2[:4] == info[:4] and torch.equal(x1, x2)
