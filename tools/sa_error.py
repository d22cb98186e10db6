"""Error of the two fused set-abstraction kernels (SIMT `sa_fused`, tensor-core `sa_mma`) against the fp64 chain on the
model's own SA level 1-2 inputs: python tools/sa_error.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caspr_b200 import ops
from caspr_b200.models import CaSPR
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences

dev = 'cuda:0'
model = CaSPR().to(dev).eval()
model.load_state_dict(synthetic_state_dict(0, cnf_init='vigorous'))
x, _ = synthetic_sequences(2, 10, 1024, seed=31)
x4 = x.to(dev).view(-1, 4)
pts = model.encoder._local_input(x4).view(20, 1024, -1)
xyz = pts.reshape(-1, 9)[:, :3].contiguous().view(20, 1024, 3)
feat = pts[:, :, 3:]
for lvl in range(2):
    sa = model.encoder.local_extract.set_abstractions[lvl]
    Bp, N, _ = xyz.shape
    M = sa.num_points_out
    idx, new_xyz = ops.fps(xyz, M)
    g0, g1 = sa.grouper_modules
    bq = ops.ball_query2(xyz, new_xyz, g0.radius, g0.num_samples, g1.radius, g1.num_samples)
    absmax = ops.sa_absmax(xyz, feat)
    outs = []
    for s, (grouper, pn) in enumerate(zip(sa.grouper_modules, sa.pointnet_modules)):
        ns = grouper.num_samples
        cin = 3 + feat.shape[2]
        rows = ops.group_points(xyz, new_xyz, feat, bq[s]).double()
        h = rows.view(Bp * M, ns, cin).transpose(1, 2)
        for i in range(3):
            c, g = pn.conv_layers[i], pn.bn_layers[i]
            h = torch.nn.functional.conv1d(h, c.weight.double(), c.bias.double())
            h = torch.nn.functional.group_norm(h, 16, g.weight.double(), g.bias.double(), eps=g.eps)
            if i < 2:
                h = h.relu()
        ref = h.max(2)[0]
        o1 = torch.empty(Bp * M, pn.feat_size, device=dev)
        o2 = torch.empty(Bp * M, pn.feat_size, device=dev)
        ops.sa_fused(xyz, new_xyz, feat, bq[s], pn.conv_layers, pn.bn_layers, o1)
        ops.sa_mma(xyz, new_xyz, feat, bq[s], pn.conv_layers, pn.bn_layers, o2, absmax)
        e1 = (o1.double() - ref).abs()
        e2 = (o2.double() - ref).abs()
        print('level %d scale %d (ns %d): |ref|max %.2f   simt max %.2e mean %.2e   mma max %.2e mean %.2e' %
              (lvl + 1, s, ns, float(ref.abs().max()), float(e1.max()), float(e1.mean()), float(e2.max()), float(e2.mean())))
        if float(e2.max()) > 1e-4:
            w = int(e2.max(1)[0].argmax())
            b = bq[s].view(Bp * M, ns)[w]
            print('   worst ball %d: %d distinct rows, err/channel max at %d' % (w, len(torch.unique(b)), int(e2[w].argmax())))
        outs.append(o1)
    feat = torch.cat(outs, 1).view(Bp, M, -1)
    xyz = new_xyz
