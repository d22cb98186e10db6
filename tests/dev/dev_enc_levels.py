import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from caspr_b200.models import CaSPR
from caspr_b200.models.encoder_train import EncoderTrainer, _Layer
from caspr_b200 import ops, train_ops as tops
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences
sd = synthetic_state_dict(0, cnf_init='vigorous')
x, nocs = synthetic_sequences(1, 2, 1024, seed=5)
model = CaSPR().cuda().eval(); model.load_state_dict(sd)
enc = model.encoder; net = enc.local_extract
rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))
with torch.no_grad():
    x4 = x.cuda().view(-1, 4)
    pts = enc._local_input(x4).view(2, 1024, -1)
    xyz = pts.reshape(-1, 9)[:, :3].contiguous().view(2, 1024, 3)
    feats = pts[:, :, 3:]
    tr = EncoderTrainer(enc)
    for lvl, sa in enumerate(net.set_abstractions):
        nx_e, f_e = sa.forward_rows(xyz, feats)
        nx_t, f_t, saved = tr._sa_forward(sa, xyz, feats)
        print('SA', lvl, 'xyz', rel(nx_t, nx_e), 'feat', rel(f_t, f_e))
        # per-layer check of scale 0
        for s in range(2):
            pn = sa.pointnet_modules[s]; ns = sa.grouper_modules[s].num_samples
            rows = ops.group_points(xyz, nx_e, feats, saved['scales'][s]['idx'])
            h_e = rows
            for i, (conv, gn) in enumerate(zip(pn.conv_layers, pn.bn_layers)):
                h_e = ops.linear(h_e, conv.weight, conv.bias)
                pre = h_e.clone()
                ops.groupnorm(h_e, h_e.shape[0] // ns, ns, 16, gn.weight, gn.bias, relu=i < 2)
                L = saved['scales'][s]['layers'][i]
                print('   scale', s, 'layer', i, 'x', rel(L.x, rows if i == 0 else prev), 'pre', rel(L.pre, pre), 'out', rel(L.out, h_e))
                prev = h_e
        xyz, feats = nx_e, f_e
