"""Training-mode TPointNet++: forward that keeps what the backward needs, and a hand-written backward.

The reference trains the encoder through torch autograd (``train_utils.py:173`` ``loss.backward()`` over
``tpointnet2.py:70-115``, ``pointnet.py:34-46``, ``pointnet2.py:217-249,361-419,483-525,649-708``).  Here every
operator's backward is a kernel of libcaspr_b200.so (``train_ops.py``); this file only sequences them.  The graph is
fixed, so instead of a generic tape the backward mirrors the forward block by block:

  layer   = Conv1d(k=1) -> GroupNorm(16) -> ReLU?           (``_Layer``)
  SA      = FPS, ball query, group gather, 3 layers, max over the ball   (per scale)
  FP      = three_nn, inverse-distance interpolation ++ skip, 2 layers
  head    = 2 layers, max over the sequence (z0) and ReLU -> Conv1d -> sigmoid (T-NOCS)

Index-producing operators (FPS, ball query, three_nn) carry no gradient, as in Kaolin.  Input coordinates carry no
gradient either (the training inputs are data), so the first layer of every chain skips its data gradient.
"""
import torch

from .. import ops
from .. import train_ops as tops

NUM_GROUPS = 16


class _Layer(object):
    """y = ReLU?(GroupNorm(conv(x))) with x, the pre-norm product and the statistics kept."""

    def __init__(self, x, conv, gn, samples, rps, relu, out=None):
        self.x, self.conv, self.gn, self.samples, self.rps, self.relu = x, conv, gn, samples, rps, relu
        self.pre = ops.linear(x, conv.weight, conv.bias)
        self.mr = tops.gn_moments(self.pre, samples, rps, NUM_GROUPS, eps=gn.eps)
        self.out = tops.gn_apply(self.pre, self.mr, samples, rps, NUM_GROUPS, gn.weight, gn.bias, relu, out=out)

    def backward(self, grads, d_out=None, d_max=None, argmax=None, need_dx=True):
        conv, gn = self.conv, self.gn
        d_pre, dgamma, dbeta = tops.gn_backward(self.pre, self.mr, self.samples, self.rps, NUM_GROUPS, gn.weight,
                                                gn.bias, self.relu, d_out=d_out, d_max=d_max, argmax=argmax)
        grads[gn.weight], grads[gn.bias] = dgamma, dbeta
        dW, db = tops.linear_wgrad(d_pre, self.x)
        grads[conv.weight], grads[conv.bias] = dW.view_as(conv.weight), db
        if not need_dx:
            return None
        w2d = conv.weight.detach().reshape(conv.weight.shape[0], conv.weight.shape[1])
        return ops.linear(d_pre, tops.transpose(w2d), None)


def _plain_linear_backward(grads, conv, x, d_y, relu_x=False):
    """y = conv(act(x)) without normalisation: parameter gradients and d(act(x))."""
    dW, db = tops.linear_wgrad(d_y, x, relu_x=relu_x)
    grads[conv.weight], grads[conv.bias] = dW.view_as(conv.weight), db
    w2d = conv.weight.detach().reshape(conv.weight.shape[0], conv.weight.shape[1])
    return ops.linear(d_y, tops.transpose(w2d), None)


class EncoderTrainer(object):
    """One forward/backward pass of a ``TPointNet2`` module (``enc``)."""

    def __init__(self, enc):
        self.enc = enc

    # ------------------------------------------------------------------------------ forward
    def forward(self, x):
        enc = self.enc
        x = x.to(torch.float32).contiguous()
        B, T, N, _ = x.shape
        R = B * T * N
        self.dims = (B, T, N)
        x4 = x.view(R, 4)
        L, G, Pf = enc.local_feat_size, enc.global_feat_size, enc.space_time_pt_feat
        feat = torch.empty(R, L + G + Pf, dtype=torch.float32, device=x.device)
        pn = enc.global_extract
        self.p1 = _Layer(x4, pn.conv1, pn.bn1, B, T * N, True, out=feat[:, L + G:])
        self.p2 = _Layer(self.p1.out, pn.conv2, pn.bn2, B, T * N, True)
        self.p3 = _Layer(self.p2.out, pn.conv3, pn.bn3, B, T * N, False)
        gmax, self.garg = tops.rowmax(self.p3.out, B, T * N)
        ops.broadcast_rows(gmax, T * N, feat[:, L:L + G])
        local_in = enc._local_input(x4)
        self._pointnet2_forward(enc.local_extract, local_in.view(B * T, N, -1), feat[:, :L])
        self.h1 = _Layer(feat, enc.conv1, enc.bn1, B, T * N, True)
        self.h2 = _Layer(self.h1.out, enc.conv2, enc.bn2, B, T * N, False)
        z0, self.zarg = tops.rowmax(self.h2.out, B, T * N)
        tnocs = None
        self.t = None
        if enc.regress_tnocs:
            self.t = ops.linear(self.h2.out, enc.conv3.weight, enc.conv3.bias, act_in=ops.ACT_RELU,
                                act_out=ops.ACT_SIGMOID)
            tnocs = self.t[:, :4].reshape(B, T, N, 4)
        return z0, tnocs

    def _pointnet2_forward(self, net, points, out):
        Bp, N, cdim = points.shape
        flat = points.reshape(Bp * N, cdim)
        xyz = flat[:, :3].contiguous().view(Bp, N, 3)
        features = points[:, :, 3:] if cdim > 3 else None
        xyz_list, feat_list = [xyz], [features]
        self.sa = []
        for sa in net.set_abstractions:
            xyz, features, saved = self._sa_forward(sa, xyz, features)
            self.sa.append(saved)
            xyz_list.append(xyz)
            feat_list.append(features)
        self.sa_feats = list(feat_list)
        self.fp = []
        ti = -2
        for fp in net.feature_propagators:
            lvl = len(feat_list) + ti
            new, saved = self._fp_forward(fp, xyz_list[ti], xyz_list[ti + 1], feat_list[ti], feat_list[ti + 1])
            saved['level'] = lvl
            self.fp.append(saved)
            feat_list[ti] = new
            ti -= 1
        h = feat_list[0].reshape(Bp * N, -1)
        conv0, gn, _, conv1 = net.final_layers
        self.f0 = _Layer(h, conv0, gn, Bp, N, True)
        self.f1_conv = conv1
        ops.linear(self.f0.out, conv1.weight, conv1.bias, out=out)

    def _sa_forward(self, sa, xyz, features):
        Bp, N, _ = xyz.shape
        M = sa.num_points_out
        idx, new_xyz = ops.fps(xyz, M)
        g0, g1 = sa.grouper_modules
        bq = ops.ball_query2(xyz, new_xyz, g0.radius, g0.num_samples, g1.radius, g1.num_samples)
        out = torch.empty(Bp * M, sa.get_num_features_out(), dtype=torch.float32, device=xyz.device)
        scales = []
        off = 0
        for s, (grouper, pointnet) in enumerate(zip(sa.grouper_modules, sa.pointnet_modules)):
            ns = grouper.num_samples
            rows = ops.group_points(xyz, new_xyz, features, bq[s])
            layers = []
            h = rows
            last = len(pointnet.conv_layers) - 1
            for i, (conv, gn) in enumerate(zip(pointnet.conv_layers, pointnet.bn_layers)):
                layers.append(_Layer(h, conv, gn, Bp * M, ns, i < last))
                h = layers[-1].out
            _, arg = tops.rowmax(h, Bp * M, ns, maxout=out[:, off:off + pointnet.feat_size])
            scales.append({'layers': layers, 'arg': arg, 'idx': bq[s], 'off': off, 'width': pointnet.feat_size})
            off += pointnet.feat_size
        C = 0 if features is None else features.shape[2]
        return new_xyz, out.view(Bp, M, -1), {'scales': scales, 'N': N, 'C': C, 'M': M}

    def _fp_forward(self, fp, xyz, xyz_prev, features, features_prev):
        Bp, n, _ = xyz.shape
        dist, idx = ops.three_nn(xyz, xyz_prev)
        h = ops.three_interp_concat(features_prev, idx, dist, features)
        layers = []
        x = h
        for i in range(0, len(fp.unit_pointnet), 3):
            layers.append(_Layer(x, fp.unit_pointnet[i], fp.unit_pointnet[i + 1], Bp, n, True))
            x = layers[-1].out
        saved = {'layers': layers, 'idx': idx, 'dist': dist, 'm': xyz_prev.shape[1], 'Cp': features_prev.shape[2],
                 'Cs': 0 if features is None else features.shape[2], 'n': n}
        return x.view(Bp, n, -1), saved

    # ----------------------------------------------------------------------------- backward
    def backward(self, g_z0, g_tnocs):
        """g_z0 (B,latent) or None, g_tnocs (B,T,N,4) or None -> {parameter: gradient}."""
        enc = self.enc
        B, T, N = self.dims
        R = B * T * N
        L, G, Pf = enc.local_feat_size, enc.global_feat_size, enc.space_time_pt_feat
        grads = {}
        dev = self.h2.out.device
        d_h2 = None
        if g_tnocs is not None and self.t is not None:
            t = self.t
            d_pre = (g_tnocs.reshape(R, 4).to(torch.float32) * t * (1.0 - t)).contiguous()
            d_h2 = _plain_linear_backward(grads, enc.conv3, self.h2.out, d_pre, relu_x=True)
            tops.rows_update(d_h2, d_h2, relu_ref=self.h2.out)
        elif self.t is not None:
            grads[enc.conv3.weight] = torch.zeros_like(enc.conv3.weight)
            grads[enc.conv3.bias] = torch.zeros_like(enc.conv3.bias)
        if g_z0 is None:
            g_z0 = torch.zeros(B, enc.latent_feat_size, dtype=torch.float32, device=dev)
        d_h1 = self.h2.backward(grads, d_out=d_h2, d_max=g_z0.contiguous(), argmax=self.zarg)
        d_feat = self.h1.backward(grads, d_out=d_h1)
        # global PointNet: the repeated global max receives the per-sequence column sums
        d_gmax = torch.empty(B, G, dtype=torch.float32, device=dev)
        for s in range(B):
            tops.colsum(d_feat[s * T * N:(s + 1) * T * N, L:L + G], d_gmax[s])
        d_p2 = self.p3.backward(grads, d_max=d_gmax, argmax=self.garg)
        d_pf = self.p2.backward(grads, d_out=d_p2)
        tops.rows_update(d_feat[:, L + G:], d_pf, accumulate=True)
        self.p1.backward(grads, d_out=d_pf, need_dx=False)
        self._pointnet2_backward(enc.local_extract, d_feat[:, :L], grads)
        return grads

    def _pointnet2_backward(self, net, d_out, grads):
        d_f0 = _plain_linear_backward(grads, self.f1_conv, self.f0.out, d_out)
        d_cur = self.f0.backward(grads, d_out=d_f0)            # gradient of the last FP output (Bp*N, 512)
        # gradient buffers of the set-abstraction outputs (levels 1..5)
        d_sa = [None] + [torch.zeros_like(f) for f in self.sa_feats[1:]]
        debug = getattr(self, 'debug', None)
        if debug is not None:
            debug['d_fp'] = [d_cur]
            debug['d_sa'] = d_sa
        for saved in reversed(self.fp):
            lvl = saved['level']
            la = saved['layers']
            d = d_cur
            for layer in reversed(la):
                d = layer.backward(grads, d_out=d)
            Cp, Cs, m = saved['Cp'], saved['Cs'], saved['m']
            Bp = saved['idx'].shape[0]
            d_prev = torch.zeros(Bp, m, Cp, dtype=torch.float32, device=d.device)
            tops.three_interp_bwd(d[:, :Cp], saved['idx'], saved['dist'], m, Cp, d_prev)
            if lvl >= 1 and Cs > 0:
                tops.rows_update(d[:, Cp:], d_sa[lvl].view(-1, Cs), accumulate=True)
            if lvl + 1 == len(self.sa_feats) - 1:
                # the coarsest propagator interpolates the last set-abstraction output itself
                tops.rows_update(d_prev.view(-1, Cp), d_sa[lvl + 1].view(-1, Cp), accumulate=True)
                d_cur = None
            else:
                d_cur = d_prev.view(-1, Cp)
                if debug is not None:
                    debug['d_fp'].append(d_cur)
        for k in range(len(self.sa) - 1, -1, -1):
            saved = self.sa[k]
            d_level = d_sa[k + 1].view(-1, d_sa[k + 1].shape[2])
            need_dx = k >= 1
            for sc in saved['scales']:
                layers = sc['layers']
                d = layers[-1].backward(grads, d_max=d_level[:, sc['off']:sc['off'] + sc['width']], argmax=sc['arg'])
                for layer in reversed(layers[1:-1]):
                    d = layer.backward(grads, d_out=d)
                d = layers[0].backward(grads, d_out=d, need_dx=need_dx)
                if need_dx:
                    tops.group_points_bwd(d, sc['idx'], saved['N'], saved['C'], d_sa[k])


class EncodeFunction(torch.autograd.Function):
    """autograd bridge: (x, encoder, *parameters) -> (z0, tnocs); the backward hands every parameter its gradient."""

    @staticmethod
    def forward(ctx, x, enc, *params):
        trainer = EncoderTrainer(enc)
        z0, tnocs = trainer.forward(x)
        ctx.trainer, ctx.params, ctx.has_tnocs = trainer, params, tnocs is not None
        if tnocs is None:
            tnocs = z0.new_zeros(0)
        return z0, tnocs

    @staticmethod
    def backward(ctx, g_z0, g_tnocs):
        grads = ctx.trainer.backward(g_z0, g_tnocs if ctx.has_tnocs else None)
        ctx.trainer = None
        return (None, None) + tuple(grads.get(p) for p in ctx.params)
