"""Independent CPU restatement of the CaSPR reconstruction hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity unpinned upstream
(the reference has no tests); pinned here against fixtures frozen from the
reference's own modules (``tests/golden/make_golden.py``).

Functional fp32 PyTorch-on-CPU code driven by a plain ``state_dict`` with the
reference's 238-key layout.  Each function cites the reference lines it follows
(paths relative to ``/root/reference/caspr/models``).  Nothing here depends on
``/root/reference`` at run time.
"""
from math import log, pi

import torch
import torch.nn.functional as F

from . import pointnet2_ops as pn2
from . import odeint001

NUM_GROUPS = 16          # pointnet2.py:12

# (num_points_out, [radius indices], [num_samples]) -- pointnet2.py:64-146
SA_POINTS = [1024, 512, 256, 64, 16]
SA_SAMPLES = [16, 32]


class CasprOracle(object):
    def __init__(self, state_dict, radii_list=(0.02, 0.05, 0.1, 0.2, 0.4, 0.8),
                 motion_feat_size=64, regress_tnocs=True, augment_quad=True, augment_pairs=True,
                 dtype=torch.float32):
        # dtype=torch.float64 evaluates the same network in double precision (geometry indices are still taken in the
        # declared fp32 arithmetic): the yardstick for how well-conditioned an fp32 result is on a given input
        self.dtype = dtype
        self.sd = {k: v.detach().to(dtype).cpu() if v.is_floating_point() else v.cpu()
                   for k, v in state_dict.items()}
        self.radii = list(radii_list)
        self.motion = motion_feat_size
        self.regress_tnocs = regress_tnocs
        self.augment_quad = augment_quad
        self.augment_pairs = augment_pairs
        self.nfe = [0, 0]            # [latent, cnf]   (caspr.py:198-202)
        self.trace = {}              # intermediate tensors for op-level parity tests
        # {tag: ReLU sign pattern (bool) / max-pool winner index} recorded by ANOTHER implementation of the encoder.  When
        # set, the encoder's discrete decisions are replayed from it instead of being taken from the oracle's own values
        # (gradient parity with frozen decisions: tests/test_train_gpu.py); trace['flips'] counts where they differ.
        self.decisions = None
        self.record = None           # set to {} to have the oracle's own decisions stored under the same tags

    # ------------------------------------------------------------------ helpers
    def _conv(self, x, key):
        """1x1 Conv1d on (B,C,L)."""
        return F.conv1d(x, self.sd[key + '.weight'], self.sd[key + '.bias'])

    def _gn(self, x, key):
        return F.group_norm(x, NUM_GROUPS, self.sd[key + '.weight'], self.sd[key + '.bias'], eps=1e-5)

    def _relu(self, x, tag):
        if self.record is not None:
            self.record[tag] = x.detach() > 0
        if self.decisions is None:
            return F.relu(x)
        mask = self.decisions[tag]
        fl = self.trace.setdefault('flips', {'relu': 0, 'relu_total': 0, 'max': 0, 'max_total': 0})
        fl['relu'] += int(((x > 0) != mask).sum())
        fl['relu_total'] += mask.numel()
        return x * mask.to(x.dtype)

    def _max(self, x, tag):
        """max over the last axis of (B, C, L)."""
        if self.record is not None:
            self.record[tag] = torch.max(x.detach(), 2)[1]
        if self.decisions is None:
            return torch.max(x, 2)[0]
        arg = self.decisions[tag].long()
        fl = self.trace.setdefault('flips', {'relu': 0, 'relu_total': 0, 'max': 0, 'max_total': 0})
        fl['max'] += int((torch.max(x, 2)[1] != arg).sum())
        fl['max_total'] += arg.numel()
        return x.gather(2, arg.unsqueeze(2)).squeeze(2)

    # ------------------------------------------------------------------ encoder
    def pointnet_global(self, x):
        """pointnet.py:34-46.  x (B,4,L) -> (B,1088,L) = [global max 1024 | pointfeat 64]."""
        p = 'encoder.global_extract.'
        L = x.shape[2]
        x = self._relu(self._gn(self._conv(x, p + 'conv1'), p + 'bn1'), 'pn_r0')
        pointfeat = x
        x = self._relu(self._gn(self._conv(x, p + 'conv2'), p + 'bn2'), 'pn_r1')
        x = self._gn(self._conv(x, p + 'conv3'), p + 'bn3')
        g = self._max(x, 'pn_max').unsqueeze(2)
        self.trace['global_max'] = g[:, :, 0]
        return torch.cat([g.repeat(1, 1, L), pointfeat], 1)

    def _sa_pointnet(self, x, prefix, tag):
        """pointnet2.py:649-708 as used by SA (global_feat=True, GroupNorm, transposed input).

        x (B'*M, C, ns): [conv,GN,ReLU] x2, conv, GN (no ReLU, :693), max over ns (:698).
        """
        x = self._relu(self._gn(self._conv(x, prefix + 'conv_layers.0'), prefix + 'bn_layers.0'), tag + '_r0')
        x = self._relu(self._gn(self._conv(x, prefix + 'conv_layers.1'), prefix + 'bn_layers.1'), tag + '_r1')
        x = self._gn(self._conv(x, prefix + 'conv_layers.2'), prefix + 'bn_layers.2')
        return self._max(x, tag + '_max')

    def set_abstraction(self, level, xyz, features):
        """pointnet2.py:361-419."""
        B = xyz.shape[0]
        M = SA_POINTS[level]
        idx = pn2.furthest_point_sampling(xyz, M)                                  # :384
        new_xyz = pn2.fps_gather_by_index(xyz.transpose(1, 2).contiguous(), idx)   # :385
        new_xyz = new_xyz.transpose(1, 2).contiguous()                             # :387
        self.trace['fps_idx_%d' % level] = idx
        outs = []
        for s in range(2):
            radius = self.radii[level + s]
            ns = SA_SAMPLES[s]
            bq = pn2.ball_query(radius, ns, xyz, new_xyz)
            self.trace['ball_idx_%d_%d' % (level, s)] = bq
            gxyz = pn2.group_gather_by_index(xyz.transpose(1, 2).contiguous(), bq)
            gxyz = gxyz - new_xyz.transpose(1, 2).unsqueeze(-1)
            g = gxyz if features is None else \
                torch.cat([gxyz, pn2.group_gather_by_index(features, bq)], dim=1)  # (B,3+C,M,ns)
            g = g.permute(0, 2, 1, 3).reshape(B * M, g.shape[1], ns)                # :397
            prefix = 'encoder.local_extract.set_abstractions.%d.pointnet_modules.%d.' % (level, s)
            f = self._sa_pointnet(g, prefix, 'sa%d_%d' % (level, s))                # :401
            outs.append(f.view(B, M, -1).transpose(1, 2))                           # :408
        return new_xyz, torch.cat(outs, dim=1)                                      # :414

    def feature_propagation(self, i, xyz, xyz_prev, features, features_prev):
        """pointnet2.py:483-525."""
        dist, idx = pn2.three_nn(xyz, xyz_prev)                                     # :514
        inv = 1.0 / (dist + 1e-8)                                                   # :516
        w = inv / torch.sum(inv, dim=2, keepdim=True)                               # :517-518
        new = pn2.three_interpolate(features_prev, idx, w)                          # :519
        if features is not None:                                                    # :521-523
            new = torch.cat([new, features], dim=1)
        p = 'encoder.local_extract.feature_propagators.%d.unit_pointnet.' % i
        new = self._relu(self._gn(self._conv(new, p + '0'), p + '1'), 'fp%d_r0' % i)
        new = self._relu(self._gn(self._conv(new, p + '3'), p + '4'), 'fp%d_r1' % i)
        return new

    def pointnet2(self, points):
        """pointnet2.py:217-249.  points (B',N,9) -> (B',N,512)."""
        xyz, features = pn2.separate_xyz_and_features(points)                       # :228
        xyz_list, feat_list = [xyz], [features]
        for level in range(5):                                                      # :232
            xyz, features = self.set_abstraction(level, xyz, features)
            xyz_list.append(xyz)
            feat_list.append(features)
        self.trace['sa_feat_4'] = feat_list[-1]
        for lvl in range(1, 6):
            self.trace['sa_out_%d' % lvl] = feat_list[lvl]
        ti = -2
        for i in range(5):                                                          # :238
            feat_list[ti] = self.feature_propagation(i, xyz_list[ti], xyz_list[ti + 1],
                                                     feat_list[ti], feat_list[ti + 1])
            self.trace['fp_out_%d' % i] = feat_list[ti]
            ti -= 1
        p = 'encoder.local_extract.final_layers.'
        x = self._relu(self._gn(self._conv(feat_list[0], p + '0'), p + '1'), 'final_r')   # :247
        x = self._conv(x, p + '3')
        return x.transpose(1, 2).contiguous()

    def encode(self, x):
        """tpointnet2.py:70-115 (via caspr.py:148-155).  x (B,T,N,4) -> z0 (B,1600), tnocs."""
        x = x.to(self.dtype)
        B, T, N, _ = x.shape
        g_in = x.view(B, T * N, 4).transpose(2, 1).contiguous()                     # :75
        g = self.pointnet_global(g_in)                                              # :76
        sp = x.view(B * T, N, 4)[:, :, :3]                                          # :79
        local_in = sp
        if self.augment_quad:
            local_in = torch.cat([sp, sp * sp], dim=2)                              # :83-84
        if self.augment_pairs:
            xz = sp[:, :, 0:1] * sp[:, :, 2:3]                                      # :87
            xy = sp[:, :, 0:1] * sp[:, :, 1:2]
            yz = sp[:, :, 2:3] * sp[:, :, 1:2]
            local_in = torch.cat([local_in, xz, xy, yz], dim=2)                     # :90
        lf = self.pointnet2(local_in).view(B, T * N, -1).transpose(2, 1).contiguous()   # :92-93
        self.trace['local_feat'] = lf
        feat = torch.cat([lf, g], dim=1)                                            # :96
        feat = self._relu(self._gn(self._conv(feat, 'encoder.conv1'), 'encoder.bn1'), 'head_r0')   # :99
        feat = self._gn(self._conv(feat, 'encoder.conv2'), 'encoder.bn2')           # :100
        tnocs = None
        if self.regress_tnocs:
            t_out = self._conv(self._relu(feat, 'head_r1'), 'encoder.conv3')        # :105
            tnocs = torch.sigmoid(t_out[:, :4, :]).transpose(2, 1).contiguous().view(B, T, N, 4)
        z0 = self._max(feat, 'head_max')                                            # :111
        return z0, tnocs

    # --------------------------------------------------------------- latent ODE
    def _dynamics(self, t, z):
        """latent_ode_model.py:139-147 (net built at :129-136: 4 Linear + 3 Tanh)."""
        self.nfe[0] += 1
        p = 'latent_ode.ode_func.dynamics_net.'
        h = torch.tanh(F.linear(z, self.sd[p + '0.weight'], self.sd[p + '0.bias']))
        h = torch.tanh(F.linear(h, self.sd[p + '2.weight'], self.sd[p + '2.bias']))
        h = torch.tanh(F.linear(h, self.sd[p + '4.weight'], self.sd[p + '4.bias']))
        return F.linear(h, self.sd[p + '6.weight'], self.sd[p + '6.bias'])

    def latent_ode(self, z0, t):
        """latent_ode_model.py:45-70; rtol=1e-3 and atol=rtol (the ':83' bug) -> 1e-3."""
        self.nfe[0] = 0                                                             # :55
        rel_t = t - t[0]                                                            # :58
        with torch.no_grad():
            pred = odeint001.odeint(self._dynamics, z0, rel_t, rtol=1e-3, atol=1e-3, method='dopri5')
        self.trace['latent_log'] = list(odeint001.LAST_SOLVER[0].log)
        return pred.permute(1, 0, 2)                                                # :68

    def aggregate_and_solve_latent(self, z0, time_tensor):
        """caspr.py:157-183."""
        B, T = time_tensor.shape
        solve_t, time_map = torch.unique(time_tensor, sorted=True, return_inverse=True)   # :166
        z_init = z0[:, :self.motion]                                                # :169
        z_global = z0[:, self.motion:]                                              # :170
        pred_z = self.latent_ode(z_init, solve_t)                                   # :173
        batch_inds = torch.arange(B).view((-1, 1)).repeat((1, T))                   # :175
        feats = pred_z[batch_inds, time_map, :]                                     # :177
        z_global = z_global.unsqueeze(1).expand(B, T, z_global.shape[1])            # :180
        return torch.cat([feats, z_global], dim=2)                                  # :181

    # ---------------------------------------------------------------------- CNF
    def _concatsquash(self, l, tc, x):
        """diffeq_layers.py:83-90."""
        p = 'point_cnf.chain.1.odefunc.diffeq.layers.%d.' % l
        gate = torch.sigmoid(F.linear(tc, self.sd[p + '_hyper_gate.weight'], self.sd[p + '_hyper_gate.bias']))
        bias = F.linear(tc, self.sd[p + '_hyper_bias.weight'])
        if x.dim() == 3:
            gate = gate.unsqueeze(1)
            bias = bias.unsqueeze(1)
        return F.linear(x, self.sd[p + '_layer.weight'], self.sd[p + '_layer.bias']) * gate + bias

    def odenet(self, tc, y):
        """odefunc.py:98-105: 4 ConcatSquashLinear, softplus (beta=1, threshold=20) between."""
        dx = y
        for l in range(4):
            dx = self._concatsquash(l, tc, dx)
            if l < 3:
                dx = F.softplus(dx)
        return dx

    def odefunc(self, t, states, e):
        """odefunc.py:119-142 with the Hutchinson VJP of odefunc.py:13-31."""
        y, _, c = states
        self.nfe[1] += 1                                                            # :122
        tt = torch.ones(y.size(0), 1).to(y) * t.clone().detach().type_as(y)         # :121
        with torch.enable_grad():
            y = y.detach().requires_grad_(True)
            tc = torch.cat([tt, c.view(y.size(0), -1)], dim=1)                      # :133
            dy = self.odenet(tc, y)                                                 # :134
            e_dzdx = torch.autograd.grad(dy, y, e)[0]                               # :14
            div = (e_dzdx * e).sum(dim=-1).unsqueeze(-1)                            # :15,:26,:135
        return dy.detach(), -div.detach(), torch.zeros_like(c)                      # :136

    def _mbn(self, i, x, logpx, reverse):
        """normalization.py:59-108 in eval mode (running stats, eps 1e-4)."""
        p = 'point_cnf.chain.%d.' % i
        mean = self.sd[p + 'running_mean'].view(1, -1).expand_as(x)
        var = self.sd[p + 'running_var'].view(1, -1).expand_as(x)
        weight = self.sd[p + 'weight'].view(1, -1).expand_as(x)
        bias = self.sd[p + 'bias'].view(1, -1).expand_as(x)
        logdet = (-0.5 * torch.log(var + 1e-4) + weight).sum(-1, keepdim=True)       # :103-108
        if not reverse:
            y = (x - mean) * torch.exp(-0.5 * torch.log(var + 1e-4))                # :70
            y = y * torch.exp(weight) + bias                                        # :74
            return y, (None if logpx is None else logpx - logdet)                   # :79
        y = (x - bias) * torch.exp(-weight)                                         # :92
        y = y * torch.exp(0.5 * torch.log(var + 1e-4)) + mean                       # :94
        return y, (None if logpx is None else logpx + logdet)                       # :99

    def cnf(self, x, context, logpx, reverse, e):
        """cnf.py:70-128 (eval mode): dopri5, atol=rtol=[1e-5]*3, times [0, sqrt_end_time^2]."""
        _logpx = torch.zeros(*x.shape[:-1], 1).to(x) if logpx is None else logpx    # :71-74
        s = self.sd['point_cnf.chain.1.sqrt_end_time']
        times = torch.stack([torch.tensor(0.0).to(x), s * s]).to(x)                 # :89-91
        if reverse:
            times = torch.flip(times, [0])                                          # :95-96
        self.nfe[1] = 0                                                             # :99
        func = lambda t, st: self.odefunc(t, st, e)
        with torch.no_grad():
            out = odeint001.odeint(func, (x, _logpx, context), times,
                                   atol=[1e-5] * 3, rtol=[1e-5] * 3, method='dopri5')
        self.trace['cnf_log'] = list(odeint001.LAST_SOLVER[0].log)
        self.trace['cnf_first_step'] = odeint001.LAST_SOLVER[0].first_step
        z_t, logpz_t = out[0][1], out[1][1]                                         # :121-124
        return z_t, (None if logpx is None else logpz_t)

    def point_cnf(self, x, context, logpx=None, reverse=False, e=None):
        """cnf.py:33-48: chain [MBN, CNF, MBN], reversed order when reverse."""
        if e is None:
            e = torch.randn_like(x)                                                 # odefunc.py:128
        order = [2, 1, 0] if reverse else [0, 1, 2]
        for i in order:
            if i == 1:
                x, logpx = self.cnf(x, context, logpx, reverse, e)
            else:
                x, logpx = self._mbn(i, x, logpx, reverse)
        return x, logpx

    # -------------------------------------------------------------- model level
    @staticmethod
    def standard_normal_logprob(z):
        """utils.py:10-12."""
        return -0.5 * log(2 * pi) - z.pow(2) / 2

    def decode(self, z, num_points=1024, constant_in_time=False, y=None, e=None):
        """caspr.py:204-267 (Gaussian base; `y`/`e` may be injected for parity tests)."""
        B, T, H = z.shape
        samp_batch = B if constant_in_time else B * T
        if y is None:
            y = torch.randn(samp_batch, num_points, 3).float()                      # utils.py:25
        if constant_in_time:
            y = y.view(B, 1, num_points, 3).expand(B, T, num_points, 3).reshape(B * T, num_points, 3)
        logp_y = self.standard_normal_logprob(y).view(B * T, num_points, -1).sum(2)  # :258
        x, _ = self.point_cnf(y, z.reshape(B * T, H), None, reverse=True, e=e)      # :262
        return (y.view(B, T, num_points, 3), logp_y.view(B, T, num_points),
                x.view(B, T, num_points, 3))

    def reconstruct(self, x, num_points=1024, constant_in_time=False, timestamps=None,
                    max_timestamp=5.0, y=None, e=None):
        """caspr.py:269-308."""
        B = x.shape[0]
        z0, tnocs = self.encode(x)                                                  # :297
        if timestamps is None:
            all_times = x[:, :, 0, 3] / max_timestamp                               # :300
        else:
            all_times = timestamps.view((1, -1)).repeat((B, 1))                     # :302
        z = self.aggregate_and_solve_latent(z0, all_times)                          # :305
        self.trace['z'] = z
        yy, logp_y, xr = self.decode(z, num_points, constant_in_time, y=y, e=e)     # :306
        return yy, logp_y, xr, tnocs

    def forward(self, x, sample_points, e=None):
        """caspr.py:76-146 (eval-mode numbers; no gradients)."""
        z0, tnocs = self.encode(x)                                                  # :85
        B, T, N, _ = sample_points.shape
        tnocs_loss = (tnocs[..., :4] - sample_points[..., :4]).abs() if self.regress_tnocs else None  # :93
        z = self.aggregate_and_solve_latent(z0, sample_points[:, :, 0, 3])          # :106-108
        pts = sample_points.reshape(B * T, N, 4)[:, :, :3].clone()                  # :113
        yy, dlogp = self.point_cnf(pts, z.reshape(B * T, -1), torch.zeros(B * T, N, 1), e=e)  # :116
        log_py = self.standard_normal_logprob(yy).sum(2)                            # :133-134
        nll = -(log_py - dlogp.view(B * T, N))                                      # :138-141
        return nll.view(B, T, -1), tnocs_loss

    def get_nfe(self):
        return list(self.nfe)


def chamfer_distance(pred, gt):
    """Symmetric mean squared-NN Chamfer distance per cloud (evaluations.py:40-43).

    pred (B,P,3), gt (B,Q,3) -> (B,)
    """
    d = torch.cdist(pred.double(), gt.double()) ** 2
    return (d.min(2)[0].mean(1) + d.min(1)[0].mean(1)).float()
