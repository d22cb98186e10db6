"""PointNet++ (MSG, GroupNorm) per-point feature network of TPointNet++.

Mirror of reference caspr/models/pointnet2.py: ``PointNet2feat`` (:14-249),
``PointNet2SetAbstraction`` (:254-419), ``PointNet2FeaturePropagator`` (:421-525) and
``PointNetFeatureExtractor`` (:528-708) — same constructor arguments, sub-module names and
parameter shapes (so ``set_abstractions.{i}.pointnet_modules.{s}.{conv_layers,bn_layers}.{l}``,
``feature_propagators.{i}.unit_pointnet.{0,1,3,4}`` and ``final_layers.{0,1,3}`` load unchanged).
Only the GroupNorm variant the reference instantiates (``batchnorm=False``) exists here.

Where the reference calls Kaolin's CUDA ops (furthest_point_sampling, ball query + group
gather, three_nn, three_interpolate; imports at pointnet2.py:7-10) and torch Conv1d/GroupNorm,
this file calls libcaspr_b200.so.  Activations stay channels-last ("rows x channels") from end to
end, which removes the transposes / contiguous() copies of pointnet2.py:386-387,:408-409.
"""
import os

import torch
import torch.nn as nn

from .. import ops

NUM_GROUPS = 16      # pointnet2.py:12
GATHER_IN_SPLIT = os.environ.get('CASPR_SA_GATHER_IN_SPLIT', '1') != '0'   # SA 3-5: group gather inside the operand split
DELAYED_FIRST_LAYER = os.environ.get('CASPR_SA_DELAYED', '1') != '0'       # SA 3-5: first layer before the gather


class PointNetFeatureExtractor(nn.Module):
    """Per-ball MLP of a set-abstraction scale: [Conv1d(k=1), GroupNorm, ReLU] x (L-1), Conv1d,
    GroupNorm (no ReLU, pointnet2.py:693), max over the ball (pointnet2.py:698)."""

    def __init__(self, in_channels=3, feat_size=1024, layer_dims=(64, 128), global_feat=True,
                 batchnorm=False, transposed_input=True):
        super(PointNetFeatureExtractor, self).__init__()
        if batchnorm:
            raise NotImplementedError('caspr_b200 implements the GroupNorm variant the reference uses')
        assert global_feat, 'set abstraction uses global_feat=True'
        dims = [in_channels] + list(layer_dims) + [feat_size]
        self.feat_size = feat_size
        self.conv_layers = nn.ModuleList()
        self.bn_layers = nn.ModuleList()
        for i in range(len(dims) - 1):
            self.conv_layers.append(nn.Conv1d(dims[i], dims[i + 1], 1))
            self.bn_layers.append(nn.GroupNorm(NUM_GROUPS, dims[i + 1]))

    def forward_rows(self, rows, balls, ns, out):
        """rows (balls*ns, C_in) grouped points -> out (balls, feat_size) view (may be a column slice)."""
        h = rows
        last = len(self.conv_layers) - 1
        if ns in (16, 32) and all(c.weight.shape[0] <= 64 for c in self.conv_layers):
            # small per-ball layers (SA levels 1-2): linear + per-ball GroupNorm + ReLU (+ max) in one kernel each
            for i, (conv, gn) in enumerate(zip(self.conv_layers, self.bn_layers)):
                if i < last:
                    h = ops.linear_gn_ball(h, conv.weight, conv.bias, gn.weight, gn.bias, ns, relu=True)
                else:
                    ops.linear_gn_ball(h, conv.weight, conv.bias, gn.weight, gn.bias, ns, relu=False,
                                       want_rows=False, maxout=out)
            return out
        widths = [c.weight.shape[0] for c in self.conv_layers]
        if ops.sa_mlp_tc_supported(ns, rows.shape[1], widths, rows.shape[0]):
            # SA levels 3-5: three tcgen05 GEMMs whose epilogues do the per-ball GroupNorm / ReLU / max
            return ops.sa_mlp_tc(rows, ns, self.conv_layers, self.bn_layers, out)
        for i, (conv, gn) in enumerate(zip(self.conv_layers, self.bn_layers)):
            h = ops.linear(h, conv.weight, conv.bias)
            if i < last:
                ops.groupnorm(h, balls, ns, NUM_GROUPS, gn.weight, gn.bias, relu=True)
            else:
                ops.groupnorm(h, balls, ns, NUM_GROUPS, gn.weight, gn.bias, relu=False, write_back=False,
                              maxout=out)
        return out


class PointNet2GroupingLayer(nn.Module):
    """Parameter-free placeholder keeping the reference's ``grouper_modules`` entries; the ball
    query of both scales runs as one fused scan in PointNet2SetAbstraction.forward_rows."""

    def __init__(self, radius, num_samples, use_xyz_feature=True, use_random_ball_query=False):
        super(PointNet2GroupingLayer, self).__init__()
        assert use_xyz_feature and not use_random_ball_query
        self.radius = radius
        self.num_samples = num_samples


class PointNet2SetAbstraction(nn.Module):
    def __init__(self, num_points_out, pointnet_in_features, pointnet_layer_dims_list, radii_list=None,
                 num_samples_list=None, batchnorm=False, use_xyz_feature=True, use_random_ball_query=False):
        super(PointNet2SetAbstraction, self).__init__()
        assert num_points_out is not None and len(radii_list) == len(num_samples_list) == \
            len(pointnet_layer_dims_list) == 2, 'the reference network uses two scales per level'
        self.num_points_out = num_points_out
        self.pointnet_layer_dims_list = pointnet_layer_dims_list
        self.grouper_modules = nn.ModuleList()
        self.pointnet_modules = nn.ModuleList()
        self.pointnet_in_channels = pointnet_in_features + (3 if use_xyz_feature else 0)
        for radius, ns, dims in zip(radii_list, num_samples_list, pointnet_layer_dims_list):
            self.grouper_modules.append(PointNet2GroupingLayer(radius, ns, use_xyz_feature, use_random_ball_query))
            self.pointnet_modules.append(PointNetFeatureExtractor(
                in_channels=self.pointnet_in_channels, feat_size=dims[-1], layer_dims=dims[:-1],
                global_feat=True, batchnorm=batchnorm, transposed_input=True))

    def get_num_features_out(self):
        return sum(d[-1] for d in self.pointnet_layer_dims_list)

    def forward_rows(self, xyz, features, trace=None):
        """xyz (B',N,3), features (B',N,C) channels-last view -> new_xyz (B',M,3), (B',M,sum D).

        pointnet2.py:384-414: FPS -> centres -> per scale ball query + grouping + per-ball MLP."""
        Bp = xyz.shape[0]
        M = self.num_points_out
        idx, new_xyz = ops.fps(xyz, M)
        g0, g1 = self.grouper_modules
        bq = ops.ball_query2(xyz, new_xyz, g0.radius, g0.num_samples, g1.radius, g1.num_samples)
        if trace is not None:
            trace['fps_idx'].append(idx)
            trace['ball_idx'].append(bq)
        out = torch.empty(Bp * M, self.get_num_features_out(), dtype=torch.float32, device=xyz.device)
        off = 0
        absmax = None
        for s, (grouper, pointnet) in enumerate(zip(self.grouper_modules, self.pointnet_modules)):
            widths = [c.weight.shape[0] for c in pointnet.conv_layers]
            if features is not None and ops.sa_mma_supported(grouper.num_samples, self.pointnet_in_channels, widths,
                                                             features):
                # levels 1-2 on the tensor cores: gather + three per-ball layers + max in one kernel, activations in
                # mma fragments; one operand bound per level serves both scales
                if absmax is None:
                    absmax = ops.sa_absmax(xyz, features)
                ops.sa_mma(xyz, new_xyz, features, bq[s], pointnet.conv_layers, pointnet.bn_layers,
                           out[:, off:off + pointnet.feat_size], absmax)
            elif ops.sa_fused_supported(grouper.num_samples, self.pointnet_in_channels, widths):
                # levels 1-2: gather + three per-ball layers + max in ONE kernel, activations stay in registers
                ops.sa_fused(xyz, new_xyz, features, bq[s], pointnet.conv_layers, pointnet.bn_layers,
                             out[:, off:off + pointnet.feat_size])
            elif (features is not None and GATHER_IN_SPLIT and
                  ops.sa_mlp_tc_supported(grouper.num_samples, self.pointnet_in_channels, widths,
                                          Bp * M * grouper.num_samples)):
                # levels 3-5: per-ball GroupNorm in the tcgen05 GEMM epilogues, the group gather folded into the operand
                # split of the first layer (the grouped tensor never exists)
                if DELAYED_FIRST_LAYER and widths[0] in (64, 128, 256):
                    # first layer's product once per source point, gathered afterwards (every point sits in ~24 balls)
                    ops.sa_mlp_tc_delayed(xyz, new_xyz, features, bq[s], pointnet.conv_layers, pointnet.bn_layers,
                                          out[:, off:off + pointnet.feat_size])
                else:
                    ops.sa_mlp_tc_grouped(xyz, new_xyz, features, bq[s], pointnet.conv_layers, pointnet.bn_layers,
                                          out[:, off:off + pointnet.feat_size])
            else:
                rows = ops.group_points(xyz, new_xyz, features, bq[s])
                pointnet.forward_rows(rows, Bp * M, grouper.num_samples, out[:, off:off + pointnet.feat_size])
            off += pointnet.feat_size
        if trace is not None:
            trace.setdefault('sa_out', []).append(out.view(Bp, M, -1))
        return new_xyz, out.view(Bp, M, -1)


class PointNet2FeaturePropagator(nn.Module):
    def __init__(self, num_features, num_features_prev, layer_dims, batchnorm=False):
        super(PointNet2FeaturePropagator, self).__init__()
        if batchnorm:
            raise NotImplementedError('caspr_b200 implements the GroupNorm variant the reference uses')
        self.layer_dims = layer_dims
        mods = []
        cin = num_features + num_features_prev
        for cout in layer_dims:
            mods += [nn.Conv1d(cin, cout, 1), nn.GroupNorm(NUM_GROUPS, cout), nn.ReLU()]
            cin = cout
        self.unit_pointnet = nn.Sequential(*mods)

    def get_num_features_out(self):
        return self.layer_dims[-1]

    def forward_rows(self, xyz, xyz_prev, features, features_prev):
        """pointnet2.py:514-525: three_nn, inverse-distance interpolation, skip concat, MLP.

        xyz (B',n,3), xyz_prev (B',m,3), features (B',n,Cs) view or None, features_prev (B',m,Cp)."""
        Bp, n, _ = xyz.shape
        dist, idx = ops.three_nn(xyz, xyz_prev)
        h = ops.three_interp_concat(features_prev, idx, dist, features)
        if len(self.unit_pointnet) == 6:
            conv_a, gn_a, _, conv_b, gn_b, _ = self.unit_pointnet
            h, st = ops.conv_gn_relu_conv(h, conv_a, gn_a, conv_b, Bp, n, NUM_GROUPS, stats_b=True)
            ops.groupnorm(h, Bp, n, NUM_GROUPS, gn_b.weight, gn_b.bias, relu=True, stats=st)
            return h.view(Bp, n, -1)
        for i in range(0, len(self.unit_pointnet), 3):
            conv, gn = self.unit_pointnet[i], self.unit_pointnet[i + 1]
            h = ops.linear(h, conv.weight, conv.bias)
            ops.groupnorm(h, Bp, n, NUM_GROUPS, gn.weight, gn.bias, relu=True)
        return h.view(Bp, n, -1)


class PointNet2feat(nn.Module):
    def __init__(self, in_features=0, num_classes=2, batchnorm=False, use_xyz_feature=True,
                 use_random_ball_query=False, radii_list=(0.02, 0.05, 0.1, 0.2, 0.4, 0.8),
                 max_feat_prop_size=512):
        super(PointNet2feat, self).__init__()
        if batchnorm:
            raise NotImplementedError('caspr_b200 implements the GroupNorm variant the reference uses')
        if len(radii_list) != 6:
            raise ValueError('Radii list must be length 6, not %d!' % len(radii_list))
        radii_list = list(radii_list)
        # (num_points_out, per-scale MLP widths) of pointnet2.py:64-146, GroupNorm branch
        spec = [(1024, [[16, 16, 32], [32, 32, 64]]),
                (512, [[32, 32, 64], [32, 32, 64]]),
                (256, [[64, 64, 128], [64, 96, 128]]),
                (64, [[128, 256, 256], [128, 256, 256]]),
                (16, [[256, 256, 512], [256, 256, 512]])]
        self.set_abstractions = nn.ModuleList()
        cin = in_features
        for lvl, (m, dims) in enumerate(spec):
            sa = PointNet2SetAbstraction(num_points_out=m, pointnet_in_features=cin,
                                         pointnet_layer_dims_list=dims,
                                         radii_list=[radii_list[lvl], radii_list[lvl + 1]],
                                         num_samples_list=[16, 32], batchnorm=batchnorm,
                                         use_xyz_feature=use_xyz_feature,
                                         use_random_ball_query=use_random_ball_query)
            self.set_abstractions.append(sa)
            cin = sa.get_num_features_out()

        self.feature_propagators = nn.ModuleList()
        sa_out = [sa.get_num_features_out() for sa in self.set_abstractions]
        prev = sa_out[-1]
        skips = [sa_out[-2], sa_out[-3], sa_out[-4], sa_out[-5], in_features]
        divisors = [1, 1, 2, 2, 4]                        # pointnet2.py:150,160,171,182,193
        layer_dims = None
        for skip, div in zip(skips, divisors):
            layer_dims = [max([max_feat_prop_size // div, num_classes])] * 2
            fp = PointNet2FeaturePropagator(num_features=skip, num_features_prev=prev,
                                            layer_dims=layer_dims, batchnorm=batchnorm)
            self.feature_propagators.append(fp)
            prev = fp.get_num_features_out()
        final_dim = layer_dims[0]
        self.final_layers = nn.Sequential(nn.Conv1d(prev, final_dim, 1), nn.GroupNorm(NUM_GROUPS, final_dim),
                                          nn.ReLU(), nn.Conv1d(final_dim, num_classes, 1))
        self.num_classes = num_classes

    def forward_rows(self, points, out=None, trace=None):
        """points (B',N,3+C) -> (B'*N, num_classes) rows (optionally written into `out`, a column
        slice of a wider buffer).  pointnet2.py:228-249."""
        Bp, N, cdim = points.shape
        flat = points.reshape(Bp * N, cdim)
        xyz = flat[:, :3].contiguous().view(Bp, N, 3)                      # separate_xyz_and_features
        features = points[:, :, 3:] if cdim > 3 else None                  # channels-last view
        xyz_list, feat_list = [xyz], [features]
        for sa in self.set_abstractions:
            xyz, features = sa.forward_rows(xyz, features, trace)
            xyz_list.append(xyz)
            feat_list.append(features)
        ti = -2
        for fp in self.feature_propagators:
            feat_list[ti] = fp.forward_rows(xyz_list[ti], xyz_list[ti + 1], feat_list[ti], feat_list[ti + 1])
            ti -= 1
        h = feat_list[0].reshape(Bp * N, -1)
        conv0, gn, _, conv1 = self.final_layers
        return ops.conv_gn_relu_conv(h, conv0, gn, conv1, Bp, N, NUM_GROUPS, out=out)

    def forward(self, points):
        """Reference layout: points (B',N,3+C) -> (B',N,num_classes)."""
        Bp, N, _ = points.shape
        return self.forward_rows(points.contiguous()).view(Bp, N, -1)
