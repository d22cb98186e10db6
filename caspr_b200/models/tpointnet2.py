"""TPointNet++ encoder (mirror of reference caspr/models/tpointnet2.py:8-123).

x (B,T,N,4) -> z0 (B,latent) and the sigmoid T-NOCS regression (B,T,N,4).  Same constructor,
sub-module names (``local_extract``, ``global_extract``, ``conv1..3``, ``bn1..2``) and
parameter shapes as the reference; all math in libcaspr_b200.so on channels-last rows, with
the [local | global-max | pointfeat] concat of tpointnet2.py:96 assembled in place.
"""
import torch
import torch.nn as nn

from .. import ops
from .. import _lib
from .pointnet import PointNetfeat
from .pointnet2 import PointNet2feat as PointNet2


class TPointNet2(nn.Module):
    def __init__(self, radii_list=[0.02, 0.05, 0.1, 0.2, 0.4, 0.8], local_feat_size=512, out_feat_size=1600,
                 augment_quad=True, augment_pairs=True, tnocs_point_size=4, regress_tnocs=True):
        super(TPointNet2, self).__init__()
        self.augment_quad = augment_quad
        self.augment_pairs = augment_pairs
        self.tnocs_point_size = tnocs_point_size
        self.local_feat_size = local_feat_size
        self.local_bottleneck_size = local_feat_size
        self.global_feat_size = 1024
        self.space_time_pt_feat = 64
        self.latent_feat_size = out_feat_size
        in_features = (3 if augment_quad else 0) + (3 if augment_pairs else 0)
        self.local_extract = PointNet2(in_features=in_features, num_classes=local_feat_size, batchnorm=False,
                                       use_xyz_feature=True, use_random_ball_query=False,
                                       radii_list=radii_list, max_feat_prop_size=self.local_bottleneck_size)
        self.global_extract = PointNetfeat(input_dim=4, out_size=self.global_feat_size)
        per_point = self.global_feat_size + self.space_time_pt_feat + self.local_feat_size
        self.conv1 = nn.Conv1d(per_point, per_point, 1)
        self.conv2 = nn.Conv1d(per_point, self.latent_feat_size, 1)
        self.bn1 = nn.GroupNorm(16, per_point)
        self.bn2 = nn.GroupNorm(16, self.latent_feat_size)
        self.regress_tnocs = regress_tnocs
        if regress_tnocs:
            self.conv3 = nn.Conv1d(self.latent_feat_size, tnocs_point_size, 1)
            self.loss_func = nn.L1Loss(reduction='none')
        self.trace = None          # set to a dict by parity tests to capture FPS / ball-query indices
        # The encoder is ~250 short kernel launches; replaying them from a CUDA graph removes the host
        # launch gaps.  One graph per (input shape, parameter storage); set use_cuda_graph=False to
        # launch eagerly.
        self.use_cuda_graph = True
        self._graphs = {}

    def _local_input(self, x4):
        """tpointnet2.py:79-90: xyz ++ (x^2,y^2,z^2) ++ (xz,xy,yz) as rows."""
        if self.augment_quad and self.augment_pairs:
            return ops.augment_xyz(x4)
        sp = x4[:, :3]
        parts = [sp]
        if self.augment_quad:
            parts.append(sp * sp)
        if self.augment_pairs:
            parts += [sp[:, 0:1] * sp[:, 2:3], sp[:, 0:1] * sp[:, 1:2], sp[:, 2:3] * sp[:, 1:2]]
        return torch.cat(parts, dim=1).contiguous()

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError('caspr_b200 runs on CUDA only (no CPU fallback): move the input to the GPU')
        x = x.to(torch.float32).contiguous()
        if self.training and torch.is_grad_enabled():
            # training: forward that keeps activations + hand-written backward (encoder_train.py)
            from .encoder_train import EncodeFunction
            params = [p for p in self.parameters() if p.requires_grad]
            z0, tnocs = EncodeFunction.apply(x, self, *params)
            return z0, (tnocs if self.regress_tnocs else None)
        if not self.use_cuda_graph or self.trace is not None:
            return self._forward_eager(x)
        key = (tuple(x.shape), x.device.index, self._param_key())
        entry = self._graphs.get(key)
        if entry is None:
            if len(self._graphs) >= 4:                  # bounded cache: graphs pin their activation pools
                self._graphs.clear()
            self._forward_eager(x)                      # lazy one-time initialisation happens outside capture
            torch.cuda.synchronize(x.device)
            x_static = x.clone()
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.lib.caspr_launch_count()
            with torch.cuda.graph(graph):
                z0_s, tnocs_s = self._forward_eager(x_static)
            entry = (graph, x_static, z0_s, tnocs_s, _lib.lib.caspr_launch_count() - n0)
            self._graphs[key] = entry
        graph, x_static, z0_s, tnocs_s, n_kernels = entry
        x_static.copy_(x)
        graph.replay()
        _lib.lib.caspr_launch_count_add(n_kernels)
        return z0_s.clone(), (None if tnocs_s is None else tnocs_s.clone())

    def _param_key(self):
        """Identity of the parameter storage the captured kernels read plus the in-place version counters
        (the graph references fp16 weight planes derived from the values: ops._prepared_weights)."""
        params = list(self.parameters())
        return (len(params), hash(tuple(p.data_ptr() for p in params)), sum(p._version for p in params))

    def _forward_eager(self, x):
        B, T, N, _ = x.shape
        R = B * T * N
        x4 = x.view(R, 4)
        L, G, Pf = self.local_feat_size, self.global_feat_size, self.space_time_pt_feat
        feat = torch.empty(R, L + G + Pf, dtype=torch.float32, device=x.device)
        # global spatio-temporal PointNet on (B, T*N) points (tpointnet2.py:75-76)
        gmax, _ = self.global_extract.forward_rows(x4, B, T * N, pointfeat_out=feat[:, L + G:])
        ops.broadcast_rows(gmax, T * N, feat[:, L:L + G])
        # per-frame PointNet++ (tpointnet2.py:79-93)
        local_in = self._local_input(x4)
        trace = None
        if self.trace is not None:
            trace = self.trace
            trace.setdefault('fps_idx', [])
            trace.setdefault('ball_idx', [])
        self.local_extract.forward_rows(local_in.view(B * T, N, -1), out=feat[:, :L], trace=trace)
        # head (tpointnet2.py:99-113)
        h2, st2 = ops.conv_gn_relu_conv(feat, self.conv1, self.bn1, self.conv2, B, T * N, 16,
                                        out=feat if self.latent_feat_size == feat.shape[1] else None, stats_b=True)
        z0 = torch.empty(B, self.latent_feat_size, dtype=torch.float32, device=x.device)
        ops.groupnorm(h2, B, T * N, 16, self.bn2.weight, self.bn2.bias, relu=False, write_back=self.regress_tnocs,
                      maxout=z0, stats=st2)
        tnocs = None
        if self.regress_tnocs:
            t = ops.linear(h2, self.conv3.weight, self.conv3.bias, act_in=ops.ACT_RELU, act_out=ops.ACT_SIGMOID)
            tnocs = t[:, :4].reshape(B, T, N, 4)
        return z0, tnocs

    def loss(self, outputs, gt):
        """tpointnet2.py:117-123: unreduced L1."""
        return self.loss_func(outputs, gt)
