"""Spatio-temporal PointNet of TPointNet++ (mirror of reference caspr/models/pointnet.py:18-46).

Parameter containers are stock torch modules so the state_dict keys / shapes match the
reference (``conv{1,2,3}.{weight,bias}``, ``bn{1,2,3}.{weight,bias}``); the math runs in
libcaspr_b200.so on channels-last rows.
"""
import os

import torch
import torch.nn as nn

from .. import ops

NUM_GROUPS = 16      # pointnet.py:12
REDUCE_ONLY = os.environ.get('CASPR_POINTNET_REDUCE_ONLY', '1') != '0'     # conv3 as statistics + extrema only


class PointNetfeat(nn.Module):
    def __init__(self, input_dim=4, out_size=1024):
        super(PointNetfeat, self).__init__()
        self.conv1 = nn.Conv1d(input_dim, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, out_size, 1)
        self.bn1 = nn.GroupNorm(NUM_GROUPS, 64)
        self.bn2 = nn.GroupNorm(NUM_GROUPS, 128)
        self.bn3 = nn.GroupNorm(NUM_GROUPS, out_size)
        self.out_size = out_size

    def forward_rows(self, rows, samples, rows_per_sample, pointfeat_out=None, global_out=None):
        """rows (samples*rows_per_sample, input_dim) -> (global max (samples,out_size), pointfeat rows (.,64)).

        pointnet.py:36-42: conv1/GN/ReLU (kept as pointfeat), conv2/GN/ReLU, conv3/GN, max over points.
        `pointfeat_out` / `global_out` may be column slices of a wider concat buffer."""
        pf = ops.linear(rows, self.conv1.weight, self.conv1.bias, out=pointfeat_out)
        ops.groupnorm(pf, samples, rows_per_sample, NUM_GROUPS, self.bn1.weight, self.bn1.bias, relu=True)
        # only the max over the points of bn3(conv3(.)) is used (pointnet.py:40-42): the last GEMM keeps statistics and
        # per-channel extrema instead of writing its (rows x 1024) output, and the max-pool is read off them
        h3, st3 = ops.conv_gn_relu_conv(pf, self.conv2, self.bn2, self.conv3, samples, rows_per_sample, NUM_GROUPS,
                                        stats_b=True, reduce_only=REDUCE_ONLY)
        if global_out is None:
            global_out = torch.empty(samples, self.out_size, dtype=torch.float32, device=rows.device)
        if h3 is None:
            ops.gn_max_from_extrema(st3, samples, rows_per_sample, NUM_GROUPS, self.bn3.weight, self.bn3.bias,
                                    global_out, eps=self.bn3.eps)
        else:
            ops.groupnorm(h3, samples, rows_per_sample, NUM_GROUPS, self.bn3.weight, self.bn3.bias, relu=False,
                          write_back=False, maxout=global_out, stats=st3)
        return global_out, pf

    def forward(self, x):
        """Reference layout: x (B, input_dim, L) -> (B, out_size+64, L) = [global repeated | pointfeat]."""
        B, C, L = x.shape
        rows = x.transpose(1, 2).reshape(B * L, C).contiguous()
        out = torch.empty(B * L, self.out_size + 64, dtype=torch.float32, device=x.device)
        g, _ = self.forward_rows(rows, B, L, pointfeat_out=out[:, self.out_size:])
        ops.broadcast_rows(g, L, out[:, :self.out_size])
        return out.view(B, L, -1).transpose(1, 2)
