"""CaSPR top-level model: TPointNet++ encoder -> Latent ODE -> conditional CNF decoder.

Drop-in mirror of reference caspr/models/caspr.py:22-308: identical constructor keywords,
``forward`` / ``encode`` / ``aggregate_and_solve_latent`` / ``gen_latent`` / ``get_nfe`` /
``decode`` / ``reconstruct`` signatures and return tuples, identical sub-module names
(``encoder``, ``latent_ode``, ``point_cnf``) and therefore the identical 238-key state_dict.
All heavy work runs in libcaspr_b200.so; this file only sequences the calls.

Differences a caller can observe: CUDA only, and two optional keyword arguments
(``y`` and ``e``) to inject the base samples / Hutchinson noise for parity tests.
"""
import numpy as np
import torch
import torch.nn as nn

from .tpointnet2 import TPointNet2
from .latent_ode_model import LatentODE
from .flow import get_point_cnf, count_nfe, PointCNFArgs
from .utils import standard_normal_logprob, sample_gaussian, sphere_surface_points


class CaSPR(nn.Module):
    def __init__(self, radii_list=[0.02, 0.05, 0.1, 0.2, 0.4, 0.8], local_feat_size=512, latent_feat_size=1600,
                 ode_hidden_size=512, motion_feat_size=64, pretrain_tnocs=False, augment_quad=True,
                 augment_pairs=True, cnf_blocks=1, regress_tnocs=True):
        super(CaSPR, self).__init__()
        self.pretrain_tnocs = pretrain_tnocs
        self.augment_quad = augment_quad
        self.augment_pairs = augment_pairs
        self.motion_feat_size = motion_feat_size
        self.regress_tnocs = regress_tnocs
        self.tnocs_point_size = 4
        self.encoder = TPointNet2(radii_list, local_feat_size=local_feat_size, out_feat_size=latent_feat_size,
                                  augment_quad=augment_quad, augment_pairs=augment_pairs,
                                  tnocs_point_size=self.tnocs_point_size, regress_tnocs=regress_tnocs)
        if self.pretrain_tnocs:
            return
        self.latent_ode = LatentODE(input_size=motion_feat_size, hidden_size=ode_hidden_size, num_layers=2,
                                    nonlinearity=nn.Tanh)
        self.cnf_args = PointCNFArgs()
        self.cnf_args.zdim = latent_feat_size
        self.cnf_args.num_blocks = cnf_blocks
        self.point_cnf = get_point_cnf(self.cnf_args)

    # ------------------------------------------------------------------------------ forward
    def forward(self, x, sample_points, aggregate_points=None, e=None):
        """caspr.py:76-122: -> (nll (B,T,N), tnocs_l1 (B,T,N,4)) or (tnocs_l1,) when pretraining.

        In training mode with autograd enabled the result is differentiable: ``loss.backward()``
        (train_utils.py:173) runs the hand-written encoder backward and the CUDA adjoint solves."""
        if not (self.training and torch.is_grad_enabled()):
            with torch.no_grad():
                return self._forward(x, sample_points, e)
        return self._forward(x, sample_points, e)

    def _forward(self, x, sample_points, e=None):
        z0, tnocs_pred = self.encode(x)
        B, H = z0.size()
        _, T, N, _ = sample_points.size()
        tnocs_loss = None
        if self.regress_tnocs:
            tnocs_loss = self.encoder.loss(tnocs_pred[:, :, :, :self.tnocs_point_size],
                                           sample_points[:, :, :, :self.tnocs_point_size].to(tnocs_pred))
        if self.pretrain_tnocs:
            return tuple([tnocs_loss])
        all_times = sample_points[:, :, 0, 3]
        z = self.aggregate_and_solve_latent(z0, all_times).reshape(B * T, self.cnf_args.zdim)
        pts = sample_points.reshape(B * T, N, 4)[:, :, :3].to(z0).contiguous()
        init_logprob = torch.zeros(B * T, N, 1, dtype=torch.float32, device=pts.device)
        cnf_result = self.point_cnf(pts, z, init_logprob, e=e)
        return tuple([self.get_nll_loss(cnf_result, B, T), tnocs_loss])

    def get_nll_loss(self, cnf_result_list, B, T):
        """caspr.py:124-146."""
        y, delta_log_py = cnf_result_list
        n = y.size(1)
        log_py = standard_normal_logprob(y).sum(2)
        log_px = log_py - delta_log_py.view(B * T, n)
        return (-log_px).view(B, T, -1)

    def encode(self, x):
        return self.encoder(x)

    def aggregate_and_solve_latent(self, z0, time_tensor):
        """caspr.py:157-183: solve the latent ODE at the unique time stamps and scatter back."""
        B, T = time_tensor.size()
        solve_t, time_map = torch.unique(time_tensor, sorted=True, return_inverse=True)
        z_init = z0[:, :self.latent_ode.input_size]
        z_global = z0[:, self.latent_ode.input_size:]
        pred_z = self.gen_latent(z_init, solve_t)
        batch_inds = torch.arange(B, device=pred_z.device).view(-1, 1).repeat(1, T)
        sample_feats = pred_z[batch_inds, time_map.to(pred_z.device), :]
        z_global = z_global.unsqueeze(1).expand(B, T, z_global.size(1))
        return torch.cat([sample_feats, z_global], dim=2)

    lockstep_group = False      # set by sharding.lockstep(): ranks of this group share one latent step sequence

    def gen_latent(self, z0, timestamps):
        if self.lockstep_group is False:
            return self.latent_ode(z0, timestamps)
        # lock-step over a sharded batch: the latent controller is batch-global, the state is tiny (B x 64), so
        # every rank solves the FULL batch (all-gather of z0) and keeps its own rows: identical step decisions
        import torch.distributed as dist
        from ..sharding import gather_rows
        full, lo, hi = gather_rows(z0.contiguous(), self.lockstep_group)
        return self.latent_ode(full, timestamps)[lo:hi]

    def get_nfe(self):
        return np.array([count_nfe(self.latent_ode), count_nfe(self.point_cnf)])

    # ------------------------------------------------------------------------------- decode
    @torch.no_grad()
    def decode(self, z, num_points=1024, constant_in_time=False, truncate_std=None, sample_contours=None,
               y=None, e=None):
        """caspr.py:204-267: -> (y (B,T,P,3), logp_y (B,T,P), x (B,T,P,3))."""
        B, T, H = z.size()
        samp_batch = B if constant_in_time else B * T
        input_dim = self.cnf_args.input_dim
        samp_size = (samp_batch, num_points, input_dim)
        if y is not None:
            y = y.to(z).reshape(samp_size)
        elif sample_contours is not None:
            radii = sample_contours
            contours, nsamp = [], 0
            for radius in radii:
                last = radius == radii[-1]
                per = (num_points - nsamp) if last else (num_points // len(radii))
                pts = sphere_surface_points(samp_batch * per, radius=radius).reshape((samp_batch, per, 3))
                contours.append(pts)
                nsamp += num_points // len(radii)
            y = torch.from_numpy(np.concatenate(contours, axis=1)).to(z).view(samp_size)
        else:
            y = sample_gaussian(samp_size, truncate_std, device=z.device)
        if constant_in_time:
            y = y.view(B, 1, num_points, input_dim).expand(B, T, num_points, input_dim)
            y = y.reshape(B * T, num_points, input_dim)
        logp_y = standard_normal_logprob(y).view(B * T, num_points, -1).sum(2)
        x = self.point_cnf(y, z.reshape(B * T, H), reverse=True, e=e)
        return (y.view(B, T, num_points, input_dim), logp_y.view(B, T, num_points),
                x.view(B, T, num_points, input_dim))

    @torch.no_grad()
    def reconstruct(self, x, num_points=1024, constant_in_time=False, timestamps=None, max_timestamp=5.0,
                    truncate_std=None, sample_contours=None, y=None, e=None):
        """caspr.py:269-308: -> (y, logp_y, x_rec, tnocs_pred)."""
        B, T, N, _ = x.size()
        z0, tnocs_pred = self.encode(x)
        if y is None and sample_contours is None:
            # The base samples come from the CPU generator (models/utils.py:25).  The encoder kernels have only been
            # enqueued at this point and nothing else touches the generator before decode, so drawing here yields the
            # reference's values while the GPU is busy encoding instead of idling during the draw.
            Tq = T if timestamps is None else int(timestamps.numel())
            samp_batch = B if constant_in_time else B * Tq
            y = sample_gaussian((samp_batch, num_points, self.cnf_args.input_dim), truncate_std, device=x.device)
            truncate_std = None
        if timestamps is None:
            all_times = x[:, :, 0, 3] / max_timestamp
        else:
            all_times = timestamps.view((1, -1)).repeat((B, 1))
        z = self.aggregate_and_solve_latent(z0, all_times.to(z0))
        y, logp_y, xr = self.decode(z, num_points, constant_in_time, truncate_std, sample_contours, y=y, e=e)
        return y, logp_y, xr, tnocs_pred
