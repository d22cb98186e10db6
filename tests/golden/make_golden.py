"""Freeze golden vectors from the UNMODIFIED reference modules (dev container only).

    python tests/golden/make_golden.py

Imports /root/reference/caspr/models over the oracle shims
(oracle/reference_loader.py), loads the seeded synthetic weights
(caspr_b200/synth.py) and writes small .npz fixtures next to this file.  The
fixtures pin (a) the oracle restatement (tests/test_oracle.py, CPU) and (b) the
CUDA path (tests/test_parity_gpu.py) on machines where /root/reference is absent.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.reference_loader import build_reference_caspr          # noqa: E402
from oracle import pointnet2_ops as pn2                            # noqa: E402
from caspr_b200.synth import synthetic_state_dict, synthetic_sequences   # noqa: E402


def main():
    torch.set_num_threads(8)
    for tag, cnf_init in (('vig', 'vigorous'), ('def', 'default')):
        sd = synthetic_state_dict(0, cnf_init=cnf_init)
        model = build_reference_caspr()
        model.load_state_dict(sd)
        model.eval()
        x, nocs = synthetic_sequences(1, 3, 1024, seed=1)
        out = {}
        with torch.no_grad():
            z0, tnocs = model.encode(x)
        out['z0'] = z0.numpy()
        out['tnocs'] = tnocs.numpy()
        # geometry of frame 0 through all five SA levels (same functions the shim binds)
        xyz = x.view(3, 1024, 4)[:, :, :3].contiguous()
        for lvl, m in enumerate([1024, 512, 256, 64, 16]):
            idx = pn2.furthest_point_sampling(xyz, m)
            out['fps_idx_%d' % lvl] = idx.numpy()
            new_xyz = pn2.fps_gather_by_index(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
            if lvl in (0, 2, 4):
                r = [0.02, 0.05, 0.1, 0.2, 0.4, 0.8][lvl + 1]
                out['ball_idx_%d_1' % lvl] = pn2.ball_query(r, 32, xyz, new_xyz).numpy()
            xyz = new_xyz
        torch.manual_seed(5)
        y, logp_y, xr, _ = model.reconstruct(x, num_points=256)
        out['rec_y'] = y.numpy()
        out['rec_logp_y'] = logp_y.numpy()
        out['rec_x'] = xr.detach().numpy()
        out['rec_nfe'] = np.asarray(model.get_nfe())
        # the Hutchinson noise the reference drew (torch.randn_like right after y)
        torch.manual_seed(5)
        _ = torch.randn(3, 256, 3)
        out['rec_e'] = torch.randn(3, 256, 3).numpy()
        # interpolated reconstruction (config 4 call shape): 5 query times, shared base sample
        torch.manual_seed(6)
        ts = torch.linspace(0, 1, 5)
        y2, _, xr2, _ = model.reconstruct(x, num_points=128, constant_in_time=True, timestamps=ts)
        out['interp_y'] = y2.numpy()
        out['interp_x'] = xr2.detach().numpy()
        out['interp_nfe'] = np.asarray(model.get_nfe())
        torch.manual_seed(6)
        _ = torch.randn(1, 128, 3)
        out['interp_e'] = torch.randn(5, 128, 3).numpy()
        # config 1: decode 512 points from one frozen latent
        g = torch.Generator().manual_seed(11)
        z = 0.8 * torch.randn(1, 1, 1600, generator=g)
        torch.manual_seed(9)
        y3, logp3, x3 = model.decode(z, num_points=512)
        out['dec_z'] = z.numpy()
        out['dec_y'] = y3.numpy()
        out['dec_x'] = x3.detach().numpy()
        out['dec_nfe'] = np.asarray(model.get_nfe())
        torch.manual_seed(9)
        _ = torch.randn(1, 512, 3)
        out['dec_e'] = torch.randn(1, 512, 3).numpy()
        # forward (NLL + T-NOCS L1), eval-mode numbers
        torch.manual_seed(7)
        nll, tl = model(x, nocs)
        out['fwd_nll'] = nll.detach().numpy()
        out['fwd_tnocs_l1_mean'] = np.asarray(tl.detach().mean().item())
        out['fwd_nfe'] = np.asarray(model.get_nfe())
        torch.manual_seed(7)
        out['fwd_e'] = torch.randn(3, 1024, 3).numpy()
        path = os.path.join(HERE, 'caspr_%s.npz' % tag)
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, 'KiB', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
