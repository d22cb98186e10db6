"""Seeded synthetic weights and input sequences for tests and benchmarks.

No pretrained weights or datasets are reachable (the reference ships download
scripts only: pretrained_weights/download_weights.sh, data/download_data.sh), so
every parity test and benchmark uses what this module generates.  Values depend
only on (seed, key) through the CPU generator, so the development container and
the GPU box produce identical tensors.

Input statistics follow the reference's demo sequences:
partial depth-camera views of an object at z in [1.1, 3.5] m, xy within +-0.3 m,
NOCS coordinates in [0,1]^3, world time stamps 5*i/(T-1), NOCS time stamps
i/(T-1) (reference caspr/data/caspr_dataset.py:158,200-205; cars.cfg max-timestamp 5.0).
"""
import json
import math
import os
import zlib

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
MANIFEST_PATH = os.path.join(_HERE, 'state_dict_manifest.json')


def load_manifest():
    """The reference's 238-key state_dict layout: {key: shape}."""
    with open(MANIFEST_PATH) as f:
        return json.load(f)


def _gen(seed, key):
    g = torch.Generator()
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


def synthetic_state_dict(seed=0, cnf_init='vigorous', manifest=None):
    """A full 238-key state_dict with seeded values.

    Encoder convs: PyTorch's default U(+-1/sqrt(fan_in)); GroupNorm affine perturbed
    away from (1,0) so the affine path is exercised; latent dynamics N(0,0.1) / zero
    bias (reference latent_ode_model.py:152-156).  CNF: ``'default'`` = PyTorch default
    init (near-trivial dynamics, NFE 14-20); ``'vigorous'`` = main weights
    N(0, sqrt(6/fan_in)), hyper weights N(0,0.05), sqrt_end_time=1 (NFE ~40).
    MovingBatchNorm buffers get trained-like statistics.
    """
    manifest = manifest or load_manifest()
    sd = {}
    for key, shape in manifest.items():
        g = _gen(seed, key)
        leaf = key.split('.')[-1]
        if leaf == '_num_evals':
            t = torch.tensor(0.)
        elif leaf == 'step':
            t = torch.zeros(1)
        elif leaf == 'sqrt_end_time':
            t = torch.tensor(1.0 if cnf_init == 'vigorous' else math.sqrt(0.5))
        elif key.startswith('point_cnf.chain.0.') or key.startswith('point_cnf.chain.2.'):
            data_side = key.startswith('point_cnf.chain.0.')
            if leaf == 'running_mean':
                t = torch.full(shape, 0.5 if data_side else 0.0) + 0.05 * (torch.rand(shape, generator=g) - 0.5)
            elif leaf == 'running_var':
                t = torch.full(shape, 0.04 if data_side else 1.0) * (0.75 + 0.5 * torch.rand(shape, generator=g))
            elif leaf == 'weight':
                t = 0.4 * (torch.rand(shape, generator=g) - 0.5)
            else:
                t = 0.2 * (torch.rand(shape, generator=g) - 0.5)
        elif key.startswith('latent_ode.'):
            if leaf == 'weight':
                # both aliases (ode_func / solver.ode_func) must hold the same tensor
                g = _gen(seed, key.replace('latent_ode.solver.', 'latent_ode.'))
                t = 0.1 * torch.randn(shape, generator=g)
            else:
                t = torch.zeros(shape)
        elif key.startswith('point_cnf.chain.1.odefunc.diffeq.'):
            fan_in = shape[1] if len(shape) == 2 else manifest[key[:-4] + 'weight'][1]
            if cnf_init == 'vigorous':
                if '_hyper' in key:
                    t = 0.05 * torch.randn(shape, generator=g)
                elif leaf == 'weight':
                    t = math.sqrt(6.0 / fan_in) * torch.randn(shape, generator=g)
                else:
                    t = 0.1 * torch.randn(shape, generator=g)
            else:
                b = 1.0 / math.sqrt(fan_in)
                t = (2 * torch.rand(shape, generator=g) - 1) * b
        elif len(shape) == 3:                       # Conv1d weight (out,in,1)
            b = 1.0 / math.sqrt(shape[1])
            t = (2 * torch.rand(shape, generator=g) - 1) * b
        else:                                       # 1-D: conv bias or GroupNorm affine
            wkey = key[:-len(leaf)] + 'weight'
            is_norm = len(manifest[wkey]) == 1
            if is_norm:
                t = (0.75 + 0.5 * torch.rand(shape, generator=g)) if leaf == 'weight' \
                    else 0.2 * (torch.rand(shape, generator=g) - 0.5)
            else:
                b = 1.0 / math.sqrt(manifest[wkey][1])
                t = (2 * torch.rand(shape, generator=g) - 1) * b
        sd[key] = t.to(torch.float32)
    return sd


def _rot(axis, angle):
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(angle) * K + (1 - math.cos(angle)) * (K @ K)


def synthetic_sequences(B, T, N, seed=0, warping=False, max_timestamp=5.0):
    """Seeded synthetic sequences.

    Returns (pcl_in (B,T,N,4) world-space points + world time stamp,
             nocs_out (B,T,N,4) T-NOCS twins of the same points).
    Rigid mode: a union of 2-4 random ellipsoids in NOCS space seen by a depth camera
    along a smooth rigid trajectory (camera-facing subset only).  ``warping=True``:
    the input is the NOCS cloud itself, deformed smoothly in time, correspondences
    kept across frames, max_timestamp 1.0 (reference data/configs/warping_cars.cfg).
    """
    rng = np.random.default_rng(seed)
    pcl = np.zeros((B, T, N, 4), dtype=np.float32)
    nocs = np.zeros((B, T, N, 4), dtype=np.float32)
    for b in range(B):
        ne = int(rng.integers(2, 5))
        centers = 0.5 + rng.uniform(-0.12, 0.12, size=(ne, 3))
        radii = rng.uniform(0.06, 0.22, size=(ne, 3))
        n_raw = 8 * N
        which = rng.integers(0, ne, size=n_raw)
        u = rng.normal(size=(n_raw, 3))
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        p = centers[which] + u * radii[which]
        # drop points that fall inside another ellipsoid (keeps a closed outer surface)
        inside = np.zeros(n_raw, dtype=bool)
        for j in range(ne):
            q = (p - centers[j]) / radii[j]
            inside |= ((q * q).sum(1) < 0.999) & (which != j)
        p = np.clip(p[~inside], 0.0, 1.0)
        nrm = (u / radii[which])[~inside]
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        axis = rng.normal(size=3)
        a0, a1 = rng.uniform(0, 2 * math.pi), rng.uniform(-1.0, 1.0)
        t0 = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(1.3, 3.2)])
        v = rng.uniform(-0.15, 0.15, size=3)
        amp, freq, ph = rng.uniform(0.02, 0.08, size=3), rng.uniform(1.0, 3.0, size=3), rng.uniform(0, 6.28, size=3)
        keep_w = rng.permutation(len(p))
        for i in range(T):
            s = i / max(T - 1, 1)
            if warping:
                d = amp * np.sin(freq * (p - 0.5) * 6.0 + ph + 2.0 * s)
                sel = keep_w[:N] if len(keep_w) >= N else np.resize(keep_w, N)
                pts = np.clip(p[sel] + d[sel] * s, 0.0, 1.0)
                pcl[b, i, :, :3] = pts
                pcl[b, i, :, 3] = s
                nocs[b, i, :, :3] = pts
                nocs[b, i, :, 3] = s
                continue
            R = _rot(axis, a0 + a1 * s)
            tt = t0 + v * s
            world = (p - 0.5) @ R.T + tt
            wn = nrm @ R.T
            facing = np.nonzero((wn * world).sum(1) < 0)[0]
            if len(facing) == 0:
                facing = np.arange(len(p))
            sel = facing[rng.permutation(len(facing))]
            sel = sel[:N] if len(sel) >= N else np.resize(sel, N)
            pcl[b, i, :, :3] = world[sel]
            pcl[b, i, :, 3] = max_timestamp * s
            nocs[b, i, :, :3] = p[sel]
            nocs[b, i, :, 3] = s
    return torch.from_numpy(pcl), torch.from_numpy(nocs)
