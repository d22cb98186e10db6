"""Freeze loader fixtures from the UNMODIFIED reference loader (dev container only).

    python tests/golden/make_golden_dataset.py

Runs the reference's own ``load_seq_path`` and ``DynamicPCLDataset.__getitem__`` (data/caspr_dataset.py) on demo frames
of /root/reference/data/demo, truncated to a few hundred points so that padding is exercised, and stores the decoded
inputs together with the outputs in tests/golden/dataset.npz.  torchvision (imported but unused by the module) is
stubbed; the dataset object is created without its directory-scanning constructor."""
import glob
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


def main():
    sys.modules.setdefault('torchvision', types.ModuleType('torchvision'))
    sys.modules['torchvision'].transforms = types.ModuleType('transforms')
    sys.modules['torchvision'].utils = types.ModuleType('utils')
    sys.path.insert(0, os.path.join(REF, 'caspr'))
    from data import caspr_dataset as cd
    frames = sorted(glob.glob(os.path.join(REF, 'data/demo/*/seq_00000000/frame_*.npz')))[:5]
    counts = [300, 512, 130, 512, 401]
    expected = 512
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        def write(tag, blank=None, no_depth=None):
            paths = []
            for i, (f, c) in enumerate(zip(frames, counts)):
                d = np.load(f)
                nocs, depth = d['nocs_data'][:c], d['depth_data'][:c]
                if blank == i:
                    nocs = np.zeros_like(nocs)
                if no_depth == i:
                    depth = np.zeros((0,))
                p = os.path.join(tmp, 'm', tag, 'frame_%08d.npz' % i)
                os.makedirs(os.path.dirname(p), exist_ok=True)
                np.savez(p, nocs_data=nocs, depth_data=depth, obj_T=d['obj_T'])
                out['%s_nocs_%d' % (tag, i)] = nocs
                out['%s_depth_%d' % (tag, i)] = depth
                paths.append(p)
            return paths

        def item(paths, seed, **attrs):
            ds = object.__new__(cd.DynamicPCLDataset)
            ds.seq_data_paths = [paths]
            ds.max_timestamp, ds.expected_num_pts = 5.0, expected
            ds.return_pose_data = False
            for k, v in attrs.items():
                setattr(ds, k, v)
            np.random.seed(seed)
            (inp, outp), _, _ = ds[0]
            return inp.numpy(), outp.numpy()

        base = write('plain')
        nocs_seq, depth_seq, _ = cd.load_seq_path(base, max_timestamp=5.0, expected_num_pts=expected)
        out['plain_nocs_seq'], out['plain_depth_seq'] = nocs_seq, depth_seq
        # A: test.py configuration (first steps, first points)
        out['A_input'], out['A_output'] = item(base, 0, return_first_steps=True, seq_len=4, random_point_sample=False,
                                               random_point_sample_per_step=False, num_pts=256, shift_time_to_zero=False)
        # B: random steps and one random point set, shifted time
        out['B_input'], out['B_output'] = item(base, 1, return_first_steps=False, seq_len=3, random_point_sample=True,
                                               random_point_sample_per_step=False, num_pts=200, shift_time_to_zero=True)
        np.random.seed(1)
        out['B_steps'] = np.asarray(sorted(np.random.choice(5, 3, replace=False)))
        out['B_pts'] = np.random.choice(expected, 200, replace=False)
        # C: per-step point sets (needs seq_len == number of frames)
        out['C_input'], out['C_output'] = item(base, 2, return_first_steps=False, seq_len=5, random_point_sample=False,
                                               random_point_sample_per_step=True, num_pts=100, shift_time_to_zero=False)
        np.random.seed(2)
        out['C_steps'] = np.asarray(sorted(np.random.choice(5, 5, replace=False)))
        out['C_pts'] = np.stack([np.random.choice(expected, 100, replace=False) for _ in range(5)], axis=0)
        # D: blank frame 3 stops the sequence (remaining frames stay zero); E: frame 1 without depth data
        d_paths = write('blank', blank=3)
        out['D_input'], out['D_output'] = item(d_paths, 0, return_first_steps=True, seq_len=5, random_point_sample=False,
                                               random_point_sample_per_step=False, num_pts=64, shift_time_to_zero=False)
        e_paths = write('nodepth', no_depth=1)
        out['E_input'], out['E_output'] = item(e_paths, 0, return_first_steps=True, seq_len=3, random_point_sample=False,
                                               random_point_sample_per_step=False, num_pts=300, shift_time_to_zero=False)
    path = os.path.join(HERE, 'dataset.npz')
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()
