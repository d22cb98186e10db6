"""Input pipeline -> device: the batch-assembly half of the reference's loader on the GPU.

Mirrors ``data/caspr_dataset.py``: ``load_seq_path`` (:148-208) decodes the ``.npz`` frames of a sequence, pads short
frames by cycling their points, stops at a blank frame and appends the NOCS / world time stamps;
``DynamicPCLDataset.__getitem__`` (:288-325) picks time steps and points and casts to float32.  Here the host only
decodes the files (``read_sequence``) and uploads the raw float64 points once; everything after that is ONE kernel
for the whole batch (``caspr_assemble_batch``), so the numpy work per item (allocate, concatenate, fancy-index, cast)
and the (B,T,N,4) host-to-device copy of the float32 batch disappear from the step.
"""
import ctypes

import numpy as np
import torch

from ._lib import lib, check
from .ops import _p, _stream, _count

DEFAULT_MAX_TIMESTAMP = 5.0          # caspr_dataset.py:16
DEFAULT_EXPECTED_NUM_PTS = 4096      # caspr_dataset.py:20


def read_sequence(frame_paths):
    """Decode the frames of one sequence (host).  Returns (nocs list, depth list, n_valid): float64 (n_i,3) arrays per
    frame, depth falling back to the NOCS cloud when a frame has none (:174-176), n_valid = frames before the first
    blank NOCS frame (:183-186)."""
    nocs, depth, n_valid = [], [], None
    for i, path in enumerate(frame_paths):
        with np.load(path) as d:
            n_pc = np.asarray(d['nocs_data'], dtype=np.float64).reshape(-1, 3)
            d_pc = np.asarray(d['depth_data'], dtype=np.float64)
        d_pc = n_pc if d_pc.size == 0 else d_pc.reshape(-1, 3)
        if n_valid is None and np.count_nonzero(n_pc) == 0:
            n_valid = i
        nocs.append(n_pc)
        depth.append(d_pc)
    return nocs, depth, len(frame_paths) if n_valid is None else n_valid


class DeviceSequences(object):
    """Raw frames of B sequences (each with the same number of frames) resident on the device."""

    def __init__(self, sequences, device):
        """sequences: list of (nocs list, depth list, n_valid) as returned by ``read_sequence``."""
        self.B = len(sequences)
        self.Tfull = len(sequences[0][0])
        counts = []
        for nocs, depth, _ in sequences:
            assert len(nocs) == self.Tfull == len(depth)
            for a, b in zip(nocs, depth):
                assert a.shape == b.shape and a.shape[0] > 0
                counts.append(a.shape[0])
        off = np.zeros(len(counts) + 1, dtype=np.int64)
        np.cumsum(counts, out=off[1:])
        nocs_all = torch.from_numpy(np.concatenate([a for s in sequences for a in s[0]], axis=0)).pin_memory()
        depth_all = torch.from_numpy(np.concatenate([a for s in sequences for a in s[1]], axis=0)).pin_memory()
        self.nocs = nocs_all.to(device, non_blocking=True)
        self.depth = depth_all.to(device, non_blocking=True)
        self.frame_off = torch.from_numpy(off).to(device)
        self.n_valid = torch.tensor([s[2] for s in sequences], dtype=torch.int32, device=device)
        self.min_count = int(min(counts))
        self.device = device

    def assemble(self, steps, pts, max_timestamp=DEFAULT_MAX_TIMESTAMP, expected_num_pts=DEFAULT_EXPECTED_NUM_PTS,
                 shift_time_to_zero=False):
        """steps (B,T) sorted time-step indices; pts (B,N) or (B,T,N) point indices (< expected_num_pts).
        -> (input (B,T,N,4), output (B,T,N,4)) float32 on the device, as ``__getitem__`` stacks them."""
        steps = torch.as_tensor(steps, dtype=torch.int32, device=self.device).contiguous()
        pts = torch.as_tensor(pts, dtype=torch.int32, device=self.device).contiguous()
        B, T = steps.shape
        assert B == self.B
        if pts.dim() == 2:
            pts = pts.unsqueeze(1)
        Tp, N = pts.shape[1], pts.shape[2]
        inp = torch.empty(B, T, N, 4, dtype=torch.float32, device=self.device)
        out = torch.empty(B, T, N, 4, dtype=torch.float32, device=self.device)
        _count('assemble_batch')
        check(lib.caspr_assemble_batch(_p(self.nocs), _p(self.depth), _p(self.frame_off), _p(self.n_valid), B,
                                       self.Tfull, int(expected_num_pts), _p(steps), T, _p(pts), Tp, N,
                                       ctypes.c_double(max_timestamp), int(shift_time_to_zero), _p(inp), _p(out),
                                       _stream()), 'caspr_assemble_batch')
        return inp, out
