"""CPU restatement of the reference loader's array work (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows data/caspr_dataset.py:148-208 (``load_seq_path``) and :288-325 (``DynamicPCLDataset.__getitem__``) for frames
that are already decoded; pinned against the reference's own functions on demo frames by
tests/golden/make_golden_dataset.py -> tests/golden/dataset.npz."""
import numpy as np


def load_seq_arrays(nocs_frames, depth_frames, max_timestamp=5.0, expected_num_pts=4096):
    """load_seq_path (:148-208) without the file reads."""
    seq_len = len(nocs_frames)
    step_size = 0.0 if seq_len == 1 else 1.0 / (seq_len - 1)                       # :155-158
    nocs_seq = np.zeros((seq_len, expected_num_pts, 4))
    depth_seq = np.zeros((seq_len, expected_num_pts, 4))
    for step_idx, (nocs_pc, depth_pc) in enumerate(zip(nocs_frames, depth_frames)):
        if depth_pc.size == 0:
            depth_pc = nocs_pc                                                     # :174-176
        if np.count_nonzero(nocs_pc) == 0:
            break                                                                  # :183-186
        if nocs_pc.shape[0] < expected_num_pts:                                    # :188-195
            pad_size = expected_num_pts - nocs_pc.shape[0]
            while pad_size > 0:
                nocs_pc = np.concatenate([nocs_pc, nocs_pc[:pad_size].reshape((-1, 3))], axis=0)
                depth_pc = np.concatenate([depth_pc, depth_pc[:pad_size].reshape((-1, 3))], axis=0)
                pad_size = expected_num_pts - nocs_pc.shape[0]
        time_stamp = np.ones((nocs_pc.shape[0], 1)) * step_size * step_idx         # :200
        nocs_seq[step_idx] = np.concatenate([nocs_pc, time_stamp], axis=1)
        time_stamp = max_timestamp * np.ones((depth_pc.shape[0], 1)) * step_size * step_idx   # :204
        depth_seq[step_idx] = np.concatenate([depth_pc, time_stamp], axis=1)
    return nocs_seq, depth_seq


def get_item(nocs_seq, depth_seq, sampled_steps, sampled_pts, shift_time_to_zero=False):
    """__getitem__ (:288-325) for given step / point choices; sampled_pts (N,) or (T,N)."""
    sampled_steps = sorted(sampled_steps)
    sampled_pts = np.asarray(sampled_pts)
    if sampled_pts.ndim == 1:
        input_data = depth_seq[sampled_steps, :, :].copy()[:, sampled_pts, :]
        output_data = nocs_seq[sampled_steps, :, :].copy()[:, sampled_pts, :]
    else:
        time_inds = np.repeat(np.arange(sampled_pts.shape[0]), sampled_pts.shape[1])
        pt_inds = sampled_pts.reshape((-1))
        shape = (sampled_pts.shape[0], sampled_pts.shape[1], -1)
        input_data = depth_seq[sampled_steps, :, :].copy()[time_inds, pt_inds, :].reshape(shape)
        output_data = nocs_seq[sampled_steps, :, :].copy()[time_inds, pt_inds, :].reshape(shape)
    if shift_time_to_zero:
        input_data[:, :, -1] -= np.min(input_data[:, :, -1])
        output_data[:, :, -1] -= np.min(output_data[:, :, -1])
    return input_data.astype(np.float32), output_data.astype(np.float32)
