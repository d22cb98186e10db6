"""Input pipeline row: the loader oracle and the device batch assembly against fixtures produced by the
reference's own ``load_seq_path`` / ``DynamicPCLDataset.__getitem__`` (tests/golden/make_golden_dataset.py).
Float32 outputs of gathers and float64 time-stamp arithmetic: the bar is bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import dataset_oracle as do

COUNTS = 5
CASES = {
    # name: (fixture tag of the frames, steps, pts ('first' N or fixture key), shift, seq_len)
    'A': ('plain', [0, 1, 2, 3], ('first', 256), False),
    'B': ('plain', 'B_steps', 'B_pts', True),
    'C': ('plain', 'C_steps', 'C_pts', False),
    'D': ('blank', [0, 1, 2, 3, 4], ('first', 64), False),
    'E': ('nodepth', [0, 1, 2], ('first', 300), False),
}


@pytest.fixture(scope='module')
def gold(golden_dir):
    return dict(np.load(os.path.join(golden_dir, 'dataset.npz')))


def _frames(gold, tag):
    nocs = [gold['%s_nocs_%d' % (tag, i)] for i in range(COUNTS)]
    depth = [gold['%s_depth_%d' % (tag, i)] for i in range(COUNTS)]
    return nocs, depth


def _choices(gold, case):
    tag, steps, pts, shift = CASES[case]
    steps = gold[steps] if isinstance(steps, str) else np.asarray(steps)
    pts = gold[pts] if isinstance(pts, str) else np.arange(pts[1])
    return tag, steps, pts, shift


def test_load_seq_oracle_matches_reference(gold):
    nocs, depth = _frames(gold, 'plain')
    nocs_seq, depth_seq = do.load_seq_arrays(nocs, depth, 5.0, 512)
    assert np.array_equal(nocs_seq, gold['plain_nocs_seq']) and np.array_equal(depth_seq, gold['plain_depth_seq'])


@pytest.mark.parametrize('case', sorted(CASES))
def test_get_item_oracle_matches_reference(gold, case):
    tag, steps, pts, shift = _choices(gold, case)
    nocs, depth = _frames(gold, tag)
    nocs_seq, depth_seq = do.load_seq_arrays(nocs, depth, 5.0, 512)
    inp, out = do.get_item(nocs_seq, depth_seq, list(steps), pts, shift)
    assert np.array_equal(inp, gold[case + '_input']) and np.array_equal(out, gold[case + '_output'])


def test_read_sequence_host_decoding(gold, tmp_path):
    """caspr_b200.data.read_sequence (host side of the loader, no GPU): depth falls back to the NOCS cloud for a frame
    without depth data (caspr_dataset.py:174-176) and the first blank NOCS frame ends the valid range (:183-186)."""
    data = pytest.importorskip('caspr_b200.data')
    for tag, expect_valid in (('plain', 5), ('blank', 3), ('nodepth', 5)):
        nocs, depth = _frames(gold, tag)
        paths = []
        for i, (a, b) in enumerate(zip(nocs, depth)):
            p = tmp_path / ('%s_%d.npz' % (tag, i))
            np.savez(p, nocs_data=a, depth_data=b, obj_T=np.eye(4))
            paths.append(str(p))
        rn, rd, n_valid = data.read_sequence(paths)
        assert n_valid == expect_valid
        for i in range(COUNTS):
            assert np.array_equal(rn[i], nocs[i])
            assert np.array_equal(rd[i], nocs[i] if depth[i].size == 0 else depth[i])


@pytest.mark.gpu
def test_device_assembly_bit_exact(gold, tmp_path):
    """Two-sequence batches through read_sequence -> DeviceSequences.assemble, every case of the fixture."""
    from caspr_b200.data import read_sequence, DeviceSequences
    dev = torch.device('cuda', 0)
    seqs = {}
    for tag in ('plain', 'blank', 'nodepth'):
        nocs, depth = _frames(gold, tag)
        paths = []
        for i, (a, b) in enumerate(zip(nocs, depth)):
            p = tmp_path / ('%s_%d.npz' % (tag, i))
            np.savez(p, nocs_data=a, depth_data=b, obj_T=np.eye(4))
            paths.append(str(p))
        seqs[tag] = read_sequence(paths)
    assert seqs['blank'][2] == 3 and seqs['plain'][2] == 5
    for case in sorted(CASES):
        tag, steps, pts, shift = _choices(gold, case)
        other = 'plain' if tag != 'plain' else 'nodepth'
        batch = DeviceSequences([seqs[tag], seqs[other]], dev)
        steps2 = np.stack([steps, steps])
        pts2 = np.stack([pts, pts])
        inp, out = batch.assemble(steps2, pts2, max_timestamp=5.0, expected_num_pts=512, shift_time_to_zero=shift)
        assert np.array_equal(inp[0].cpu().numpy(), gold[case + '_input']), case
        assert np.array_equal(out[0].cpu().numpy(), gold[case + '_output']), case
        # second sequence of the batch against the oracle
        nocs, depth = _frames(gold, other)
        ns, ds = do.load_seq_arrays(nocs, depth, 5.0, 512)
        ri, ro = do.get_item(ns, ds, list(steps), pts, shift)
        assert np.array_equal(inp[1].cpu().numpy(), ri) and np.array_equal(out[1].cpu().numpy(), ro), case


@pytest.mark.gpu
def test_device_assembly_feeds_the_model(gold, tmp_path):
    """The assembled device batch goes straight into CaSPR.encode (no host round trip)."""
    from caspr_b200.data import DeviceSequences
    from caspr_b200.models import CaSPR
    from caspr_b200.synth import synthetic_state_dict
    nocs, depth = _frames(gold, 'plain')
    batch = DeviceSequences([(nocs, depth, 5)], torch.device('cuda', 0))
    x, _ = batch.assemble(np.arange(5)[None], np.arange(512)[None] % 512, expected_num_pts=512)
    x = torch.cat([x, x], dim=2)                         # 1024 points per frame
    model = CaSPR().cuda().eval()
    model.load_state_dict(synthetic_state_dict(0))
    z0, tnocs = model.encode(x)
    assert torch.isfinite(z0).all() and tnocs.shape == (1, 5, 1024, 4)
