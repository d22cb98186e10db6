"""Pin the oracle restatement (oracle/caspr_oracle.py, pure CPU) against the fixtures frozen from the
UNMODIFIED reference modules (tests/golden/make_golden.py), and — when /root/reference is present —
against the live reference modules themselves."""
import os

import numpy as np
import pytest
import torch

from caspr_b200.synth import synthetic_state_dict, synthetic_sequences
from oracle.caspr_oracle import CasprOracle, chamfer_distance
from oracle import pointnet2_ops as pn2
from oracle.reference_loader import reference_available


@pytest.fixture(scope='module', params=['vig', 'def'])
def case(request, golden_dir):
    tag = request.param
    gold = dict(np.load(os.path.join(golden_dir, 'caspr_%s.npz' % tag)))
    sd = synthetic_state_dict(0, cnf_init='vigorous' if tag == 'vig' else 'default')
    x, nocs = synthetic_sequences(1, 3, 1024, seed=1)
    torch.set_num_threads(8)
    return tag, gold, CasprOracle(sd), x, nocs


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


def test_geometry_indices_bit_exact(case):
    _, gold, _, x, _ = case
    xyz = x.view(3, 1024, 4)[:, :, :3].contiguous()
    for lvl, m in enumerate([1024, 512, 256, 64, 16]):
        idx = pn2.furthest_point_sampling(xyz, m)
        assert np.array_equal(idx.numpy(), gold['fps_idx_%d' % lvl])
        new_xyz = pn2.fps_gather_by_index(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
        if lvl in (0, 2, 4):
            r = [0.02, 0.05, 0.1, 0.2, 0.4, 0.8][lvl + 1]
            assert np.array_equal(pn2.ball_query(r, 32, xyz, new_xyz).numpy(), gold['ball_idx_%d_1' % lvl])
        xyz = new_xyz


def test_encode_matches_reference_fixture(case):
    _, gold, oracle, x, _ = case
    z0, tnocs = oracle.encode(x)
    assert _rel(z0, gold['z0']) < 1e-5
    assert _rel(tnocs, gold['tnocs']) < 1e-5


def test_reconstruct_matches_reference_fixture(case):
    _, gold, oracle, x, _ = case
    y = torch.from_numpy(gold['rec_y']).reshape(3, 256, 3)
    e = torch.from_numpy(gold['rec_e'])
    _, _, xr, _ = oracle.reconstruct(x, num_points=256, y=y, e=e)
    assert list(oracle.get_nfe()) == list(gold['rec_nfe'])
    assert _rel(xr, gold['rec_x']) < 1e-5


def test_interpolated_reconstruct_matches_reference_fixture(case):
    _, gold, oracle, x, _ = case
    y = torch.from_numpy(gold['interp_y'])[:, 0]
    e = torch.from_numpy(gold['interp_e'])
    _, _, xr, _ = oracle.reconstruct(x, num_points=128, constant_in_time=True,
                                     timestamps=torch.linspace(0, 1, 5), y=y, e=e)
    assert list(oracle.get_nfe()) == list(gold['interp_nfe'])
    assert _rel(xr, gold['interp_x']) < 1e-5


def test_decode_config1_matches_reference_fixture(case):
    """BASELINE config 1: decode 512 points from one frozen latent on the CPU."""
    _, gold, oracle, _, _ = case
    z = torch.from_numpy(gold['dec_z'])
    _, _, xd = oracle.decode(z, num_points=512, y=torch.from_numpy(gold['dec_y']).reshape(1, 512, 3),
                             e=torch.from_numpy(gold['dec_e']))
    assert oracle.get_nfe()[1] == gold['dec_nfe'][1]
    assert _rel(xd, gold['dec_x']) < 1e-5


def test_forward_nll_matches_reference_fixture(case):
    _, gold, oracle, x, nocs = case
    nll, tl = oracle.forward(x, nocs, e=torch.from_numpy(gold['fwd_e']))
    assert list(oracle.get_nfe()) == list(gold['fwd_nfe'])
    assert _rel(nll, gold['fwd_nll']) < 1e-4
    assert abs(float(tl.mean()) - float(gold['fwd_tnocs_l1_mean'])) < 1e-6


def test_chamfer_definition():
    a = torch.tensor([[[0., 0, 0], [1, 0, 0]]])
    b = torch.tensor([[[0., 0, 0.5]]])
    cd = chamfer_distance(a, b)
    assert abs(float(cd) - ((0.25 + 1.25) / 2 + 0.25)) < 1e-6


@pytest.mark.skipif(not reference_available(), reason='/root/reference only exists in the dev container')
def test_oracle_matches_live_reference_modules():
    from oracle.reference_loader import build_reference_caspr
    sd = synthetic_state_dict(3)
    x, _ = synthetic_sequences(1, 2, 1024, seed=9)
    ref = build_reference_caspr()
    ref.load_state_dict(sd)
    ref.eval()
    with torch.no_grad():
        z0_ref, tn_ref = ref.encode(x)
    z0, tn = CasprOracle(sd).encode(x)
    assert _rel(z0, z0_ref) < 1e-5 and _rel(tn, tn_ref) < 1e-5


def test_emd_oracle_properties():
    """The approximate-EMD restatement (third-party algorithm, parity unpinned): exact on a permuted copy, the
    translation cost of a rigid shift, and agreement with the exact assignment on a small well-separated set."""
    from oracle.emd_oracle import approx_emd
    rng = np.random.default_rng(0)
    a = rng.random((2, 256, 3)) - 0.5
    perm = rng.permutation(256)
    assert approx_emd(a, a[:, perm]).max() / 256 < 1e-4
    shift = np.array([0.003, -0.002, 0.001])
    c = approx_emd(a, a[:, perm] + shift) / 256
    assert np.allclose(c, np.linalg.norm(shift), rtol=0.02)
    from scipy.optimize import linear_sum_assignment
    p, q = rng.random((1, 64, 3)), rng.random((1, 64, 3))
    d = np.linalg.norm(p[0][:, None] - q[0][None], axis=-1)
    r, cidx = linear_sum_assignment(d)
    exact = d[r, cidx].sum()
    approx = approx_emd(p, q)[0]
    assert exact <= approx * 1.001 and approx < 1.35 * exact


# ------------------------------------------------------------------ round-2 fixtures (tests/golden/make_golden_r2.py)
@pytest.fixture(scope='module')
def r2(golden_dir):
    torch.set_num_threads(8)
    return dict(np.load(os.path.join(golden_dir, 'caspr_r2.npz'))), CasprOracle(synthetic_state_dict(0, 'vigorous'))


def _seeded_y_e(seed, shape):
    torch.manual_seed(seed)
    return torch.randn(*shape), torch.randn(*shape)


def test_two_sequences_match_reference_fixture(r2):
    """B=2 with two different sequences: per-sequence head statistics, the unique / batch_inds scatter of
    caspr.py:166-177 and the batch-global controllers, against the unmodified reference modules."""
    gold, oracle = r2
    x, _ = synthetic_sequences(2, 10, 1024, seed=31)
    y, e = _seeded_y_e(15, (20, 512, 3))
    _, logp, xr, tn = oracle.reconstruct(x, num_points=512, y=y, e=e)
    assert list(oracle.get_nfe()) == list(gold['b2_nfe'])
    assert _rel(xr, gold['b2_x_rec']) < 1e-5
    assert _rel(logp, gold['b2_logp_y']) < 1e-6
    assert _rel(tn[:, ::3, ::8], gold['b2_tnocs_frame']) < 1e-5


def test_evaluation_call_and_demo_sequences_match_reference_fixture(r2):
    """utils/evaluations.py:105-114 (3 observed steps in, 10 query times) and two real demo sequences."""
    gold, oracle = r2
    x, nocs = synthetic_sequences(2, 10, 1024, seed=32)
    y, e = _seeded_y_e(16, (20, 256, 3))
    _, _, xr, _ = oracle.reconstruct(x[:, [0, 5, 9]], num_points=256, timestamps=nocs[0, :, 0, 3], y=y, e=e)
    assert list(oracle.get_nfe()) == list(gold['eval_nfe'])
    assert _rel(xr, gold['eval_x_rec']) < 1e-5
    y, e = _seeded_y_e(17, (10, 256, 3))
    _, _, xr, tn = oracle.reconstruct(torch.from_numpy(gold['demo_x']), num_points=256, y=y, e=e)
    assert list(oracle.get_nfe()) == list(gold['demo_nfe'])
    assert _rel(xr, gold['demo_x_rec']) < 1e-5
    assert _rel(tn[:, :, ::4], gold['demo_tnocs']) < 1e-5


def test_contour_and_truncated_samples_match_reference_fixture(r2):
    """The base-sample variants of decode (caspr.py:236-252): the fixture's contour / truncated points decode to the
    fixture's clouds, and the host-side samplers of the product reproduce the reference's RNG streams."""
    from caspr_b200.models.utils import sample_gaussian, sphere_surface_points
    gold, oracle = r2
    x, _ = synthetic_sequences(1, 3, 1024, seed=1)
    _, _, xr, _ = oracle.reconstruct(x, num_points=128, timestamps=torch.linspace(0, 1, 4), constant_in_time=True,
                                     y=torch.from_numpy(gold['cont_y'])[:, 0], e=torch.from_numpy(gold['cont_e']))
    assert list(oracle.get_nfe()) == list(gold['cont_nfe'])
    assert _rel(xr, gold['cont_x_rec']) < 1e-5
    _, _, xr, _ = oracle.reconstruct(x, num_points=128, y=torch.from_numpy(gold['trunc_y']).view(3, 128, 3),
                                     e=torch.from_numpy(gold['trunc_e']))
    assert list(oracle.get_nfe()) == list(gold['trunc_nfe'])
    assert _rel(xr, gold['trunc_x_rec']) < 1e-5
    torch.manual_seed(9)
    assert np.array_equal(sample_gaussian((3, 128, 3), 1.5).numpy(), gold['trunc_y'].reshape(3, 128, 3))
    np.random.seed(3)
    c = np.concatenate([sphere_surface_points(64, r).reshape(1, 64, 3) for r in (0.5, 1.0)], axis=1)
    assert np.array_equal(c[0].astype(np.float32), gold['cont_y'][0, 0])
