"""GPU parity at the shapes the benchmark and the reference's callers actually use (round 2).

Round 1 proved model-level parity on one sequence only.  Here: two different sequences per batch, the full BASELINE
config-2 shape (B=8, T=10, N=1024, P=2048) against the oracle run on the box's host cores, the evaluation call of
utils/evaluations.py:105-114, two real demo sequences, the reference's CPU RNG stream for the base samples, and the
`truncate_std` / `sample_contours` branches of `decode` (caspr.py:236-252).  Tolerance: 1e-4 relative on coordinates
(north_star), identical NFE.  Fixtures: tests/golden/caspr_r2.npz (tests/golden/make_golden_r2.py, frozen from the
unmodified reference modules)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from caspr_b200.synth import synthetic_state_dict, synthetic_sequences   # noqa: E402
from oracle.caspr_oracle import CasprOracle, chamfer_distance             # noqa: E402

DEV = 'cuda:0'


def _rel(a, b):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


@pytest.fixture(scope='module')
def setup(golden_dir, lib_built):
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from caspr_b200.models import CaSPR
    gold = dict(np.load(os.path.join(golden_dir, 'caspr_r2.npz')))
    gold1 = dict(np.load(os.path.join(golden_dir, 'caspr_vig.npz')))
    sd = synthetic_state_dict(0, cnf_init='vigorous')
    model = CaSPR().to(DEV).eval()
    model.load_state_dict(sd)
    return gold, gold1, model, sd


def _seeded_y_e(seed, shape):
    """The reference's draws under torch.manual_seed(seed): base samples (models/utils.py:25), then the Hutchinson
    noise (odefunc.py:128), both from the CPU generator when the reference runs on the CPU."""
    torch.manual_seed(seed)
    return torch.randn(*shape), torch.randn(*shape)


def test_two_sequences_match_reference_fixture(setup):
    """B=2, T=10, N=1024, P=512 with two different sequences: per-sequence head GroupNorm and max-pool, the
    torch.unique / batch_inds scatter (caspr.py:166-177), batch-global controllers over 20 frames."""
    gold, _, model, _ = setup
    x, _ = synthetic_sequences(2, 10, 1024, seed=31)
    y, e = _seeded_y_e(15, (20, 512, 3))
    z0, _ = model.encode(x.to(DEV))
    assert _rel(z0, gold['b2_z0']) < 1e-4
    yy, logp, xr, tn = model.reconstruct(x.to(DEV), num_points=512, y=y, e=e.to(DEV))
    assert list(model.get_nfe().astype(int)) == list(gold['b2_nfe'].astype(int))
    assert _rel(xr, gold['b2_x_rec']) < 1e-4
    assert _rel(logp, gold['b2_logp_y']) < 1e-5
    assert _rel(tn[:, ::3, ::8], gold['b2_tnocs_frame']) < 1e-4
    # the two sequences really went through different per-sequence statistics
    assert float((z0[0] - z0[1]).abs().max()) > 1e-3


def test_cpu_rng_stream_reproduces_reference_samples(setup):
    """Without injected samples `reconstruct` draws the base points exactly as the reference does (CPU generator,
    models/utils.py:25): under torch.manual_seed(5) it returns the fixture's `rec_y` bit for bit."""
    _, gold1, model, _ = setup
    x, _ = synthetic_sequences(1, 3, 1024, seed=1)
    torch.manual_seed(5)
    yy, logp, xr, _ = model.reconstruct(x.to(DEV), num_points=256, e=torch.from_numpy(gold1['rec_e']).to(DEV))
    assert np.array_equal(yy.cpu().numpy(), gold1['rec_y'])
    assert list(model.get_nfe().astype(int)) == list(gold1['rec_nfe'].astype(int))
    assert _rel(xr, gold1['rec_x']) < 1e-4


def test_evaluation_protocol_call_shape(setup):
    """utils/evaluations.py:105-114: observed steps [0,5,9] in, all ten NOCS time stamps queried."""
    gold, _, model, _ = setup
    x, nocs = synthetic_sequences(2, 10, 1024, seed=32)
    y, e = _seeded_y_e(16, (20, 256, 3))
    obs = x[:, [0, 5, 9]].contiguous().to(DEV)
    _, _, xr, _ = model.reconstruct(obs, num_points=256, timestamps=nocs[0, :, 0, 3].to(DEV),
                                    constant_in_time=False, y=y, e=e.to(DEV))
    assert xr.shape == (2, 10, 256, 3)
    assert list(model.get_nfe().astype(int)) == list(gold['eval_nfe'].astype(int))
    assert _rel(xr, gold['eval_x_rec']) < 1e-4
    # strided (non-contiguous) observed input gives the same answer
    _, _, xr2, _ = model.reconstruct(x.to(DEV)[:, [0, 5, 9]], num_points=256, timestamps=nocs[0, :, 0, 3].to(DEV),
                                     y=y, e=e.to(DEV))
    assert torch.equal(xr, xr2)


def test_real_demo_sequences(setup):
    """Two real sequences of /root/reference/data/demo (decoded by the reference's loader, stored in the fixture).

    Real depth data is quantised (millimetre steps, many exactly tied distances), which makes this the hardest geometry
    case: FPS and ball-query indices of all five levels must still equal the oracle's bit for bit.  It is also the one
    input on which the reference's OWN fp32 result is ill-conditioned: per-ball GroupNorm divides rounding noise of
    camera-space features (|z^2| ~ 6) by in-ball spreads of ~1e-2, so the unmodified reference differs from its float64
    evaluation by 9e-4 (z0), 2e-4 (reconstruction), and a one-ulp change of the input moves the reference's z0 by 1.7e-3
    (measured).  A 1e-4 bar against the fp32 fixture is therefore not meaningful here; the yardstick
    is the float64 evaluation (`demo_*_f64`): the CUDA path must be as close to it as the reference is, within a factor
    of 4, and stay within 5e-3 of the fp32 fixture.  Measured on B200: 0.13x for z0, 0.27x for the reconstruction - the
    tensor-core set-abstraction kernel normalises differences to each ball's first row (sa_mma.cu), which removes most of
    that cancellation, so the CUDA path is closer to the float64 evaluation than the reference is."""
    gold, _, model, sd = setup
    x = torch.from_numpy(gold['demo_x'])
    y, e = _seeded_y_e(17, (10, 256, 3))
    oracle = CasprOracle(sd)
    oracle.encode(x)
    model.encoder.trace = {}
    z0, tn = model.encode(x.to(DEV))
    trace, model.encoder.trace = model.encoder.trace, None
    for lvl in range(5):
        assert torch.equal(trace['fps_idx'][lvl].cpu(), oracle.trace['fps_idx_%d' % lvl]), 'FPS level %d' % lvl
        for s_ in range(2):
            assert torch.equal(trace['ball_idx'][lvl][s_].cpu(), oracle.trace['ball_idx_%d_%d' % (lvl, s_)])
    _, _, xr, _ = model.reconstruct(x.to(DEV), num_points=256, y=y, e=e.to(DEV))
    assert list(model.get_nfe().astype(int)) == list(gold['demo_nfe'].astype(int))
    for name, got, ref32, ref64 in (('z0', z0, gold['demo_z0'], gold['demo_z0_f64']),
                                    ('x_rec', xr, gold['demo_x_rec'], gold['demo_x_rec_f64']),
                                    ('tnocs', tn[:, :, ::4], gold['demo_tnocs'], gold['demo_tnocs_f64'])):
        err_gpu, err_ref = _rel(got, ref64), _rel(ref32, ref64)
        assert err_ref > 1e-4, name              # the premise: the reference itself is not reproducible to 1e-4 here
        assert err_gpu < 4.0 * err_ref, (name, err_gpu, err_ref)
        assert _rel(got, ref32) < 5e-3, name
    cd = chamfer_distance(xr.cpu().view(10, 256, 3), torch.from_numpy(gold['demo_x_rec_f64']).view(10, 256, 3))
    assert float(cd.max()) < 1e-6


def test_sample_contours_branch(setup):
    """decode(sample_contours=[0.5, 1.0]) through the interpolated call of utils/viz_utils.py:142-148: the numpy RNG
    stream (utils/transform_utils.py:80-85) gives the fixture's base points exactly; constant_in_time shares them."""
    gold, _, model, _ = setup
    x, _ = synthetic_sequences(1, 3, 1024, seed=1)
    np.random.seed(3)
    yy, logp, xr, _ = model.reconstruct(x.to(DEV), num_points=128, timestamps=torch.linspace(0, 1, 4).to(DEV),
                                        constant_in_time=True, sample_contours=[0.5, 1.0],
                                        e=torch.from_numpy(gold['cont_e']).to(DEV))
    assert _rel(yy, gold['cont_y']) < 1e-6          # float64 numpy points cast to float32 on either side
    assert torch.equal(yy[:, 0], yy[:, 3])
    r = yy[0, 0].norm(dim=1)
    assert torch.allclose(r[:64], torch.full_like(r[:64], 0.5), atol=1e-5)
    assert torch.allclose(r[64:], torch.full_like(r[64:], 1.0), atol=1e-5)
    assert list(model.get_nfe().astype(int)) == list(gold['cont_nfe'].astype(int))
    assert _rel(logp, gold['cont_logp_y']) < 1e-5
    assert _rel(xr, gold['cont_x_rec']) < 1e-4


def test_truncate_std_branch(setup):
    """reconstruct(truncate_std=1.5).  (1) With the fixture's base points injected the decoded cloud matches the
    reference run; (2) the un-injected branch executes models/utils.py:15-22 on the device (as the reference does on
    a GPU: `tensor.new_empty(...).normal_()` is a CUDA draw, so its values are not comparable with a CPU run) and the
    samples obey the truncation rule; (3) `truncated_normal` itself reproduces the CPU fixture draw for draw."""
    from caspr_b200.models.utils import sample_gaussian
    gold, _, model, sd = setup
    x, _ = synthetic_sequences(1, 3, 1024, seed=1)
    _, _, xr, _ = model.reconstruct(x.to(DEV), num_points=128, y=torch.from_numpy(gold['trunc_y']).view(3, 128, 3),
                                    e=torch.from_numpy(gold['trunc_e']).to(DEV))
    assert list(model.get_nfe().astype(int)) == list(gold['trunc_nfe'].astype(int))
    assert _rel(xr, gold['trunc_x_rec']) < 1e-4
    torch.manual_seed(9)
    assert np.array_equal(sample_gaussian((3, 128, 3), 1.5).numpy(), gold['trunc_y'].reshape(3, 128, 3))
    torch.manual_seed(21)
    yy, logp, xr2, _ = model.reconstruct(x.to(DEV), num_points=2048, truncate_std=1.5)
    # four candidates per value: P(all four outside 1.5 sigma) = 0.1336^4 = 3.2e-4 -> a handful of 18 432 values
    frac_out = float((yy.abs() >= 1.5).float().mean())
    assert frac_out < 2e-3
    assert float(yy.std()) < 0.8                     # truncated at 1.5 sigma: std 0.74
    assert torch.isfinite(xr2).all()
    oracle = CasprOracle(sd)
    e = torch.randn(3, 2048, 3)
    _, _, xr3, _ = model.reconstruct(x.to(DEV), num_points=2048, y=yy.cpu().view(3, 2048, 3), e=e.to(DEV))
    _, _, xr_ref, _ = oracle.reconstruct(x, num_points=2048, y=yy.cpu().view(3, 2048, 3), e=e)
    assert list(model.get_nfe().astype(int)) == oracle.get_nfe()
    assert _rel(xr3, xr_ref) < 1e-4


def test_full_config2_shape_matches_oracle(setup):
    """The benchmarked shape: B=8, T=10, N=1024, P=2048 (BASELINE configs[1]) against the oracle on the host cores
    (about a minute).  FPS indices of every level bit-exact for all 80 clouds, identical NFE, coordinates 1e-4."""
    _, _, model, sd = setup
    torch.set_num_threads(os.cpu_count() or 8)
    B, T, N, P = 8, 10, 1024, 2048
    x, _ = synthetic_sequences(B, T, N, seed=0)          # bench.py's input
    g = torch.Generator().manual_seed(1)
    y = torch.randn(B * T, P, 3, generator=g)
    e = torch.randn(B * T, P, 3, generator=g)
    model.encoder.trace = {}
    yy, logp, xr, tn = model.reconstruct(x.to(DEV), num_points=P, y=y, e=e.to(DEV))
    trace, model.encoder.trace = model.encoder.trace, None
    nfe = list(model.get_nfe().astype(int))
    oracle = CasprOracle(sd)
    _, logp_ref, xr_ref, tn_ref = oracle.reconstruct(x, num_points=P, y=y, e=e)
    for lvl in range(5):
        assert torch.equal(trace['fps_idx'][lvl].cpu(), oracle.trace['fps_idx_%d' % lvl]), 'FPS level %d' % lvl
    assert nfe == oracle.get_nfe()
    print('full config-2 shape vs oracle: T-NOCS %.2e, logp_y %.2e, x_rec %.2e (relative, bars 1e-4 / 1e-5 / 1e-4)'
          % (_rel(tn, tn_ref), _rel(logp, logp_ref), _rel(xr, xr_ref)))
    assert _rel(tn, tn_ref) < 1e-4
    assert _rel(logp, logp_ref) < 1e-5
    assert _rel(xr, xr_ref) < 1e-4
    cd = chamfer_distance(xr.cpu().view(B * T, P, 3), xr_ref.view(B * T, P, 3))
    assert float(cd.max()) < 5e-8
