"""Host-side mirror of the reference interface: state_dict layout, loaders, loud failure without CUDA."""
import pytest
import torch

from caspr_b200.synth import load_manifest, synthetic_state_dict, synthetic_sequences


def test_state_dict_matches_reference_manifest(lib_built):
    from caspr_b200.models import CaSPR
    model = CaSPR()
    sd = model.state_dict()
    man = load_manifest()
    assert list(sd.keys()) == list(man.keys())          # 238 keys, same order as the reference
    assert len(sd) == 238
    for k, shape in man.items():
        assert list(sd[k].shape) == shape, k
    assert sum(p.numel() for p in model.parameters()) == 16262189


def test_load_reference_style_checkpoint(lib_built):
    from caspr_b200.models import CaSPR
    sd = synthetic_state_dict(0)
    model = CaSPR()
    # reference torch_utils.load_weights strips a DataParallel 'module.' prefix and uses strict=False
    wrapped = {'module.' + k: v for k, v in sd.items()}
    stripped = {k[len('module.'):]: v for k, v in wrapped.items()}
    missing, unexpected = model.load_state_dict(stripped, strict=False)
    assert not missing and not unexpected
    enc = {k[len('encoder.'):]: v for k, v in sd.items() if k.startswith('encoder.')}
    model.encoder.load_state_dict(enc)                  # torch_utils.load_encoder_weights_from_full
    a = model.latent_ode.ode_func.dynamics_net[0].weight
    b = model.latent_ode.solver.ode_func.dynamics_net[0].weight
    assert a.data_ptr() == b.data_ptr()                 # aliased module registered twice


def test_ctor_variants(lib_built):
    from caspr_b200.models import CaSPR
    m = CaSPR(pretrain_tnocs=True)
    assert not hasattr(m, 'point_cnf')
    m = CaSPR(regress_tnocs=False)
    assert not hasattr(m.encoder, 'conv3')
    m = CaSPR(augment_quad=False, augment_pairs=False)
    assert m.encoder.local_extract.set_abstractions[0].pointnet_modules[0].conv_layers[0].weight.shape[1] == 3


def test_no_cpu_fallback(lib_built):
    from caspr_b200.models import CaSPR
    model = CaSPR().eval()
    x, _ = synthetic_sequences(1, 2, 64, seed=0)
    with pytest.raises((RuntimeError, TypeError)):
        model.encode(x)                                  # CPU tensor: must fail loudly, never fall back


def test_synthetic_generators_are_deterministic():
    a = synthetic_state_dict(0)
    b = synthetic_state_dict(0)
    assert all(torch.equal(a[k], b[k]) for k in a)
    x1, n1 = synthetic_sequences(2, 3, 128, seed=4)
    x2, n2 = synthetic_sequences(2, 3, 128, seed=4)
    assert torch.equal(x1, x2) and torch.equal(n1, n2)
    assert x1.shape == (2, 3, 128, 4) and float(x1[..., 2].min()) > 0.5


def test_training_gradient_layouts_match_parameter_order():
    """The flat gradient buffers the adjoint entry points fill (caspr_cnf_adjoint / caspr_latent_ode_adjoint) are cut
    into per-parameter views in ``parameters()`` order: sizes and order must match the modules (no GPU needed: the
    *_param_count queries are host functions)."""
    from caspr_b200._lib import lib
    from caspr_b200.models import CaSPR
    model = CaSPR()
    cnf = model.point_cnf.chain[1]
    names = [n for n, _ in cnf.odefunc.named_parameters()]
    per_layer = ['_layer.weight', '_layer.bias', '_hyper_bias.weight', '_hyper_gate.weight', '_hyper_gate.bias']
    assert names == ['diffeq.layers.%d.%s' % (l, s) for l in range(4) for s in per_layer]
    assert lib.caspr_cnf_param_count(512, 1600) == sum(p.numel() for p in cnf.odefunc.parameters())
    net = model.latent_ode.ode_func.dynamics_net
    lat = [p for l in (net[0], net[2], net[4], net[6]) for p in (l.weight, l.bias)]
    assert [tuple(p.shape) for p in lat] == [(512, 64), (512,), (512, 512), (512,), (512, 512), (512,), (64, 512), (64,)]
    assert lib.caspr_latent_ode_param_count(64, 512) == sum(p.numel() for p in lat)
    # shared module registered twice (latent_ode.ode_func / latent_ode.solver.ode_func): one set of Parameters
    assert len(list(model.latent_ode.parameters())) == 8


def test_derived_weight_cache_follows_parameter_versions(lib_built):
    """Matrices derived from a parameter (column subsets of the head's / set-abstraction first layers, ops.derived_weight)
    are rebuilt when the parameter is modified in place (version counter) and after `invalidate_weight_cache()` (edits
    through `.data`, which bypass the counter) - host logic only, no kernel runs."""
    import torch
    from caspr_b200 import ops
    p = torch.nn.Parameter(torch.arange(12.0).view(3, 4))
    calls = []

    def tail(w):
        calls.append(1)
        return w[:, 2:]

    a = ops.derived_weight(p, 'tail', tail)
    assert a.is_contiguous() and torch.equal(a, p.detach()[:, 2:]) and len(calls) == 1
    assert ops.derived_weight(p, 'tail', tail) is a and len(calls) == 1            # cached
    with torch.no_grad():
        p.mul_(2.0)                                                                # bumps the version
    b = ops.derived_weight(p, 'tail', tail)
    assert len(calls) == 2 and torch.equal(b, p.detach()[:, 2:])
    p.data.add_(1.0)                                                               # does NOT bump the version
    assert ops.derived_weight(p, 'tail', tail) is b
    ops.invalidate_weight_cache()
    c = ops.derived_weight(p, 'tail', tail)
    assert len(calls) == 3 and torch.equal(c, p.detach()[:, 2:])
    # a different tag on the same parameter is a different entry
    d = ops.derived_weight(p, 'head', lambda w: w[:, :2])
    assert torch.equal(d, p.detach()[:, :2]) and ops.derived_weight(p, 'tail', tail) is c
