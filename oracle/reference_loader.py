"""Import the UNMODIFIED reference model code over the oracle shims.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``/root/reference/caspr/models`` cannot be imported as-is in this image:
``kaolin`` (pointnet2.py:7-10), ``torchdiffeq`` (latent_ode_model.py:9,
cnf.py:11-12) and ``open3d`` (cnf.py:14, transform_utils.py:2) are not
installed, and ``get_point_cnf`` hard-calls ``.cuda()`` (flow.py:81).  This
loader puts ``oracle/shims`` (packages of those names backed by the oracle
restatements) and the reference's ``caspr/`` directory on ``sys.path`` and makes
``nn.Module.cuda`` a no-op when no GPU is present.  The reference sources are
read where they lie; nothing is copied.

``/root/reference`` exists only in the development container — callers must
check :func:`reference_available` first (the GPU box does not have it).
"""
import contextlib
import io
import os
import sys

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get('CASPR_REFERENCE_ROOT', '/root/reference')
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shims')
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'caspr', 'models', 'caspr.py'))


def load_reference_models():
    """Returns the reference's ``models.caspr`` module (class ``CaSPR`` etc.)."""
    if not reference_available():
        raise RuntimeError('reference sources not present at %s' % REFERENCE_ROOT)
    for p in (_REPO, os.path.join(REFERENCE_ROOT, 'caspr'), _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    if not torch.cuda.is_available():
        nn.Module.cuda = lambda self, device=None: self      # flow.py:81
    import importlib
    return importlib.import_module('models.caspr')


def build_reference_caspr(quiet=True, **kwargs):
    """Instantiate the reference ``CaSPR`` (prints from its ctor are swallowed)."""
    mod = load_reference_models()
    ctx = contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()
    with ctx:
        model = mod.CaSPR(**kwargs)
    return model
