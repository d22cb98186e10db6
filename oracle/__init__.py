"""CPU oracle for the CaSPR reconstruction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker (or the timed CPU baseline), never as the thing shipped.  The product
path (``caspr_b200``) must fail loudly when its CUDA library is missing and
must never route through this package.

PARITY STATUS: **parity unpinned**.  The reference repository
(davrempe/caspr @ b8360da) ships no tests, golden vectors or known-answer
fixtures for this path, and the native arithmetic
it executes lives in two un-vendored third-party packages that are absent from
``/root/reference`` and cannot be installed here (no network):

* Kaolin (NVIDIAGameWorks/kaolin, unpinned git master of early/mid 2020,
  ``kaolin.models.PointNet2`` ops derived from erikwijmans/Pointnet2_PyTorch):
  restated in ``oracle/pointnet2_ops.py``.
* torchdiffeq == 0.0.1 (pinned by the reference README:22): restated in
  ``oracle/odeint001.py``.

What IS pinned: the reference's own model code.  ``oracle/reference_loader.py``
imports ``/root/reference/caspr/models`` UNCHANGED over shim packages backed by
the two restatements above, and ``tests/golden/make_golden.py`` froze its
outputs as fixtures; the independent restatement in ``oracle/caspr_oracle.py``
is checked against those fixtures (and against the live reference modules when
``/root/reference`` is present).
"""
