"""Shim for torchdiffeq==0.0.1 (absent here): the oracle restatement of its dopri5."""
from oracle.odeint001 import odeint, odeint_adjoint  # noqa: F401
__version__ = '0.0.1-oracle-restatement'
