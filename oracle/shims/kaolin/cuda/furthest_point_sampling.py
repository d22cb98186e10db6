"""Shim: reference pointnet2.py:10 imports this module only for its side effect."""
