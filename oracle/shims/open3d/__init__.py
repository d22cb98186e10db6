"""Shim: open3d is imported by reference cnf.py:14 and transform_utils.py:2 but never
called on the hot path."""
