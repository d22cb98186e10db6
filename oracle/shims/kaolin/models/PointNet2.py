"""Shim for `kaolin.models.PointNet2` — the six names bound at reference pointnet2.py:7."""
from oracle.pointnet2_ops import (separate_xyz_and_features, PointNet2GroupingLayer,  # noqa: F401
                                  furthest_point_sampling, fps_gather_by_index,
                                  three_nn, three_interpolate)
