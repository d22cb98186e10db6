"""Drop-in mirror of the reference's ``caspr/models`` package for the reconstruction hot path.

Same class names, constructor arguments, ``forward`` / ``encode`` / ``reconstruct`` / ``decode`` /
``get_nfe`` signatures and 238-key ``state_dict`` layout as /root/reference/caspr/models (so the
reference's train.py / test.py / viz.py and their checkpoints load unchanged); every heavy
operation runs in libcaspr_b200.so (hand-written sm_100a CUDA) through ``caspr_b200.ops``.
"""
from .caspr import CaSPR  # noqa: F401
