"""soccernerfs_b200: the K-Planes train/render hot path of iSach/SoccerNeRFs on B200 (sm_100a) CUDA kernels.

Python host code mirrors the reference's nerfstudio plugin surface (same class names, arguments, outputs);
all arithmetic on the path runs in hand-written kernels reached through the C-ABI of ``libkplanes_b200.so``
(``include/kplanes_b200.h``).  There is no CPU, Triton or PyTorch-eager fallback.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
