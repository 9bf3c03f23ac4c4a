"""nnuzoo_b200 -- B200-native (sm_100a) selective scan + SS2D CrossScan/CrossMerge for nnUZoo.

Public surface (mirrors the reference's operator / module interface for this one path):
  selective_scan_fn, SelectiveScanFn   <- nnunetv2/nets/seg_mamba/selective_scan_interface.py:14-83
  cross_scan, cross_merge              <- m2net.py:175-177, 202-206, 218; ssnd2net.py:249-255, 285-299
  SS2D, SSND                           <- m2net.py:39-225; ssnd2net.py:73-318
  Mamba                                <- seg_mamba/mamba_simple.py:37-357 (1-D nets; uni / bi / tri-directional)
  causal_conv1d_fn                     <- mamba_simple.py:316-324 (the block's depthwise causal conv + SiLU)

Everything runs on hand-written CUDA behind the C ABI of include/nnuzoo_b200.h; there is no CPU
or PyTorch fallback (the CPU oracle under oracle/ is test infrastructure only).
"""
from .selective_scan_interface import SelectiveScanFn, selective_scan_fn  # noqa: F401
from .cross_scan import cross_merge, cross_scan  # noqa: F401
from .ss2d import SS2D, SSND  # noqa: F401
from .causal_conv1d import causal_conv1d_fn  # noqa: F401
from .mamba import Mamba  # noqa: F401

__all__ = ["selective_scan_fn", "SelectiveScanFn", "cross_scan", "cross_merge", "SS2D", "SSND", "Mamba", "causal_conv1d_fn"]
