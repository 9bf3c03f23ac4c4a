"""nnuzoo_b200 -- B200-native (sm_100a) selective scan + SS2D CrossScan/CrossMerge for nnUZoo.

Public surface (mirrors the reference's operator / module interface for this one path):
  selective_scan_fn, SelectiveScanFn   <- nnunetv2/nets/seg_mamba/selective_scan_interface.py:14-83
  cross_scan, cross_merge              <- m2net.py:175-177, 202-206, 218; ssnd2net.py:249-255, 285-299
  SS2D, SSND                           <- m2net.py:39-225; ssnd2net.py:73-318
  Mamba                                <- seg_mamba/mamba_simple.py:37-357 (1-D nets; uni / bi / tri-directional)
  causal_conv1d_fn                     <- mamba_simple.py:316-324 (the block's depthwise causal conv + SiLU)
  mamba_inner_fn(_no_out_proj)         <- selective_scan_interface.py:159-434, :609-628 (conv -> projections -> scan, one node)
  MambaLayer, ResMambaBlock, nd_mamba_order, Block, create_block, MambaNDCore (nnuzoo_b200.mamba_nd)
                                       <- lm2net.py:64-176, mamba_nd2net.py:565-722, :725-1001 (callers of the 1-D path)
  M2Net, get_m2net (nnuzoo_b200.m2net) <- m2net.py:805-971, :1187-1208 (SS2D^2-Net, the config-2 network)
  Trainer (nnuzoo_b200.train)          <- nnUNetTrainer.py:1112-1144 + nnUNetTrainerM2Net.py (DDP training step)
  SlidingWindowPredictor (.predict)    <- inference/predict_from_raw_data.py:515-690 (rank-sharded tiles)
SS2D glue kernels behind the same C ABI: grouped_proj / proj_wgrad, LayerNorm, dwconv3x3_silu, ss2d_core (fused
scan + CrossMerge + out_norm + gate).

Everything runs on hand-written CUDA behind the C ABI of include/nnuzoo_b200.h; there is no CPU
or PyTorch fallback (the CPU oracle under oracle/ is test infrastructure only).
"""
from .selective_scan_interface import SelectiveScanFn, selective_scan_fn  # noqa: F401
from .cross_scan import cross_merge, cross_scan  # noqa: F401
from .ss2d import SS2D, SSND  # noqa: F401
from .causal_conv1d import causal_conv1d_fn  # noqa: F401
from .mamba import Mamba  # noqa: F401
from .mamba_inner import mamba_inner_fn, mamba_inner_fn_no_out_proj  # noqa: F401
from .mamba_nd import Block, MambaLayer, MambaNDCore, ResMambaBlock, create_block, nd_mamba_order  # noqa: F401
from .norm import LayerNorm, layer_norm  # noqa: F401
from .proj import grouped_proj, proj_wgrad  # noqa: F401
from .dwconv import dwconv3x3_silu  # noqa: F401
from .fused import ss2d_core  # noqa: F401
from .m2net import M2Net, get_m2net  # noqa: F401

__all__ = ["selective_scan_fn", "SelectiveScanFn", "cross_scan", "cross_merge", "SS2D", "SSND", "Mamba", "causal_conv1d_fn",
           "mamba_inner_fn", "mamba_inner_fn_no_out_proj", "MambaLayer", "ResMambaBlock", "nd_mamba_order", "Block",
           "create_block", "MambaNDCore",
           "LayerNorm", "layer_norm", "grouped_proj", "proj_wgrad", "dwconv3x3_silu", "ss2d_core", "M2Net", "get_m2net"]
