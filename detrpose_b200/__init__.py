"""B200-native (sm_100a) multi-scale deformable attention for DETRPose.

One hot path, built from scratch: the sampling core of the pose decoder's
deformable cross-attention (bilinear gather over the feature pyramid, attention
weighted reduction, forward + backward), as hand-written CUDA behind a C ABI
(``include/msda_b200.h``), with the host side mirroring the reference's module
and function interface.  See DESIGN.md and INTEGRATION.md.
"""
from .functional import (ms_deform_attn_core, sample_indices, level_start_index, locations_and_weights,
                         pack_value, clear_repack_cache, set_default_coord_mode, get_default_coord_mode)
from .ms_deform_attn import MSDeformAttn
from .gate import Gate, gate_epilogue
from .lqe import LQE, lqe_statistics
from . import patch, synthetic, shard

__all__ = ["MSDeformAttn", "Gate", "gate_epilogue", "LQE", "lqe_statistics", "ms_deform_attn_core", "sample_indices", "level_start_index",
           "locations_and_weights", "pack_value", "clear_repack_cache", "set_default_coord_mode",
           "get_default_coord_mode", "patch", "synthetic", "shard"]
__version__ = "0.1.0"
