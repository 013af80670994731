"""dvp_mvs_b200 — B200-native PatchMatch-MVS engine behind DVP-MVS's APD interface.

Only what the hot path needs lives here:
  csrc/        hand-written sm_100a CUDA kernels + the flat C ABI (include/dvp_mvs.h) -> libdvp_mvs.so
  _lib.py      ctypes binding of that ABI (`Engine`)
  (the C++ mirror of the reference's `APD` class is include/dvp_apd_adapter.hpp)
  synth.py     deterministic synthetic scenes (the reference ships no data)
  farm.py      per-view sharding across the GPUs of one box (NCCL-free)
"""
from ._lib import (Engine, Scene, Farm, Fusion, FusionView, make_fusion_view, edge_segment, label_segment, resize_linear_f32, Params, Inputs, DvpError, default_params, FIRST_INIT, REFINE_INIT, REFINE_ITER,
                   WEAK, STRONG, UNKNOWN, STAGES, PRODUCT_LIB)
from . import synth
from . import formats

__all__ = ["Engine", "Scene", "Farm", "Fusion", "FusionView", "make_fusion_view", "edge_segment", "label_segment", "resize_linear_f32", "Params", "Inputs", "DvpError", "default_params", "FIRST_INIT", "REFINE_INIT", "REFINE_ITER",
           "WEAK", "STRONG", "UNKNOWN", "STAGES", "PRODUCT_LIB", "synth", "formats"]
