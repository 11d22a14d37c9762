"""elvis_b200 -- B200-native (sm_100a) implementation of the ELVIS / PRESLEY pre- and
post-processing hot path: removability scoring, per-row top-k removal masks, shrink,
stretch and the v2 per-block degradations.

  elvis_b200.elvis / .utils / .presley   drop-in mirrors of the reference's functions
  elvis_b200.ops                         device-tensor operators (one C-ABI call each)
  elvis_b200.pipeline                    batched planar-YUV pipelines, frame sharding
  include/elvis_b200.h                   the C ABI of libelvis_b200.so

There is no CPU fallback: importing the operator modules without the built library raises
(python -m elvis_b200.build)."""
__version__ = "0.1.0"
