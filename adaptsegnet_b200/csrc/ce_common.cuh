// Shared by the cross-entropy kernels (pointwise.cu: full-res logits; lazy_up.cu: interpolated on the fly).
#pragma once
#include "common.cuh"

namespace asn {

// device-side accumulator behind the `stats` argument of asn_softmax_ce_* / asn_upsample_ce_fwd_bwd
struct CeStats {
  double loss_sum;
  double weight_sum;
  long long n_valid;
  long long n_bad;
};

// label classification shared by fwd and bwd:  1 = contributes, 0 = ignored, -1 = out of bounds
__device__ __forceinline__ int classify_label(long long y, int C, int ignore, int mask_negative) {
  if (y == (long long)ignore) return 0;
  if (mask_negative && y < 0) return 0;
  if (y < 0 || y >= C) return -1;
  return 1;
}

}  // namespace asn
