// "Lazy upsample" (Tier-B): consumers of the bilinearly upsampled logits that read the LOW-RES
// logits and interpolate on the fly, so that the full-resolution (N,C,H,W) fp32 tensors of the reference
// (interp(...) model/deeplab_multi.py:188-189, F.softmax(...) train...:617-618) never exist in HBM.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>

namespace asn {
namespace lazy {

// true when the strip kernels cover this case (C == 19, upsampling in both directions)
bool supported(int C, int h, int w, int H, int W);
// bytes of the per-CTA partial sums the transposed interpolation needs (CE and the discriminator-input backward)
size_t partial_bytes(int N, int C, int h, int w, int H, int W);

// discriminator input: A0[n][Y][X+1][0..31] (bf16, zero padded) = softmax_c(upsample(z_low))[n, :, Y, X]
int pack_input(const float* z_low, __nv_bfloat16* a0, int N, int C, int h, int w, int H, int W, int W0p,
               cudaStream_t st);
// its backward: dz_low = Up^T( p * (dA0 - sum_c p * dA0) ),  p recomputed from z_low
int unpack_dx(const __nv_bfloat16* da0, const float* z_low, float* dz_low, int N, int C, int h, int w, int H, int W,
              int W0p, void* partial, size_t partial_size, cudaStream_t st);

}  // namespace lazy
}  // namespace asn
