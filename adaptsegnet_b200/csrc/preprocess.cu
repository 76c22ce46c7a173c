// SURVEY.md section 8f row 4 (input pipeline), the part with exact semantics: what dataset/gta5_dataset.py:58-71 does to a
// decoded, resized sample before it becomes a tensor --
//     image = image[:, :, ::-1]; image -= mean; image = image.transpose((2, 0, 1))        uint8 RGB HWC -> fp32 BGR CHW
//     label_copy[label == k] = v for the 19 (id, train id) pairs, 255 elsewhere           uint8 ids   -> train ids
// -- as two HBM-bound kernels working on the raw 8-bit buffers (3 B/px in, 12 B/px out; 1 B/px in, 8 B/px out), so that a
// loader hands 2.8 MB + 0.9 MB per 720x1280 frame to the device instead of 11 MB + 7.4 MB of host-side float arrays.
// The PIL resize in front of it (BICUBIC / NEAREST, :51-52) is not restated: Pillow's fixed-point resampling has no
// reference vectors in the repository to pin it against (dataset files are not available).
#include "common.cuh"

namespace asn {

// thread = 4 consecutive pixels of one image: 12 input bytes (three 32-bit loads), three 16-byte plane stores
__global__ void __launch_bounds__(256)
image_u8_to_bgr_f32_kernel(const uint8_t* __restrict__ rgb, float* __restrict__ out, int64_t n_img, int64_t HW,
                           float mean_b, float mean_g, float mean_r) {
  const int64_t q_per_img = HW / 4;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_img * q_per_img; i += (int64_t)gridDim.x * 256) {
    const int64_t img = i / q_per_img, q = i - img * q_per_img;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(rgb + (img * HW + q * 4) * 3);
    const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);   // r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
    const float r[4] = {(float)(w0 & 0xff), (float)(w0 >> 24), (float)((w1 >> 16) & 0xff), (float)((w2 >> 8) & 0xff)};
    const float g[4] = {(float)((w0 >> 8) & 0xff), (float)(w1 & 0xff), (float)(w1 >> 24), (float)((w2 >> 16) & 0xff)};
    const float b[4] = {(float)((w0 >> 16) & 0xff), (float)((w1 >> 8) & 0xff), (float)(w2 & 0xff), (float)(w2 >> 24)};
    float* o = out + img * 3 * HW + q * 4;   // planes: B, G, R
    st_stream(reinterpret_cast<float4*>(o), make_float4(b[0] - mean_b, b[1] - mean_b, b[2] - mean_b, b[3] - mean_b));
    st_stream(reinterpret_cast<float4*>(o + HW), make_float4(g[0] - mean_g, g[1] - mean_g, g[2] - mean_g, g[3] - mean_g));
    st_stream(reinterpret_cast<float4*>(o + 2 * HW), make_float4(r[0] - mean_r, r[1] - mean_r, r[2] - mean_r, r[3] - mean_r));
  }
}

// thread = 16 consecutive labels: one 16-byte load, the 256-entry table in shared memory, eight 16-byte stores
__global__ void __launch_bounds__(256)
label_u8_to_trainid_i64_kernel(const uint8_t* __restrict__ ids, const uint8_t* __restrict__ lut_g, long long* __restrict__ out,
                               int64_t n16) {
  __shared__ uint8_t lut[256];
  lut[threadIdx.x] = lut_g[threadIdx.x];
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += (int64_t)gridDim.x * 256) {
    const uint4 v = ld_stream(reinterpret_cast<const uint4*>(ids) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    longlong2* o = reinterpret_cast<longlong2*>(out + i * 16);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t a = (w[k >> 1] >> ((k & 1) * 16)) & 0xff, b2 = (w[k >> 1] >> ((k & 1) * 16 + 8)) & 0xff;
      o[k] = make_longlong2((long long)lut[a], (long long)lut[b2]);
    }
  }
}

}  // namespace asn

using namespace asn;

extern "C" int asn_image_u8_to_bgr_f32(const uint8_t* rgb_hwc, float* out_chw, int N, int H, int W, float mean_b,
                                       float mean_g, float mean_r, void* stream) {
  ASN_CHECK_ARG(rgb_hwc && out_chw && N > 0 && H > 0 && W > 0, "asn_image_u8_to_bgr_f32: bad argument");
  ASN_CHECK_ARG(((int64_t)H * W) % 4 == 0 && (reinterpret_cast<uintptr_t>(rgb_hwc) & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(out_chw) & 15) == 0,
                "asn_image_u8_to_bgr_f32: H*W must be a multiple of 4, buffers 4- / 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t HW = (int64_t)H * W;
  prof::Scope ps("image_u8_to_bgr_f32", 0, 15.0 * N * HW, st);
  image_u8_to_bgr_f32_kernel<<<full_grid(N * HW / 4, 256), 256, 0, st>>>(rgb_hwc, out_chw, N, HW, mean_b, mean_g, mean_r);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

extern "C" int asn_label_u8_to_trainid_i64(const uint8_t* ids, const uint8_t* lut256, int64_t* out, int64_t n_px,
                                           void* stream) {
  ASN_CHECK_ARG(ids && lut256 && out && n_px > 0, "asn_label_u8_to_trainid_i64: bad argument");
  ASN_CHECK_ARG(n_px % 16 == 0 && (reinterpret_cast<uintptr_t>(ids) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "asn_label_u8_to_trainid_i64: n_px must be a multiple of 16, buffers 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  prof::Scope ps("label_u8_to_trainid_i64", 0, 9.0 * n_px, st);
  label_u8_to_trainid_i64_kernel<<<full_grid(n_px / 16, 256), 256, 0, st>>>(ids, lut256, reinterpret_cast<long long*>(out),
                                                                             n_px / 16);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}
