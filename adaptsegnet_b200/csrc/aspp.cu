// K1 / K1b: ASPP classifier head (model/deeplab_multi.py:106-121) on tcgen05.
//
// The four dilated 3x3 convolutions with N = 19 outputs are hostile to a direct implicit GEMM
// (N padded 19 -> 32, and the pixel x 36*Cin im2col operand would stream ~2 GB through L2 per
// call).  Instead the head is computed as ONE dense GEMM over the un-shifted activations
// followed by a tiny shifted gather:
//     Z[p, t*19+c] = sum_ci X[p, ci] * W_t[c, ci]          (M = pixels, N = 36*19 = 684 -> 688, K = Cin)
//     y[c, p]      = bias_sum[c] + sum_t Z[p + shift_t, t*19+c]     (zero outside the image)
// The branch sum of deeplab_multi.py:117-121 is the sum over t.  Backward:
//     dYcol[q, t*19+c] = dy[c, q - shift_t]                 (tiny: dy has 19 channels)
//     dX^T[ci, q]  = sum_j WpT[ci, j] * dYcol[q, j]         (M = Cin, N = pixels, K = 688; lands in NCHW)
//     dWp[ci, j]   = sum_q X[ci, q] * dYcolT[j, q]          (M = Cin, N = 688, K = pixels; split-K)
// All three GEMMs have dense, tensor-core-friendly shapes; X is read once per GEMM as bf16.
#include "umma_host.cuh"

namespace asn {


// ---- layout kernels ------------------------------------------------------------------------

// x [N][C][P] fp32 -> out [N*P][C] bf16 (64 x 64 tiles through shared memory; 256-byte reads,
// 128-byte writes per row)
__global__ void __launch_bounds__(256)
nchw_to_nhwc_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int C, int P) {
  __shared__ float tile[64][65];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  const float* src = x + (int64_t)n * C * P;
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    int cl = i / 64, pl = i % 64;
    int c = c0 + cl, p = p0 + pl;
    tile[cl][pl] = (c < C && p < P) ? __ldg(src + (int64_t)c * P + p) : 0.f;
  }
  __syncthreads();
  __nv_bfloat16* dst = out + (int64_t)n * P * C;
  for (int i = threadIdx.x; i < 64 * 32; i += 256) {
    int pl = i / 32, cp = (i % 32) * 2;
    int p = p0 + pl, c = c0 + cp;
    if (p < P && c + 1 < C) {
      __nv_bfloat162 v = __floats2bfloat162_rn(tile[cp][pl], tile[cp + 1][pl]);
      *reinterpret_cast<__nv_bfloat162*>(dst + (int64_t)p * C + c) = v;
    } else if (p < P && c < C) {
      dst[(int64_t)p * C + c] = __float2bfloat16(tile[cp][pl]);
    }
  }
}

// x [N][C][P] fp32 -> out [C][ld] bf16, column n*P + p  (K-contiguous operand for the wgrad GEMM).
// blockIdx.y = (n, c) row; VEC = 4 pixels per thread (16-byte loads, 8-byte stores).
template <int VEC>
__global__ void __launch_bounds__(256)
nchw_to_ckp_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int C, int P,
                        int64_t ld) {
  const int n_rows = N * C;
  const int pv = (P + VEC - 1) / VEC;
  for (int row = blockIdx.y; row < n_rows; row += gridDim.y) {
    const int n = row / C, c = row - n * C;
    const float* src = x + (int64_t)row * P;
    __nv_bfloat16* dst = out + (int64_t)c * ld + (int64_t)n * P;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < pv; i += gridDim.x * 256) {
      if (VEC == 4) {
        const float4 v = ld_stream(reinterpret_cast<const float4*>(src) + i);
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        *reinterpret_cast<uint2*>(dst + (int64_t)i * 4) = pk;
      } else {
        dst[i] = __float2bfloat16(__ldg(src + i));
      }
    }
  }
}

// channels_last input: x [N*P][C] fp32 -> bf16, same layout (the trunk already produced NHWC)
__global__ void __launch_bounds__(256)
nhwc_f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 v = ld_stream(reinterpret_cast<const float4*>(x) + i);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    reinterpret_cast<uint2*>(out)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

// channels_last input: x [NP][C] fp32 -> out [C][ld] bf16 (64 x 64 tiles through shared memory)
__global__ void __launch_bounds__(256)
nhwc_to_ckp_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int C, int64_t NP_, int64_t ld) {
  __shared__ float tile[64][65];
  const int64_t p0 = (int64_t)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int pl = i / 64, cl = i % 64;
    const int64_t p = p0 + pl;
    const int c = c0 + cl;
    tile[pl][cl] = (p < NP_ && c < C) ? __ldg(x + p * C + c) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 32; i += 256) {
    const int cl = i / 32, pp = (i % 32) * 2;
    const int c = c0 + cl;
    const int64_t p = p0 + pp;
    if (c >= C) continue;
    if (p + 1 < NP_ && (((uintptr_t)(out + (int64_t)c * ld + p)) & 3) == 0) {
      __nv_bfloat162 v = __floats2bfloat162_rn(tile[pp][cl], tile[pp + 1][cl]);
      *reinterpret_cast<__nv_bfloat162*>(out + (int64_t)c * ld + p) = v;
    } else {
      if (p < NP_) out[(int64_t)c * ld + p] = __float2bfloat16(tile[pp][cl]);
      if (p + 1 < NP_) out[(int64_t)c * ld + p + 1] = __float2bfloat16(tile[pp + 1][cl]);
    }
  }
}

struct AsppTaps {
  int n_taps;       // 9 * n_active
  int dh[36], dw[36];
};

// y[n][c][p] = bias_sum[c] + sum_t Z[n*P + p + shift_t][t*n_cls + c].
// A CTA owns 32 consecutive pixels; each of its 8 warps accumulates 4 of them with lane = class
// (the n_cls floats of one (pixel, tap) are contiguous in Z, all taps of a pixel are independent
// loads), then the CTA writes the [n_cls][32] tile as coalesced 128-byte NCHW rows.
template <int MAX_CLS>
__global__ void __launch_bounds__(256)
aspp_gather_kernel(const float* __restrict__ Z, const float* __restrict__ bias_sum, float* __restrict__ y, int N,
                   int H, int W, int n_cls, int NP, AsppTaps taps) {
  __shared__ float stage[MAX_CLS][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = H * W;
  const int groups_per_img = (P + 31) / 32;
  const int n = blockIdx.x / groups_per_img;
  const int p0 = (blockIdx.x - n * groups_per_img) * 32;
  const float b = lane < n_cls ? __ldg(bias_sum + lane) : 0.f;
  const float* Zn = Z + (int64_t)n * P * NP + lane;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pl = warp * 4 + i;
    const int p = p0 + pl;
    float acc = b;
    if (p < P && lane < n_cls) {
      const int h = p / W, w = p - h * W;
#pragma unroll 6
      for (int t = 0; t < taps.n_taps; ++t) {
        const int hh = h + taps.dh[t], ww = w + taps.dw[t];
        if ((unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W)
          acc += __ldg(Zn + (int64_t)(hh * W + ww) * NP + t * n_cls);
      }
    }
    if (lane < n_cls) stage[lane][pl] = acc;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < n_cls * 32; idx += 256) {
    const int c = idx >> 5, pl = idx & 31;
    if (p0 + pl < P) y[((int64_t)n * n_cls + c) * P + p0 + pl] = stage[c][pl];
  }
}

// Round-2 form of the gather: warp = ONE pixel, lane = class.  All (up to 36) taps of the pixel are independent 76-byte
// loads issued back to back (fully unrolled, the bounds test is warp-uniform), so an SM keeps 32 warps x 36 loads in
// flight instead of 24 warps x 6; a CTA owns 8 consecutive pixels and writes 32-byte runs of the NCHW rows.  The sum runs
// over the taps in the same order as before (bit-identical results).
constexpr int GATHER_PX = 8;
__global__ void __launch_bounds__(32 * GATHER_PX, 4)
aspp_gather_px_kernel(const float* __restrict__ Z, const float* __restrict__ bias_sum, float* __restrict__ y, int N,
                      int H, int W, int n_cls, int NP, const AsppTaps taps) {
  __shared__ float stage[32][GATHER_PX + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = H * W;
  const int groups_per_img = (P + GATHER_PX - 1) / GATHER_PX;
  const int n = blockIdx.x / groups_per_img;
  const int p0 = (blockIdx.x - n * groups_per_img) * GATHER_PX;
  const int p = p0 + warp;
  if (p < P && lane < n_cls) {
    const int h = p / W, w = p - h * W;
    const float* Zn = Z + (int64_t)n * P * NP + lane;
    float v[36];
#pragma unroll
    for (int t = 0; t < 36; ++t) {
      const int hh = h + taps.dh[t], ww = w + taps.dw[t];
      const bool ok = t < taps.n_taps && (unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W;
      v[t] = ok ? __ldg(Zn + (int64_t)(hh * W + ww) * NP + t * n_cls) : 0.f;
    }
    float acc = __ldg(bias_sum + lane);
#pragma unroll
    for (int t = 0; t < 36; ++t) acc += v[t];
    stage[lane][warp] = acc;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < n_cls * GATHER_PX; idx += 32 * GATHER_PX) {
    const int c = idx / GATHER_PX, pl = idx - c * GATHER_PX;
    if (p0 + pl < P) y[((int64_t)n * n_cls + c) * P + p0 + pl] = stage[c][pl];
  }
}

// dYcol[n*P + q][t*n_cls + c] = dy[n][c][q - shift_t] (0 outside).  One CTA per 32 pixels; the [32][NP] tile is
// staged in shared memory so the rows are written with coalesced 16-byte stores.
__global__ void __launch_bounds__(256)
aspp_dycols_kernel(const float* __restrict__ dy, __nv_bfloat16* __restrict__ dYcol, int N, int H, int W, int n_cls,
                   int NP, AsppTaps taps) {
  extern __shared__ __nv_bfloat16 tile[];  // [32][NP + 2]: an odd number of 32-bit words per row -> lanes (= rows) hit
  const int TP = NP + 2;                   // 32 different banks when a warp writes one column

  const int P = H * W;
  const int groups_per_img = (P + 31) / 32;
  const int n = blockIdx.x / groups_per_img;
  const int q0 = (blockIdx.x % groups_per_img) * 32;
  const int J = taps.n_taps * n_cls;
  const int ql = threadIdx.x & 31;       // lane = pixel -> coalesced reads of the dy rows
  const int q = q0 + ql;
  const int qh = q / W, qw = q - qh * W;
  const float* dyn = dy + (int64_t)n * n_cls * P;
  // zero the padding columns [J, NP) once
  for (int idx = threadIdx.x; idx < (NP - J) * 32; idx += 256) tile[(idx & 31) * TP + J + (idx >> 5)] = __float2bfloat16(0.f);
  for (int t = threadIdx.x >> 5; t < taps.n_taps; t += 8) {   // warp = tap
    const int h = qh - taps.dh[t], w = qw - taps.dw[t];
    const bool ok = q < P && (unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W;
    const float* src = dyn + (int64_t)h * W + w;
    for (int c = 0; c < n_cls; ++c) {
      const __nv_bfloat16 bv = __float2bfloat16(ok ? __ldg(src + (int64_t)c * P) : 0.f);
      const int j = t * n_cls + c;
      tile[ql * TP + j] = bv;
    }
  }
  __syncthreads();
  // rows q0..q0+31 of dYcol are contiguous (NP*2 bytes each): 32-bit words, a warp writes 128 contiguous bytes
  const int wpr = NP / 2;
  const uint32_t* srcw = reinterpret_cast<const uint32_t*>(tile);
  uint32_t* dst = reinterpret_cast<uint32_t*>(dYcol + ((int64_t)n * P + q0) * NP);
  const int rows = min(32, P - q0);
  for (int i = threadIdx.x; i < rows * wpr; i += 256) {
    const int r = i / wpr, w = i - r * wpr;
    dst[i] = srcw[r * (TP / 2) + w];
  }
}

// Round-2 form of dYcol: no shared-memory tile.  A warp owns 32 consecutive pixels x one 32-byte chunk (16 columns
// j = t*n_cls + c) of their dYcol rows: lane = pixel, so each of the 16 reads of dy[c][q - shift_t] is a coalesced
// 128-byte row segment (dy is 1.1 MB and stays in L1/L2), all 16 are independent, and the lane writes its 32 bytes with
// one 256-bit store (a full sector; rows are NP*2 = 32-byte multiples apart).  The first `extra_blocks` blocks compute the bias
// gradient db[c] = sum_p dy[c][p] (what used to be a separate 19-CTA launch) beside the rest.
__global__ void __launch_bounds__(256)
aspp_dycols_chunk_kernel(const float* __restrict__ dy, __nv_bfloat16* __restrict__ dYcol, float* __restrict__ db, int N,
                         int H, int W, int n_cls, int NP, int extra_blocks, const AsppTaps taps) {
  const int P = H * W;
  if ((int)blockIdx.x < extra_blocks) {   // the (longer, latency-bound) channel sums start first and run beside the rest
    block_channel_sum(dy, db, N, n_cls, P, blockIdx.x);
    return;
  }
  const int chunks = NP / 16;
  const int groups_per_img = (P + 31) / 32;
  const int64_t gw = (int64_t)(blockIdx.x - extra_blocks) * 8 + (threadIdx.x >> 5);   // global warp = (image, pixel group, chunk)
  if (gw >= (int64_t)N * groups_per_img * chunks) return;
  const int ck = (int)(gw % chunks);
  const int grp = (int)(gw / chunks);
  const int n = grp / groups_per_img;
  const int q = (grp - n * groups_per_img) * 32 + (threadIdx.x & 31);
  if (q >= P) return;
  const int qh = q / W, qw = q - qh * W;
  const float* dyn = dy + (int64_t)n * n_cls * P;
  int t = (ck * 16) / n_cls, c = ck * 16 - t * n_cls;                  // warp-uniform
  // the source pixel and its bounds test depend on the tap only: computed when the tap changes (a chunk spans <= 2 taps at
  // 19 classes), the class steps the pointer by one plane
  const float* src = dyn;
  bool ok = false;
  auto set_tap = [&]() {
    ok = false;
    if (t < taps.n_taps) {
      const int h = qh - taps.dh[t], w = qw - taps.dw[t];
      ok = (unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W;
      src = dyn + (int64_t)c * P + (ok ? h * W + w : 0);
    }
  };
  set_tap();
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    v[i] = ok ? __ldg(src) : 0.f;
    src += P;
    if (++c == n_cls) { c = 0; ++t; set_tap(); }
  }
  uint32_t pk[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    pk[i] = *reinterpret_cast<const uint32_t*>(&b);
  }
  __nv_bfloat16* dst = dYcol + ((int64_t)n * P + q) * NP + ck * 16;
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
               "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
               : "memory");
}

// Wp[j][ci] (bf16, j = t*n_cls + c, rows >= J zero) and WpT[ci][j] from the fp32 OIHW branch weights
struct WeightPtrs {
  const float* w[4];
};
__global__ void __launch_bounds__(256)
aspp_pack_kernel(WeightPtrs wp, __nv_bfloat16* __restrict__ Wp, __nv_bfloat16* __restrict__ WpT, int n_active,
                 int n_cls, int Cin, int NP) {
  const int64_t total = (int64_t)NP * Cin;
  const int J = n_active * 9 * n_cls;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int ci = (int)(i % Cin), j = (int)(i / Cin);
    float v = 0.f;
    if (j < J) {
      const int t = j / n_cls, c = j - t * n_cls;
      const int b = t / 9, k = t - b * 9;
      v = __ldg(wp.w[b] + ((int64_t)c * Cin + ci) * 9 + k);
    }
    const __nv_bfloat16 bv = __float2bfloat16(v);
    if (Wp) Wp[i] = bv;
    if (WpT) WpT[(int64_t)ci * NP + j] = bv;
  }
}

// dw_b[c][ci][k] = sum_s part[s][ci][(b*9+k)*n_cls + c]
struct GradPtrs {
  float* w[4];
};
// dW_b[c][ci][k] = sum_s part[s][ci][(b*9 + k)*n_cls + c].  One CTA per UDW_CI input channels: their rows of the
// partials are read coalesced (NP contiguous floats each) and summed over the splits into shared memory; the output is
// written as runs of UDW_CI * 9 contiguous floats per (branch, class).
constexpr int UDW_CI = 2;
__global__ void __launch_bounds__(256)
aspp_unpack_dw_kernel(const float* __restrict__ part, int S, GradPtrs gp, int n_active, int n_cls, int Cin,
                      int NP) {
  extern __shared__ float t[];  // [UDW_CI][NP]
  const int ci0 = blockIdx.x * UDW_CI;
  const int nci = min(UDW_CI, Cin - ci0);
  const int64_t zs = (int64_t)Cin * NP;  // elements between split-K partials
  for (int e = threadIdx.x; e < nci * NP; e += 256) {
    const float* src = part + (int64_t)ci0 * NP + e;
    float acc = 0.f;
#pragma unroll 4
    for (int s = 0; s < S; ++s) acc += __ldg(src + s * zs);
    t[e] = acc;
  }
  __syncthreads();
  const int run = nci * 9;
  for (int e = threadIdx.x; e < n_active * n_cls * run; e += 256) {
    const int bc = e / run, r = e - bc * run;   // r = ci_local * 9 + k
    const int b = bc / n_cls, c = bc - b * n_cls;
    const int cl = r / 9, k = r - cl * 9;
    if (gp.w[b]) gp.w[b][((int64_t)c * Cin + ci0) * 9 + r] = t[cl * NP + (b * 9 + k) * n_cls + c];
  }
}

// Round-2 form of the unpack: one CTA per input channel.  Its NP-float row of every split-K partial is read with 16-byte
// loads, all splits of a thread's four columns in flight at once; the 9-float runs of dW_b[c][ci][:] leave from shared memory.
__global__ void __launch_bounds__(192)
aspp_unpack_dw_row_kernel(const float* __restrict__ part, int S, GradPtrs gp, int n_active, int n_cls, int Cin, int NP) {
  extern __shared__ float t[];  // [NP]
  const int ci = blockIdx.x;
  const int64_t zs = (int64_t)Cin * NP;
  for (int e = threadIdx.x; e < NP / 4; e += 192) {
    const float4* src = reinterpret_cast<const float4*>(part + (int64_t)ci * NP) + e;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int s = 0; s < S; ++s) {
      const float4 v = ld_stream(src + s * (zs / 4));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(t)[e] = acc;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n_active * n_cls * 9; e += 192) {
    const int bc = e / 9, k = e - bc * 9;
    const int b = bc / n_cls, c = bc - b * n_cls;
    float* wb = b == 0 ? gp.w[0] : b == 1 ? gp.w[1] : b == 2 ? gp.w[2] : gp.w[3];   // (no dynamic index into the parameters)
    if (wb) wb[((int64_t)c * Cin + ci) * 9 + k] = t[(b * 9 + k) * n_cls + c];
  }
}

// Round-2 form of the weight pack: one CTA per PKW_CI input channels owns ALL their packed columns.  Reads: for every
// (branch, class) the PKW_CI*9 contiguous floats of w_b[c][ci0..][:]; writes: whole rows of WpT (16-byte stores) and 16-byte
// pieces of the Wp rows -- no 2-byte scattered stores (the old kernel wrote WpT one bf16 at a time, 1376 bytes apart).
constexpr int PKW_CI = 8;
__global__ void __launch_bounds__(256)
aspp_pack_rows_kernel(WeightPtrs wp, __nv_bfloat16* __restrict__ Wp, __nv_bfloat16* __restrict__ WpT, int n_active,
                      int n_cls, int Cin, int NP) {
  extern __shared__ __align__(16) unsigned char pk_raw[];
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(pk_raw);  // [PKW_CI][NP]
  const int ci0 = blockIdx.x * PKW_CI;
  const int J = n_active * 9 * n_cls;
  for (int e = threadIdx.x; e < PKW_CI * (NP - J); e += 256)      // padding columns
    tile[(e / (NP - J)) * NP + J + e % (NP - J)] = __float2bfloat16(0.f);
  // per (branch, class): PKW_CI * 9 = 72 contiguous floats = 18 float4 (288-byte aligned runs)
  constexpr int RUN4 = PKW_CI * 9 / 4;
#pragma unroll 6
  for (int e = threadIdx.x; e < n_active * n_cls * RUN4; e += 256) {
    const int bc = e / RUN4, r4 = e - bc * RUN4;
    const int b = bc / n_cls, c = bc - b * n_cls;
    const float* wb = b == 0 ? wp.w[0] : b == 1 ? wp.w[1] : b == 2 ? wp.w[2] : wp.w[3];
    const float4 v = ld_stream(reinterpret_cast<const float4*>(wb + ((int64_t)c * Cin + ci0) * 9) + r4);
    const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r4 * 4 + i;
      const int cl = r / 9, k = r - cl * 9;
      tile[cl * NP + (b * 9 + k) * n_cls + c] = __float2bfloat16(f[i]);
    }
  }
  __syncthreads();
  if (WpT) {   // rows ci0 .. ci0+PKW_CI-1 of WpT are the tile itself
    const uint4* src = reinterpret_cast<const uint4*>(tile);
    uint4* dst = reinterpret_cast<uint4*>(WpT + (int64_t)ci0 * NP);
    for (int e = threadIdx.x; e < PKW_CI * NP / 8; e += 256) dst[e] = src[e];
  }
  if (Wp) {    // Wp[j][ci0 .. ci0+7]: one 16-byte store per packed row
    for (int j = threadIdx.x; j < NP; j += 256) {
      uint32_t w4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t lo = *reinterpret_cast<const unsigned short*>(tile + (2 * i) * NP + j);
        const uint32_t hi = *reinterpret_cast<const unsigned short*>(tile + (2 * i + 1) * NP + j);
        w4[i] = lo | (hi << 16);
      }
      *reinterpret_cast<uint4*>(Wp + (int64_t)j * Cin + ci0) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
  }
}

// column sums over pixels of an NCHW tensor: out[o] = sum_{n,p} src[n][o][p]
__global__ void __launch_bounds__(256)
nchw_channel_sum_kernel(const float* __restrict__ src, float* __restrict__ out, int N, int O, int P) {
  block_channel_sum(src, out, N, O, P, blockIdx.x);
}

int channel_sum_nchw(const float* src, float* out, int N, int O, int P, cudaStream_t st) {
  prof::Scope ps("channel_sum", 0, 4.0 * N * O * P, st);
  nchw_channel_sum_kernel<<<O, 256, 0, st>>>(src, out, N, O, P);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

// ---- host orchestration --------------------------------------------------------------------

// BLOCK_N of the forward GEMM: least padding of N (176 for the 688 packed taps), except that with CTA pairs the
// 256-wide MMA is ~20 % cheaper per MAC than the 176-wide one (tools/gemm_probe.py) and wins, in spite of padding
// 688 -> 768, once there are at least two full rounds of 256 x 256 pair tiles.
static int pick_block_n(int n, int64_t rows) {
  const int cand[3] = {128, 176, 256};
  int best = 256;
  int64_t best_pad = -1;
  for (int i = 0; i < 3; ++i) {
    int64_t pad = round_up(n, cand[i]);
    if (best_pad < 0 || pad < best_pad || (pad == best_pad && cand[i] > best)) {
      best = cand[i];
      best_pad = pad;
    }
  }
  if (umma::cluster_size(umma::MODE_GEMM, 256) == 2) {
    const int64_t pair_tiles = cdiv(cdiv(rows, (int64_t)128), (int64_t)2) * cdiv(n, 256);
    if (pair_tiles >= sm_count()) best = 256;
  }
  return best;
}

static int make_taps(AsppTaps& t, const int* dil, int n_active, int W) {
  (void)W;
  t.n_taps = 9 * n_active;
  for (int b = 0; b < n_active; ++b)
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        t.dh[b * 9 + kh * 3 + kw] = (kh - 1) * dil[b];
        t.dw[b * 9 + kh * 3 + kw] = (kw - 1) * dil[b];
      }
  return ASN_OK;
}

// ASN_GLUE=0 selects the round-1 forms of the layout kernels (gather, dYcol, unpack, pack) for A/B measurements
static bool glue_v2() {
  static const bool on = !(getenv("ASN_GLUE") != nullptr && getenv("ASN_GLUE")[0] == '0');
  return on;
}

struct AsppWs {
  size_t x_nhwc, z, dycol, dwpart, total;
  int NP, S;
};

// MN-major B tiles are made of 64-column boxes.  CTA pairs (256 x 256 tiles, each CTA loading 128 of the 256 columns):
// 3 x 256 >= 688; single CTAs: 4 x 192.
static int aspp_wgrad_bn(int Cin) {
  return umma::cluster_size(umma::MODE_GEMM_MN, 256) == 2 && cdiv(Cin, 128) % 2 == 0 ? 256 : 192;
}

static AsppWs aspp_ws(int N, int Cin, int H, int W, int n_cls, int n_active) {
  AsppWs w;
  const int64_t P = (int64_t)N * H * W;
  w.NP = (int)round_up((int64_t)9 * n_active * n_cls, 16);
  // split-K of the wgrad GEMM: one wave of persistent CTAs
  const int bn = aspp_wgrad_bn(Cin);
  const int tiles = bn == 256 ? cdiv(Cin, 256) * cdiv(w.NP, bn) : cdiv(Cin, 128) * cdiv(w.NP, bn);  // (pair) tiles
  const int slots = bn == 256 ? sm_count() / 2 : sm_count();
  int S = slots / tiles;
  const int k_steps = cdiv(P, 64);
  if (S > k_steps / 4) S = k_steps / 4;
  if (S < 1) S = 1;
  w.S = umma::effective_split((int)P, S);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) / 256 * 256;
    return o;
  };
  w.x_nhwc = take((size_t)P * Cin * 2);   // forward only, when the caller does not keep the bf16 copy
  w.z = take((size_t)P * w.NP * 4);
  w.dycol = take((size_t)P * w.NP * 2);
  w.dwpart = take((size_t)w.S * Cin * w.NP * 4);
  w.total = off;
  return w;
}

}  // namespace asn

using namespace asn;

extern "C" int asn_aspp_np(int n_cls, int n_active) { return (int)round_up((int64_t)9 * n_active * n_cls, 16); }

extern "C" int asn_aspp_pack_weights(const float* const* w_oihw, int n_active, int n_cls, int Cin, void* wp_bf16,
                                     void* wpt_bf16, void* stream) {
  ASN_CHECK_ARG(w_oihw && n_active >= 1 && n_active <= 4 && n_cls >= 1 && n_cls <= 32 && Cin >= 1,
                "asn_aspp_pack_weights: bad argument");
  WeightPtrs wp{};
  for (int b = 0; b < n_active; ++b) {
    ASN_CHECK_ARG(w_oihw[b], "asn_aspp_pack_weights: null branch weight %d", b);
    wp.w[b] = w_oihw[b];
  }
  const int NP = asn_aspp_np(n_cls, n_active);
  prof::Scope ps("aspp_pack_weights", 0, (36.0 * n_active / 4 + 4.0) * NP * Cin, static_cast<cudaStream_t>(stream));
  const bool aligned = ((reinterpret_cast<uintptr_t>(wp_bf16) | reinterpret_cast<uintptr_t>(wpt_bf16)) & 15) == 0;
  bool w_aligned = true;
  for (int b = 0; b < n_active; ++b) w_aligned = w_aligned && (reinterpret_cast<uintptr_t>(w_oihw[b]) & 15) == 0;
  if (glue_v2() && Cin % PKW_CI == 0 && aligned && w_aligned) {
    const size_t smem = (size_t)PKW_CI * NP * sizeof(__nv_bfloat16);
    aspp_pack_rows_kernel<<<Cin / PKW_CI, 256, smem, static_cast<cudaStream_t>(stream)>>>(
        wp, static_cast<__nv_bfloat16*>(wp_bf16), static_cast<__nv_bfloat16*>(wpt_bf16), n_active, n_cls, Cin, NP);
  } else {
    aspp_pack_kernel<<<full_grid((int64_t)NP * Cin, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        wp, static_cast<__nv_bfloat16*>(wp_bf16), static_cast<__nv_bfloat16*>(wpt_bf16), n_active, n_cls, Cin, NP);
  }
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

extern "C" size_t asn_aspp_workspace_bytes(int N, int Cin, int H, int W, int n_cls, int n_active) {
  return aspp_ws(N, Cin, H, W, n_cls, n_active).total;
}

static int aspp_check(int N, int Cin, int H, int W, int n_cls, int n_active, const int* dil) {
  ASN_CHECK_ARG(N > 0 && H > 0 && W > 0, "aspp: bad shape");
  ASN_CHECK_ARG(Cin % 8 == 0 && Cin >= 64, "aspp: Cin=%d must be a multiple of 8 and >= 64", Cin);
  ASN_CHECK_ARG(n_cls >= 1 && n_cls <= 32, "aspp: n_cls=%d outside [1,32]", n_cls);
  ASN_CHECK_ARG(n_active >= 1 && n_active <= 4 && dil, "aspp: n_active=%d outside [1,4]", n_active);
  ASN_CHECK_ARG((int64_t)N * H * W < (1LL << 30), "aspp: too many pixels");
  return ASN_OK;
}

extern "C" int asn_aspp_fwd(const float* x_nchw, int x_channels_last, void* x_bf16_keep, const void* wp_bf16,
                            const float* bias_sum, float* y_nchw, int N, int Cin, int H, int W, int n_cls,
                            const int* dil_host, int n_active, void* workspace, size_t workspace_bytes, void* stream) {
  ASN_CHECK_ARG(x_nchw && wp_bf16 && bias_sum && y_nchw && workspace, "asn_aspp_fwd: null pointer");
  int rc = aspp_check(N, Cin, H, W, n_cls, n_active, dil_host);
  if (rc) return rc;
  const AsppWs ws = aspp_ws(N, Cin, H, W, n_cls, n_active);
  if (workspace_bytes < ws.total) {
    set_error("asn_aspp_fwd: workspace %zu < %zu", workspace_bytes, ws.total);
    return ASN_EWORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* base = static_cast<uint8_t*>(workspace);
  __nv_bfloat16* xn = x_bf16_keep ? static_cast<__nv_bfloat16*>(x_bf16_keep)
                                  : reinterpret_cast<__nv_bfloat16*>(base + ws.x_nhwc);
  float* Z = reinterpret_cast<float*>(base + ws.z);
  const int P = H * W;
  if (x_channels_last) {
    ASN_CHECK_ARG((reinterpret_cast<uintptr_t>(x_nchw) & 15) == 0, "asn_aspp_fwd: channels_last x must be 16-byte aligned");
    prof::Scope ps("aspp_x_cast_bf16", 0, 6.0 * N * P * Cin, st);
    const int64_t n4 = (int64_t)N * P * Cin / 4;
    nhwc_f32_to_bf16_kernel<<<full_grid(n4, 256), 256, 0, st>>>(x_nchw, xn, n4);
    ASN_LAUNCH_CHECK();
  } else {
    prof::Scope ps("aspp_x_to_nhwc_bf16", 0, 6.0 * N * P * Cin, st);
    nchw_to_nhwc_bf16_kernel<<<dim3(cdiv(P, 64), cdiv(Cin, 64), N), 256, 0, st>>>(x_nchw, xn, Cin, P);
    ASN_LAUNCH_CHECK();
  }
  // algorithmic flops: 2 * px * (9*n_active*n_cls) * Cin  (N padding not counted)
  rc = umma::gemm_tn(xn, wp_bf16, Z, N * P, ws.NP, Cin, Cin, Cin, ws.NP, 1, 0, pick_block_n(ws.NP, (int64_t)N * P), st,
                     "aspp_fwd_gemm", 2.0 * N * P * (9.0 * n_active * n_cls) * Cin);
  if (rc) return rc;
  AsppTaps taps;
  make_taps(taps, dil_host, n_active, W);
  const int groups = N * cdiv(P, 32);
  prof::Scope ps("aspp_gather", 0, 4.0 * N * P * (9.0 * n_active * n_cls + n_cls), st);
  if (glue_v2())
    aspp_gather_px_kernel<<<N * cdiv(P, GATHER_PX), 32 * GATHER_PX, 0, st>>>(Z, bias_sum, y_nchw, N, H, W, n_cls, ws.NP, taps);
  else
    aspp_gather_kernel<32><<<groups, 256, 0, st>>>(Z, bias_sum, y_nchw, N, H, W, n_cls, ws.NP, taps);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

extern "C" int asn_aspp_bwd(const void* x_bf16, int dx_channels_last, const void* wpt_bf16, const float* dy_nchw,
                            float* dx_nchw, float* const* dw_oihw, float* db, int N, int Cin, int H, int W, int n_cls,
                            const int* dil_host, int n_active, void* workspace, size_t workspace_bytes,
                            void* stream) {
  ASN_CHECK_ARG(dy_nchw && workspace, "asn_aspp_bwd: null pointer");
  ASN_CHECK_ARG(!dx_nchw || wpt_bf16, "asn_aspp_bwd: dx needs the transposed weight pack");
  ASN_CHECK_ARG(!dw_oihw || x_bf16, "asn_aspp_bwd: dw needs the bf16 activations kept by the forward");
  int rc = aspp_check(N, Cin, H, W, n_cls, n_active, dil_host);
  if (rc) return rc;
  const AsppWs ws = aspp_ws(N, Cin, H, W, n_cls, n_active);
  if (workspace_bytes < ws.total) {
    set_error("asn_aspp_bwd: workspace %zu < %zu", workspace_bytes, ws.total);
    return ASN_EWORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* base = static_cast<uint8_t*>(workspace);
  __nv_bfloat16* dycol = reinterpret_cast<__nv_bfloat16*>(base + ws.dycol);
  const int P = H * W;
  AsppTaps taps;
  make_taps(taps, dil_host, n_active, W);
  const double flops = 2.0 * N * P * (9.0 * n_active * n_cls) * Cin;
  bool db_done = false;
  if (dx_nchw || dw_oihw) {
    prof::Scope ps("aspp_dy_cols", 0, 2.0 * N * P * ws.NP + 4.0 * N * P * n_cls, st);
    if (glue_v2() && (reinterpret_cast<uintptr_t>(dycol) & 31) == 0) {
      // the bias gradient rides along as n_cls extra CTAs of the same launch
      const int main_blocks = cdiv((int64_t)N * cdiv(P, 32) * (ws.NP / 16), 8);
      const int extra = db ? n_cls : 0;
      aspp_dycols_chunk_kernel<<<main_blocks + extra, 256, 0, st>>>(dy_nchw, dycol, db, N, H, W, n_cls, ws.NP, extra, taps);
      db_done = db != nullptr;
    } else {
      const size_t smem = (size_t)32 * (ws.NP + 2) * 2;
      if (smem > 48 * 1024)
        ASN_CUDA(cudaFuncSetAttribute(aspp_dycols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      aspp_dycols_kernel<<<N * cdiv(P, 32), 256, smem, st>>>(dy_nchw, dycol, N, H, W, n_cls, ws.NP, taps);
    }
    ASN_LAUNCH_CHECK();
  }
  if (dx_nchw && dx_channels_last) {
    // dX[N*P][Cin] = dYcol[N*P][NP] . WpT[Cin][NP]^T : the gradient lands in channels_last, all images at once
    rc = umma::gemm_tn(dycol, wpt_bf16, dx_nchw, N * P, Cin, ws.NP, ws.NP, ws.NP, Cin, 1, 0, 256, st,
                       "aspp_dgrad_gemm", flops);
    if (rc) return rc;
  } else if (dx_nchw) {
    // dX^T[Cin][P] = WpT[Cin][NP] . dYcol[P][NP]^T per image: lands directly in NCHW
    for (int n = 0; n < N; ++n) {
      rc = umma::gemm_tn(wpt_bf16, dycol + (int64_t)n * P * ws.NP, dx_nchw + (int64_t)n * Cin * P, Cin, P, ws.NP,
                         ws.NP, ws.NP, P, 1, 0, 256, st, "aspp_dgrad_gemm", flops / N);
      if (rc) return rc;
    }
  }
  if (dw_oihw) {
    // dWp[Cin][NP] = X[N*P][Cin]^T . dYcol[N*P][NP]: both operands as they are (MN-major), split-K over the pixels
    float* part = reinterpret_cast<float*>(base + ws.dwpart);
    rc = umma::gemm_nt_mn(x_bf16, dycol, part, Cin, ws.NP, N * P, Cin, ws.NP, ws.NP, ws.S, (long long)Cin * ws.NP,
                          aspp_wgrad_bn(Cin), st, "aspp_wgrad_gemm", flops);
    if (rc) return rc;
    GradPtrs gp{};
    for (int b = 0; b < n_active; ++b) gp.w[b] = dw_oihw[b];
    prof::Scope ps("aspp_unpack_dw", 0, 4.0 * Cin * ws.NP * (ws.S + 1), st);
    if (glue_v2() && (reinterpret_cast<uintptr_t>(part) & 15) == 0)
      aspp_unpack_dw_row_kernel<<<Cin, 192, (size_t)ws.NP * sizeof(float), st>>>(part, ws.S, gp, n_active, n_cls, Cin, ws.NP);
    else
      aspp_unpack_dw_kernel<<<cdiv(Cin, UDW_CI), 256, (size_t)UDW_CI * ws.NP * sizeof(float), st>>>(
          part, ws.S, gp, n_active, n_cls, Cin, ws.NP);
    ASN_LAUNCH_CHECK();
  }
  if (db && !db_done) {
    rc = channel_sum_nchw(dy_nchw, db, N, n_cls, P, st);
    if (rc) return rc;
  }
  return ASN_OK;
}
