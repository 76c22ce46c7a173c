// K3 (softmax + cross entropy, ignore label), K4 (channel softmax), K6 (GAN losses).
// All HBM-bound (K6: latency-bound).  Layout: fp32 NCHW; a thread owns VEC=4 consecutive
// pixels of one image and walks the C channel planes with 16-byte loads, so every warp
// request is a fully coalesced 512-byte segment; the C values per pixel stay in registers
// (C = 19 is specialised; other class counts take a three-pass generic kernel).
#include "common.cuh"
#include "ce_common.cuh"

namespace asn {

constexpr int PW_THREADS = 128;

template <int VEC>
struct PixVec;
template <>
struct PixVec<4> {
  static __device__ __forceinline__ void load(const float* p, float* v) {
    float4 t = ld_stream(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* v) {
    st_stream(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
};
template <>
struct PixVec<1> {
  static __device__ __forceinline__ void load(const float* p, float* v) { v[0] = __ldg(p); }
  static __device__ __forceinline__ void store(float* p, const float* v) { p[0] = v[0]; }
};

__device__ __forceinline__ void block_accumulate(double a, double b, long long c, long long d,
                                                 CeStats* stats) {
  __shared__ double sa[PW_THREADS / 32], sb[PW_THREADS / 32];
  __shared__ long long sc[PW_THREADS / 32], sd[PW_THREADS / 32];
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c); d = warp_sum(d);
  int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sa[wid] = a; sb[wid] = b; sc[wid] = c; sd[wid] = d; }
  __syncthreads();
  if (wid == 0) {
    a = lane < PW_THREADS / 32 ? sa[lane] : 0.0;
    b = lane < PW_THREADS / 32 ? sb[lane] : 0.0;
    c = lane < PW_THREADS / 32 ? sc[lane] : 0;
    d = lane < PW_THREADS / 32 ? sd[lane] : 0;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c); d = warp_sum(d);
    if (lane == 0) {
      if (a != 0.0 || a != a) atomicAdd(&stats->loss_sum, a);
      if (b != 0.0) atomicAdd(&stats->weight_sum, b);
      if (c) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_valid), (unsigned long long)c);
      if (d) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_bad), (unsigned long long)d);
    }
  }
}

// ---- C known at compile time: values live in registers --------------------------------
template <int C, int VEC, bool BWD>
__global__ void __launch_bounds__(PW_THREADS)
ce_kernel(const float* __restrict__ z, const long long* __restrict__ y, int N, int HW, int ignore,
          int mask_negative, const float* __restrict__ cw, int size_average, CeStats* stats,
          const float* __restrict__ gscale, float* __restrict__ dz) {
  const int gpi = HW / VEC;  // pixel groups per image
  const int64_t total = (int64_t)N * gpi;
  double loss = 0.0, wsum = 0.0;
  long long nvalid = 0, nbad = 0;
  float bscale = 1.f;
  if (BWD) {
    bscale = gscale ? __ldg(gscale) : 1.f;
    if (size_average) bscale = (float)((double)bscale / stats->weight_sum);
  }
  for (int64_t g = (int64_t)blockIdx.x * PW_THREADS + threadIdx.x; g < total;
       g += (int64_t)gridDim.x * PW_THREADS) {
    const int n = (int)(g / gpi);
    const int p = (int)(g % gpi) * VEC;
    const float* zp = z + (int64_t)n * C * HW + p;
    float v[C][VEC];
#pragma unroll
    for (int c = 0; c < C; ++c) PixVec<VEC>::load(zp + (int64_t)c * HW, v[c]);
    long long lab[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) lab[k] = __ldg(y + (int64_t)n * HW + p + k);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float m = v[0][k];
#pragma unroll
      for (int c = 1; c < C; ++c) m = fmaxf(m, v[c][k]);
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) s += expf(v[c][k] - m);
      const int cls = classify_label(lab[k], C, ignore, mask_negative);
      const int yi = cls == 1 ? (int)lab[k] : -1;
      const float wgt = cls == 1 ? (cw ? __ldg(cw + yi) : 1.f) : 0.f;
      if (!BWD) {
        float zy = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) zy = (c == yi) ? v[c][k] : zy;
        if (cls == 1) {
          loss += (double)(wgt * ((m + logf(s)) - zy));
          wsum += (double)wgt;
          ++nvalid;
        } else if (cls < 0) {
          ++nbad;
        }
      } else {
        const float inv = 1.f / s;
        const float f = wgt * bscale;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float pc = expf(v[c][k] - m) * inv;
          v[c][k] = cls == 1 ? f * (pc - (c == yi ? 1.f : 0.f)) : 0.f;
        }
      }
    }
    if (BWD) {
      float* dp = dz + (int64_t)n * C * HW + p;
#pragma unroll
      for (int c = 0; c < C; ++c) PixVec<VEC>::store(dp + (int64_t)c * HW, v[c]);
    }
  }
  if (!BWD) block_accumulate(loss, wsum, nvalid, nbad, stats);
}

// ---- generic class count: three passes over the channel planes (L1/L2 resident) ------
template <bool BWD>
__global__ void __launch_bounds__(PW_THREADS)
ce_generic_kernel(const float* __restrict__ z, const long long* __restrict__ y, int N, int C, int HW,
                  int ignore, int mask_negative, const float* __restrict__ cw, int size_average,
                  CeStats* stats, const float* __restrict__ gscale, float* __restrict__ dz) {
  const int64_t total = (int64_t)N * HW;
  double loss = 0.0, wsum = 0.0;
  long long nvalid = 0, nbad = 0;
  float bscale = 1.f;
  if (BWD) {
    bscale = gscale ? __ldg(gscale) : 1.f;
    if (size_average) bscale = (float)((double)bscale / stats->weight_sum);
  }
  for (int64_t g = (int64_t)blockIdx.x * PW_THREADS + threadIdx.x; g < total;
       g += (int64_t)gridDim.x * PW_THREADS) {
    const int n = (int)(g / HW);
    const int p = (int)(g % HW);
    const float* zp = z + (int64_t)n * C * HW + p;
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, zp[(int64_t)c * HW]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(zp[(int64_t)c * HW] - m);
    const long long lab = y[(int64_t)n * HW + p];
    const int cls = classify_label(lab, C, ignore, mask_negative);
    const int yi = cls == 1 ? (int)lab : -1;
    const float wgt = cls == 1 ? (cw ? cw[yi] : 1.f) : 0.f;
    if (!BWD) {
      if (cls == 1) {
        loss += (double)(wgt * ((m + logf(s)) - zp[(int64_t)yi * HW]));
        wsum += (double)wgt;
        ++nvalid;
      } else if (cls < 0) {
        ++nbad;
      }
    } else {
      float* dp = dz + (int64_t)n * C * HW + p;
      const float inv = 1.f / s, f = wgt * bscale;
      for (int c = 0; c < C; ++c) {
        float pc = expf(zp[(int64_t)c * HW] - m) * inv;
        dp[(int64_t)c * HW] = cls == 1 ? f * (pc - (c == yi ? 1.f : 0.f)) : 0.f;
      }
    }
  }
  if (!BWD) block_accumulate(loss, wsum, nvalid, nbad, stats);
}

__global__ void ce_finalize_kernel(const CeStats* stats, int size_average, float* loss) {
  double v = size_average ? stats->loss_sum / stats->weight_sum : stats->loss_sum;
  if (stats->n_bad) v = __longlong_as_double(0x7ff8000000000000LL);  // out-of-bounds target
  *loss = (float)v;
}

// ---- K4 softmax over channels -----------------------------------------------------------
template <int C, int VEC, bool BWD>
__global__ void __launch_bounds__(PW_THREADS)
softmax_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int N,
               int HW) {
  // fwd: a = z, out = p.   bwd: a = p, b = dp, out = dz = p * (dp - sum_c p*dp)
  const int gpi = HW / VEC;
  const int64_t total = (int64_t)N * gpi;
  for (int64_t g = (int64_t)blockIdx.x * PW_THREADS + threadIdx.x; g < total;
       g += (int64_t)gridDim.x * PW_THREADS) {
    const int n = (int)(g / gpi);
    const int p = (int)(g % gpi) * VEC;
    const int64_t off = (int64_t)n * C * HW + p;
    float v[C][VEC];
#pragma unroll
    for (int c = 0; c < C; ++c) PixVec<VEC>::load(a + off + (int64_t)c * HW, v[c]);
    if (!BWD) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        float m = v[0][k];
#pragma unroll
        for (int c = 1; c < C; ++c) m = fmaxf(m, v[c][k]);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          v[c][k] = expf(v[c][k] - m);
          s += v[c][k];
        }
        const float inv = 1.f / s;
#pragma unroll
        for (int c = 0; c < C; ++c) v[c][k] *= inv;
      }
    } else {
      float dot[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) dot[k] = 0.f;
      // second operand is streamed channel by channel twice (dot, then result): keep p in regs
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float d[VEC];
        PixVec<VEC>::load(b + off + (int64_t)c * HW, d);
#pragma unroll
        for (int k = 0; k < VEC; ++k) dot[k] += v[c][k] * d[k];
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float d[VEC];
        PixVec<VEC>::load(b + off + (int64_t)c * HW, d);
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[c][k] = v[c][k] * (d[k] - dot[k]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) PixVec<VEC>::store(out + off + (int64_t)c * HW, v[c]);
  }
}

template <bool BWD>
__global__ void __launch_bounds__(PW_THREADS)
softmax_generic_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                       int N, int C, int HW) {
  const int64_t total = (int64_t)N * HW;
  for (int64_t g = (int64_t)blockIdx.x * PW_THREADS + threadIdx.x; g < total;
       g += (int64_t)gridDim.x * PW_THREADS) {
    const int64_t off = (int64_t)(g / HW) * C * HW + (g % HW);
    if (!BWD) {
      float m = -INFINITY;
      for (int c = 0; c < C; ++c) m = fmaxf(m, a[off + (int64_t)c * HW]);
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += expf(a[off + (int64_t)c * HW] - m);
      const float inv = 1.f / s;
      for (int c = 0; c < C; ++c) out[off + (int64_t)c * HW] = expf(a[off + (int64_t)c * HW] - m) * inv;
    } else {
      float dot = 0.f;
      for (int c = 0; c < C; ++c) dot += a[off + (int64_t)c * HW] * b[off + (int64_t)c * HW];
      for (int c = 0; c < C; ++c)
        out[off + (int64_t)c * HW] = a[off + (int64_t)c * HW] * (b[off + (int64_t)c * HW] - dot);
    }
  }
}

// ---- K6 GAN losses: one CTA, fixed reduction tree -> deterministic ----------------------
constexpr int GAN_THREADS = 512;
__global__ void __launch_bounds__(GAN_THREADS)
gan_loss_kernel(const float* __restrict__ x, long long n, float t, int kind, float grad_scale,
                float* __restrict__ loss, float* __restrict__ dx) {
  __shared__ double part[GAN_THREADS / 32];
  double acc = 0.0;
  const float invn = 1.f / (float)n;
  for (long long i = threadIdx.x; i < n; i += GAN_THREADS) {
    float v = x[i];
    float l, g;
    if (kind == ASN_GAN_BCE) {
      // max(x,0) - x*t + log1p(exp(-|x|)) ; d/dx = sigmoid(x) - t
      float e = expf(-fabsf(v));
      l = fmaxf(v, 0.f) - v * t + log1pf(e);
      float sig = v >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
      g = sig - t;
    } else {
      float d = v - t;
      l = d * d;
      g = 2.f * d;
    }
    acc += (double)l;
    if (dx) dx[i] = grad_scale * g * invn;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < GAN_THREADS / 32 ? part[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) *loss = (float)(v / (double)n);
  }
}

}  // namespace asn

using namespace asn;

template <bool BWD>
static int launch_ce(const float* z, const int64_t* y, int N, int C, int H, int W, int ignore,
                     int mask_negative, const float* cw, int size_average, void* stats,
                     const float* gscale, float* dz, cudaStream_t st) {
  const int HW = H * W;
  CeStats* s = static_cast<CeStats*>(stats);
  const long long* yl = reinterpret_cast<const long long*>(y);
  const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(z) & 15) == 0) &&
                   (!BWD || (reinterpret_cast<uintptr_t>(dz) & 15) == 0);
  prof::Scope ps(BWD ? "softmax_ce_bwd" : "softmax_ce_fwd", 0,
                 (double)N * HW * (4.0 * C * (BWD ? 2 : 1) + 8.0), st);
  if (C == 19 && vec) {
    int grid = full_grid((int64_t)N * (HW / 4), PW_THREADS);
    ce_kernel<19, 4, BWD><<<grid, PW_THREADS, 0, st>>>(z, yl, N, HW, ignore, mask_negative, cw, size_average, s, gscale, dz);
  } else if (C == 19) {
    int grid = full_grid((int64_t)N * HW, PW_THREADS);
    ce_kernel<19, 1, BWD><<<grid, PW_THREADS, 0, st>>>(z, yl, N, HW, ignore, mask_negative, cw, size_average, s, gscale, dz);
  } else {
    int grid = full_grid((int64_t)N * HW, PW_THREADS);
    ce_generic_kernel<BWD><<<grid, PW_THREADS, 0, st>>>(z, yl, N, C, HW, ignore, mask_negative, cw, size_average, s, gscale, dz);
  }
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

extern "C" int asn_softmax_ce_fwd(const float* z, const int64_t* y, int N, int C, int H, int W,
                                  int ignore_label, int mask_negative, const float* class_weight,
                                  int size_average, void* stats, float* loss, void* stream) {
  ASN_CHECK_ARG(z && y && stats && loss, "asn_softmax_ce_fwd: null pointer");
  ASN_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0, "asn_softmax_ce_fwd: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ASN_CUDA(cudaMemsetAsync(stats, 0, sizeof(CeStats), st));
  int rc = launch_ce<false>(z, y, N, C, H, W, ignore_label, mask_negative, class_weight, size_average, stats, nullptr, nullptr, st);
  if (rc) return rc;
  ce_finalize_kernel<<<1, 1, 0, st>>>(static_cast<const CeStats*>(stats), size_average, loss);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

extern "C" int asn_softmax_ce_bwd(const float* z, const int64_t* y, int N, int C, int H, int W,
                                  int ignore_label, int mask_negative, const float* class_weight,
                                  int size_average, const void* stats, const float* gscale, float* dz,
                                  void* stream) {
  ASN_CHECK_ARG(z && y && stats && dz, "asn_softmax_ce_bwd: null pointer");
  ASN_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0, "asn_softmax_ce_bwd: bad shape");
  return launch_ce<true>(z, y, N, C, H, W, ignore_label, mask_negative, class_weight, size_average,
                         const_cast<void*>(stats), gscale, dz, static_cast<cudaStream_t>(stream));
}

template <bool BWD>
static int launch_softmax(const float* a, const float* b, float* out, int N, int C, int H, int W,
                          cudaStream_t st) {
  const int HW = H * W;
  const bool vec = (HW % 4 == 0) && (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                                        reinterpret_cast<uintptr_t>(out)) & 15) == 0);
  prof::Scope ps(BWD ? "softmax_bwd" : "softmax_fwd", 0, 4.0 * N * C * HW * (BWD ? 3 : 2), st);
  if (C == 19 && vec) {
    softmax_kernel<19, 4, BWD><<<full_grid((int64_t)N * (HW / 4), PW_THREADS), PW_THREADS, 0, st>>>(a, b, out, N, HW);
  } else if (C == 19) {
    softmax_kernel<19, 1, BWD><<<full_grid((int64_t)N * HW, PW_THREADS), PW_THREADS, 0, st>>>(a, b, out, N, HW);
  } else {
    softmax_generic_kernel<BWD><<<full_grid((int64_t)N * HW, PW_THREADS), PW_THREADS, 0, st>>>(a, b, out, N, C, HW);
  }
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

extern "C" int asn_softmax_fwd(const float* z, int N, int C, int H, int W, float* p, void* stream) {
  ASN_CHECK_ARG(z && p, "asn_softmax_fwd: null pointer");
  ASN_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0, "asn_softmax_fwd: bad shape");
  return launch_softmax<false>(z, nullptr, p, N, C, H, W, static_cast<cudaStream_t>(stream));
}

extern "C" int asn_softmax_bwd(const float* p, const float* dp, int N, int C, int H, int W, float* dz,
                               void* stream) {
  ASN_CHECK_ARG(p && dp && dz, "asn_softmax_bwd: null pointer");
  ASN_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0, "asn_softmax_bwd: bad shape");
  return launch_softmax<true>(p, dp, dz, N, C, H, W, static_cast<cudaStream_t>(stream));
}

extern "C" int asn_gan_loss_fwd_bwd(const float* x, int64_t n, float target, int kind, float grad_scale,
                                    float* loss, float* dx, void* stream) {
  ASN_CHECK_ARG(x && loss, "asn_gan_loss_fwd_bwd: null pointer");
  ASN_CHECK_ARG(n > 0, "asn_gan_loss_fwd_bwd: empty input");
  ASN_CHECK_ARG(kind == ASN_GAN_BCE || kind == ASN_GAN_MSE, "asn_gan_loss_fwd_bwd: unknown kind %d", kind);
  prof::Scope ps("gan_loss", 0, 8.0 * n, static_cast<cudaStream_t>(stream));
  gan_loss_kernel<<<1, GAN_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, (long long)n, target, kind, grad_scale, loss, dx);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}
