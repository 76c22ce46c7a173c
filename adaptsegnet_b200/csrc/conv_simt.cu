// fp32 convolution on CUDA cores (ASN_PREC_FP32): the "fp32 mode" (1e-4) of K1 / K5 and the
// on-device cross-check for the tcgen05 kernels.  One tiled implicit-GEMM kernel serves
// forward, data-gradient and weight-gradient through three gather functors:
//   FWD  : C[m=(n,p,q)][j=o]      = sum_k A[m][k=(c,r,t)] * w[o][k]
//   DGRAD: C[m=(n,ih,iw)][j=c]    = sum_k dy[m @ k=(o,r,t)] * w[o][c][r][t]
//   WGRAD: C[m=o][j=(c,r,t)]      = sum_k dy[o][k=(n,p,q)] * x[k @ j]      (split-K, atomics)
// 64x64 output tile per CTA, K step 16, 4x4 register micro-tile per thread.
#include "common.cuh"

namespace asn {

struct ConvGeom {
  int N, C, H, W, O, KH, KW, stride, pad, dil, OH, OW;
};

enum { MODE_FWD = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };

constexpr int BM = 64, BN = 64, BK = 16, CONV_THREADS = 256;

template <int MODE>
__device__ __forceinline__ float load_a(const float* __restrict__ src, const ConvGeom& g, int m, int k,
                                        int M, int K) {
  if (m >= M || k >= K) return 0.f;
  if (MODE == MODE_FWD) {
    int q = m % g.OW, p = (m / g.OW) % g.OH, n = m / (g.OW * g.OH);
    int t = k % g.KW, r = (k / g.KW) % g.KH, c = k / (g.KW * g.KH);
    int ih = p * g.stride - g.pad + r * g.dil, iw = q * g.stride - g.pad + t * g.dil;
    if ((unsigned)ih >= (unsigned)g.H || (unsigned)iw >= (unsigned)g.W) return 0.f;
    return __ldg(src + (((int64_t)n * g.C + c) * g.H + ih) * g.W + iw);
  } else if (MODE == MODE_DGRAD) {
    int iw = m % g.W, ih = (m / g.W) % g.H, n = m / (g.W * g.H);
    int t = k % g.KW, r = (k / g.KW) % g.KH, o = k / (g.KW * g.KH);
    int ph = ih + g.pad - r * g.dil, pw = iw + g.pad - t * g.dil;
    if (ph < 0 || pw < 0 || ph % g.stride || pw % g.stride) return 0.f;
    ph /= g.stride;
    pw /= g.stride;
    if (ph >= g.OH || pw >= g.OW) return 0.f;
    return __ldg(src + (((int64_t)n * g.O + o) * g.OH + ph) * g.OW + pw);
  } else {
    int q = k % g.OW, p = (k / g.OW) % g.OH, n = k / (g.OW * g.OH);
    return __ldg(src + (((int64_t)n * g.O + m) * g.OH + p) * g.OW + q);
  }
}

template <int MODE>
__device__ __forceinline__ float load_b(const float* __restrict__ src, const ConvGeom& g, int k, int j,
                                        int K, int Nn) {
  if (k >= K || j >= Nn) return 0.f;
  if (MODE == MODE_FWD) {
    return __ldg(src + (int64_t)j * K + k);  // w[o][c][r][t], k = (c,r,t)
  } else if (MODE == MODE_DGRAD) {
    int rt = k % (g.KH * g.KW), o = k / (g.KH * g.KW);
    return __ldg(src + ((int64_t)o * g.C + j) * (g.KH * g.KW) + rt);
  } else {
    int q = k % g.OW, p = (k / g.OW) % g.OH, n = k / (g.OW * g.OH);
    int t = j % g.KW, r = (j / g.KW) % g.KH, c = j / (g.KW * g.KH);
    int ih = p * g.stride - g.pad + r * g.dil, iw = q * g.stride - g.pad + t * g.dil;
    if ((unsigned)ih >= (unsigned)g.H || (unsigned)iw >= (unsigned)g.W) return 0.f;
    return __ldg(src + (((int64_t)n * g.C + c) * g.H + ih) * g.W + iw);
  }
}

template <int MODE>
__global__ void __launch_bounds__(CONV_THREADS)
conv_simt_kernel(const float* __restrict__ a_src, const float* __restrict__ b_src,
                 const float* __restrict__ bias, float* __restrict__ out, ConvGeom g, int M, int Nn,
                 int K, int k_per_split, float slope, int accumulate) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, j0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int ty = tid / 16, tx = tid % 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * CONV_THREADS;
      int ml, kl;
      if (MODE == MODE_WGRAD) {
        kl = e % BK;
        ml = e / BK;
      } else {
        ml = e % BM;
        kl = e / BM;
      }
      int kk = k0 + kl;
      As[kl][ml] = load_a<MODE>(a_src, g, m0 + ml, kk < kend ? kk : K, M, K);
      int kl2 = e % BK, jl = e / BK;
      int kk2 = k0 + kl2;
      Bs[kl2][jl] = load_b<MODE>(b_src, g, kk2 < kend ? kk2 : K, j0 + jl, K, Nn);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      int j = j0 + tx * 4 + jj;
      if (j >= Nn) continue;
      float v = acc[i][jj];
      if (MODE == MODE_FWD) {
        int q = m % g.OW, p = (m / g.OW) % g.OH, n = m / (g.OW * g.OH);
        int64_t o = (((int64_t)n * g.O + j) * g.OH + p) * g.OW + q;
        if (bias) v += __ldg(bias + j);
        if (slope != 1.f) v = v > 0.f ? v : v * slope;
        out[o] = accumulate ? out[o] + v : v;
      } else if (MODE == MODE_DGRAD) {
        int iw = m % g.W, ih = (m / g.W) % g.H, n = m / (g.W * g.H);
        int64_t o = (((int64_t)n * g.C + j) * g.H + ih) * g.W + iw;
        out[o] = accumulate ? out[o] + v : v;
      } else {
        int64_t o = (int64_t)m * Nn + j;
        if (gridDim.z > 1)
          atomicAdd(out + o, v);
        else
          out[o] = v;
      }
    }
  }
}

// db[o] = sum_{n,p,q} dy[n,o,p,q]; one CTA per output channel, fixed reduction order.
__global__ void __launch_bounds__(256)
bias_grad_kernel(const float* __restrict__ dy, float* __restrict__ db, int N, int O, int P) {
  __shared__ double part[8];
  const int o = blockIdx.x;
  double acc = 0.0;
  for (int n = 0; n < N; ++n) {
    const float* src = dy + ((int64_t)n * O + o) * P;
    for (int i = threadIdx.x; i < P; i += 256) acc += (double)src[i];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < 8 ? part[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) db[o] = (float)v;
  }
}

__global__ void __launch_bounds__(256)
lrelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ post, float* __restrict__ dx,
                 int64_t n, float slope) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
    dx[i] = dy[i] * (post[i] > 0.f ? 1.f : slope);
}

static int make_geom(ConvGeom& g, int N, int C, int H, int W, int O, int KH, int KW, int stride, int pad,
                     int dil) {
  ASN_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && O > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0 && dil > 0,
                "conv2d: bad geometry");
  g = {N, C, H, W, O, KH, KW, stride, pad, dil, 0, 0};
  g.OH = (H + 2 * pad - dil * (KH - 1) - 1) / stride + 1;
  g.OW = (W + 2 * pad - dil * (KW - 1) - 1) / stride + 1;
  ASN_CHECK_ARG(g.OH > 0 && g.OW > 0, "conv2d: empty output");
  ASN_CHECK_ARG((int64_t)N * H * W < (1LL << 31) && (int64_t)C * KH * KW < (1LL << 31), "conv2d: index overflow");
  return ASN_OK;
}

}  // namespace asn

using namespace asn;

extern "C" int asn_conv2d_fwd_f32(const float* x, const float* w, const float* bias, float* y, int N,
                                  int C, int H, int W, int O, int KH, int KW, int stride, int pad, int dil,
                                  float lrelu_slope, int accumulate, void* stream) {
  ASN_CHECK_ARG(x && w && y, "asn_conv2d_fwd_f32: null pointer");
  ConvGeom g;
  int rc = make_geom(g, N, C, H, W, O, KH, KW, stride, pad, dil);
  if (rc) return rc;
  int M = N * g.OH * g.OW, K = C * KH * KW;
  dim3 grid(cdiv(M, BM), cdiv(O, BN), 1);
  prof::Scope ps("conv_f32_fwd", 2.0 * M * O * K, 0, static_cast<cudaStream_t>(stream));
  conv_simt_kernel<MODE_FWD><<<grid, CONV_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      x, w, bias, y, g, M, O, K, K, lrelu_slope, accumulate);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

extern "C" int asn_conv2d_dgrad_f32(const float* dy, const float* w, float* dx, int N, int C, int H,
                                    int W, int O, int KH, int KW, int stride, int pad, int dil,
                                    int accumulate, void* stream) {
  ASN_CHECK_ARG(dy && w && dx, "asn_conv2d_dgrad_f32: null pointer");
  ConvGeom g;
  int rc = make_geom(g, N, C, H, W, O, KH, KW, stride, pad, dil);
  if (rc) return rc;
  int M = N * H * W, K = O * KH * KW;
  dim3 grid(cdiv(M, BM), cdiv(C, BN), 1);
  prof::Scope ps("conv_f32_dgrad", 2.0 * M * C * K, 0, static_cast<cudaStream_t>(stream));
  conv_simt_kernel<MODE_DGRAD><<<grid, CONV_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      dy, w, nullptr, dx, g, M, C, K, K, 1.f, accumulate);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

extern "C" int asn_conv2d_wgrad_f32(const float* x, const float* dy, float* dw, float* db, int N, int C,
                                    int H, int W, int O, int KH, int KW, int stride, int pad, int dil,
                                    void* stream) {
  ASN_CHECK_ARG(x && dy && dw, "asn_conv2d_wgrad_f32: null pointer");
  ConvGeom g;
  int rc = make_geom(g, N, C, H, W, O, KH, KW, stride, pad, dil);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int M = O, Nn = C * KH * KW, K = N * g.OH * g.OW;
  int tiles = cdiv(M, BM) * cdiv(Nn, BN);
  int split = (2 * sm_count() + tiles - 1) / tiles;
  int max_split = cdiv(K, 4 * BK);
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  int kps = (int)round_up(cdiv(K, split), BK);
  split = cdiv(K, kps);
  if (split > 1) ASN_CUDA(cudaMemsetAsync(dw, 0, (size_t)M * Nn * sizeof(float), st));
  dim3 grid(cdiv(M, BM), cdiv(Nn, BN), split);
  {
    prof::Scope ps("conv_f32_wgrad", 2.0 * M * Nn * K, 0, st);
    conv_simt_kernel<MODE_WGRAD><<<grid, CONV_THREADS, 0, st>>>(dy, x, nullptr, dw, g, M, Nn, K, kps, 1.f, 0);
    ASN_LAUNCH_CHECK();
  }
  if (db) {
    bias_grad_kernel<<<O, 256, 0, st>>>(dy, db, N, O, g.OH * g.OW);
    ASN_LAUNCH_CHECK();
  }
  return ASN_OK;
}

extern "C" int asn_lrelu_bwd_f32(const float* dy, const float* post, float* dx, int64_t n, float slope,
                                 void* stream) {
  ASN_CHECK_ARG(dy && post && dx && n >= 0, "asn_lrelu_bwd_f32: bad argument");
  if (n == 0) return ASN_OK;
  lrelu_bwd_kernel<<<wave_grid(n, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, post, dx, n, slope);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}
