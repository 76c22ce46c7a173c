// K2 / K2b / K9: bilinear resize with align_corners=True (model/deeplab_multi.py:188-189,
// evaluate_cityscapes.py:153,163,168-169).  HBM-bound: the forward is a pure 16-byte
// streaming write of the full-resolution tensor (the low-res source stays in L1/L2), the
// backward a pure streaming read.  Float op order follows ATen's upsample_bilinear2d so
// that the fused argmax resolves near-ties like the reference.
#include "common.cuh"

namespace asn {

constexpr int UP_THREADS = 256;

// one linear interpolation exactly as ATen's CPU kernel evaluates it: w0 * a + w1 * b with the second product rounded
// and the first fused into the addition (pinned bit for bit against the reference: tests/golden/eval2.npz)
__device__ __forceinline__ float lerp3(float w0, float w1, float a, float b) { return __fmaf_rn(w0, a, __fmul_rn(w1, b)); }

// one thread = VEC consecutive output columns; it keeps their x-interpolation (indices + weights) in
// registers and walks down the rows (nc, Y) assigned to its CTA row-slice, so the per-pixel work is
// 4 gathers (L1-resident low-res rows) + 6 FMAs and the kernel stays HBM-write-bound.
template <int VEC>
__global__ void upsample_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int NC, int h, int w, int H,
                                    int W, float sh, float sw) {
  const int wv = (W + VEC - 1) / VEC;
  const int xv = blockIdx.x * blockDim.x + threadIdx.x;
  if (xv >= wv) return;
  int i0[VEC], i1[VEC];
  float l0[VEC], l1[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const Lerp lx = lerp_at(min(xv * VEC + k, W - 1), sw, w);
    i0[k] = lx.i0; i1[k] = lx.i1; l0[k] = lx.l0; l1[k] = lx.l1;
  }
  const int base = i0[0], b1 = min(base + 1, w - 1), b2 = min(base + 2, w - 1);
  const bool narrow = (i0[VEC - 1] - base) <= 1 && (i1[VEC - 1] - base) <= 2;
  float wx[VEC][3];
#pragma unroll
  for (int k = 0; k < VEC; ++k)
#pragma unroll
    for (int m = 0; m < 3; ++m)
      wx[k][m] = (i0[k] - base == m ? l0[k] : 0.f) + (i1[k] - base == m ? l1[k] : 0.f);
  const int n_rows = NC * H;
  for (int row = blockIdx.y; row < n_rows; row += gridDim.y) {
    const int nc = row / H;
    const int Y = row - nc * H;
    const Lerp ly = lerp_at(Y, sh, h);
    const float* r0 = x + ((int64_t)nc * h + ly.i0) * w;
    const float* r1 = x + ((int64_t)nc * h + ly.i1) * w;
    float out[VEC];
    if (VEC == 4 && narrow) {
      // the 4 outputs touch at most source columns base, base+1, base+2: 6 gathers, then
      // out[k] = sum_m wx[k][m] * (ly.l0 * r0[m] + ly.l1 * r1[m]) with the x-weights folded once per thread
      const float v0 = ly.l0 * __ldg(r0 + base) + ly.l1 * __ldg(r1 + base);
      const float v1 = ly.l0 * __ldg(r0 + b1) + ly.l1 * __ldg(r1 + b1);
      const float v2 = ly.l0 * __ldg(r0 + b2) + ly.l1 * __ldg(r1 + b2);
#pragma unroll
      for (int k = 0; k < VEC; ++k) out[k] = wx[k][0] * v0 + wx[k][1] * v1 + wx[k][2] * v2;
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        out[k] = ly.l0 * (l0[k] * __ldg(r0 + i0[k]) + l1[k] * __ldg(r0 + i1[k])) +
                 ly.l1 * (l0[k] * __ldg(r1 + i0[k]) + l1[k] * __ldg(r1 + i1[k]));
    }
    float* dst = y + (int64_t)row * W + (int64_t)xv * VEC;
    if (VEC == 4) {
      st_stream(reinterpret_cast<float4*>(dst), make_float4(out[0], out[1], out[2], out[3]));
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        if (xv * VEC + k < W) dst[k] = out[k];
    }
  }
}

// backward pass 1: collapse the width.  A CTA stages UPB_ROWS full-res rows (nc, Y) in shared memory
// with coalesced 16-byte loads, then every thread sums, for one (row, j), the (<= ~2/scale) columns
// whose x0 or x1 is j.  T[nc, Y, j] (N*C*H*w floats) is the workspace.
constexpr int UPB_ROWS = 4;
__global__ void __launch_bounds__(UP_THREADS)
upsample_bwd_w_kernel(const float* __restrict__ dy, float* __restrict__ T, int n_rows, int w, int W, float sw,
                      int vec_ok) {
  extern __shared__ float rows_sh[];                          // [UPB_ROWS][W] | i0[W] | l1[W]
  int* tab_i0 = reinterpret_cast<int*>(rows_sh + UPB_ROWS * W);
  float* tab_l1 = rows_sh + (UPB_ROWS + 1) * W;
  for (int X = threadIdx.x; X < W; X += UP_THREADS) {          // same x-interpolation for every row
    const Lerp lx = lerp_at(X, sw, w);
    tab_i0[X] = lx.i0;
    tab_l1[X] = lx.l1;
  }
  const float inv = sw > 0.f ? 1.f / sw : 0.f;
  for (int row0 = blockIdx.x * UPB_ROWS; row0 < n_rows; row0 += gridDim.x * UPB_ROWS) {
    const int rows_here = min(UPB_ROWS, n_rows - row0);
    const float* src = dy + (int64_t)row0 * W;
    const int n_el = rows_here * W;  // the rows are contiguous in memory
    if (vec_ok) {
      for (int i = threadIdx.x; i < n_el / 4; i += UP_THREADS)
        reinterpret_cast<float4*>(rows_sh)[i] = ld_stream(reinterpret_cast<const float4*>(src) + i);
    } else {
      for (int i = threadIdx.x; i < n_el; i += UP_THREADS) rows_sh[i] = src[i];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < rows_here * w; idx += UP_THREADS) {
      const int rl = idx / w;
      const int j = idx - rl * w;
      const float* row = rows_sh + rl * W;
      // candidate columns: source coordinate in (j-1, j+1); widen by 2 and test against the table
      const int lo = sw > 0.f ? max(0, (int)floorf((float)(j - 1) * inv) - 2) : 0;
      const int hi = sw > 0.f ? min(W - 1, (int)ceilf((float)(j + 1) * inv) + 2) : W - 1;
      float acc = 0.f;
      for (int X = lo; X <= hi; ++X) {
        const int a0 = tab_i0[X];
        const float b1 = tab_l1[X];
        const float v = row[X];
        const int a1 = a0 + (a0 < w - 1 ? 1 : 0);
        if (a0 == j) acc += __fsub_rn(1.f, b1) * v;
        if (a1 == j) acc += b1 * v;
      }
      T[(int64_t)(row0 + rl) * w + j] = acc;
    }
    __syncthreads();
  }
}

// Fast variant of the width pass for moderate magnifications (<= UPB_MAXT source columns per output):
// thread j keeps the weights w(X, j) of its ~2/scale + 5 candidate columns in registers (they are the
// same for every row) and the staged rows are skewed (addr = X + X/32) so that lanes striding the row
// by the magnification hit distinct banks.  Per output: <= UPB_MAXT x (LDS + FMA).
constexpr int UPB_MAXT = 24;
__device__ __forceinline__ int skew(int X) { return X + (X >> 5); }
// BOUND = 256 (w <= 256, every training shape): registers to spare, so the skewed offsets of the candidate columns are
// computed once per thread and a tap is LDS + FFMA; BOUND = 1024: 64 registers, the offsets are recomputed per tap.
template <int BOUND>
__global__ void __launch_bounds__(BOUND)
upsample_bwd_w_fast_kernel(const float* __restrict__ dy, float* __restrict__ T, int n_rows, int w, int W,
                           float sw, int vec_ok) {
  constexpr bool OFFS = BOUND <= 256;
  extern __shared__ float rows_sh[];  // [UPB_ROWS][skew(W) + 1]
  const int pitch = skew(W) + 1;
  const int j = threadIdx.x;
  float wgt[UPB_MAXT];
  int off[OFFS ? UPB_MAXT : 1];   // skewed shared-memory offset of candidate column k
  int lo = 0;
  if (j < w) {
    const float inv = 1.f / sw;
    lo = max(0, (int)floorf((float)(j - 1) * inv) - 2);
    const int hi = min(W - 1, (int)ceilf((float)(j + 1) * inv) + 2);
#pragma unroll
    for (int k = 0; k < UPB_MAXT; ++k) {
      const int X = lo + k;
      float f = 0.f;
      if (X <= hi) {
        const Lerp lx = lerp_at(X, sw, w);
        f = (lx.i0 == j ? lx.l0 : 0.f) + (lx.i1 == j ? lx.l1 : 0.f);
      }
      wgt[k] = f;
      if (OFFS) off[k] = skew(min(X, W - 1));
    }
  }
  for (int row0 = blockIdx.x * UPB_ROWS; row0 < n_rows; row0 += gridDim.x * UPB_ROWS) {
    const int rows_here = min(UPB_ROWS, n_rows - row0);
    const float* src = dy + (int64_t)row0 * W;
    if (vec_ok) {
      // batches of 4 independent 16-byte loads per thread before the first shared-memory store
      const int n4 = rows_here * (W / 4);
      for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * blockDim.x) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * blockDim.x;
          if (i < n4) v[u] = ld_stream(reinterpret_cast<const float4*>(src) + i);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * blockDim.x;
          if (i < n4) {
            const int rl = (i * 4) / W, X = i * 4 - rl * W;
            float* d = rows_sh + rl * pitch;
            d[skew(X)] = v[u].x; d[skew(X + 1)] = v[u].y; d[skew(X + 2)] = v[u].z; d[skew(X + 3)] = v[u].w;
          }
        }
      }
    } else {
      for (int i = threadIdx.x; i < rows_here * W; i += blockDim.x) {
        const int rl = i / W, X = i - rl * W;
        rows_sh[rl * pitch + skew(X)] = src[i];
      }
    }
    __syncthreads();
    if (j < w) {
      for (int rl = 0; rl < rows_here; ++rl) {
        const float* row = rows_sh + rl * pitch;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < UPB_MAXT; ++k) acc = fmaf(wgt[k], row[OFFS ? off[k] : skew(min(lo + k, W - 1))], acc);
        T[(int64_t)(row0 + rl) * w + j] = acc;
      }
    }
    __syncthreads();
  }
}

// backward pass 2: collapse the height.  thread = (nc, i, j), j fastest -> coalesced reads of T.
__global__ void __launch_bounds__(UP_THREADS)
upsample_bwd_h_kernel(const float* __restrict__ T, float* __restrict__ dx, int NC, int h, int w, int H,
                      float sh) {
  const int64_t total = (int64_t)NC * h * w;
  const float inv = sh > 0.f ? 1.f / sh : 0.f;
  for (int64_t idx = (int64_t)blockIdx.x * UP_THREADS + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * UP_THREADS) {
    int j = (int)(idx % w);
    int i = (int)((idx / w) % h);
    int nc = (int)(idx / ((int64_t)w * h));
    int lo = sh > 0.f ? max(0, (int)floorf((float)(i - 1) * inv) - 2) : 0;
    int hi = sh > 0.f ? min(H - 1, (int)ceilf((float)(i + 1) * inv) + 2) : H - 1;
    float acc = 0.f;
    const float* base = T + (int64_t)nc * H * w + j;
    for (int Y = lo; Y <= hi; ++Y) {
      Lerp ly = lerp_at(Y, sh, h);
      float wgt = (ly.i0 == i ? ly.l0 : 0.f) + (ly.i1 == i ? ly.l1 : 0.f);
      if (wgt != 0.f) acc += wgt * __ldg(base + (int64_t)Y * w);
    }
    dx[idx] = acc;
  }
}

// K9: thread = 4 consecutive output pixels; interpolate all C channels, keep the first maximum.
__global__ void __launch_bounds__(UP_THREADS)
upsample_argmax_kernel(const float* __restrict__ x, uint8_t* __restrict__ pred, int N, int C, int h,
                       int w, int H, int W, float sh, float sw) {
  const int wv = (W + 3) / 4;
  const int64_t total = (int64_t)N * H * wv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int xv = (int)(i % wv);
    int64_t row = i / wv;
    int Y = (int)(row % H);
    int n = (int)(row / H);
    Lerp ly = lerp_at(Y, sh, h);
    Lerp lx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) lx[k] = lerp_at(min(xv * 4 + k, W - 1), sw, w);
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int arg[4] = {0, 0, 0, 0};
    for (int c = 0; c < C; ++c) {
      const float* r0 = x + (((int64_t)n * C + c) * h + ly.i0) * w;
      const float* r1 = x + (((int64_t)n * C + c) * h + ly.i1) * w;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // ATen's CPU kernel (UpSampleKernel.cpp `interpolate`, what the reference's evaluate script runs) computes
        // each lerp as fma(w0, a, w1 * b): spelled out, so the values -- and the ties -- are the reference's bit for bit
        const float v = lerp3(ly.l0, ly.l1, lerp3(lx[k].l0, lx[k].l1, __ldg(r0 + lx[k].i0), __ldg(r0 + lx[k].i1)),
                              lerp3(lx[k].l0, lx[k].l1, __ldg(r1 + lx[k].i0), __ldg(r1 + lx[k].i1)));
        // strict '>' keeps the first maximum (np.argmax); NaN handling: numpy treats the
        // first NaN as the maximum
        if (v > best[k] || (v != v && best[k] == best[k])) {
          best[k] = v;
          arg[k] = c;
        }
      }
    }
    uint8_t* dst = pred + row * W + (int64_t)xv * 4;
    if ((W & 3) == 0) {
      *reinterpret_cast<uint32_t*>(dst) =
          (uint32_t)arg[0] | ((uint32_t)arg[1] << 8) | ((uint32_t)arg[2] << 16) | ((uint32_t)arg[3] << 24);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (xv * 4 + k < W) dst[k] = (uint8_t)arg[k];
    }
  }
}

// K9, two stages: the fork upsamples the heads' logits to the INPUT size inside ResNetMulti.forward
// (model/deeplab_multi.py:188-189) and evaluate_cityscapes.py:153,163 resizes that again to the label size before the
// argmax (:168-169).  Two align_corners bilinear resizes compose into one only when (out-1) % (in-1) == 0, which the
// Cityscapes shapes (64x128 -> 512x1024 -> 1024x2048) do not satisfy, so both stages are computed -- but the 39.8 MB
// intermediate never leaves the SM: a CTA owns a TH2 x TW2 output tile, interpolates the (few) intermediate pixels the
// tile touches for all C channels into shared memory (stage 1, from the L1/L2-resident low-res logits), then every
// thread interpolates its four outputs from shared memory and keeps the first maximum (stage 2).  Per frame 0.6 MB are
// read and 2 MB written.  Both stages use ATen's op order (lerp3), so the result equals the reference's bit for bit.
constexpr int TH2 = 8, TW2 = 128;   // output tile: 256 threads x 4 pixels
struct HistArgs {           // optional fused confusion matrix (compute_iou.py:15-17,55-57); label == nullptr: off
  const void* label;        // [N][H][W], ASN_LABEL_U8 / I32 / I64
  int label_dtype;
  const uint8_t* lut;       // nullable 256-entry label_mapping table (compute_iou.py:24-28)
  int n_cls;
  unsigned long long* hist;      // [n_cls][n_cls], accumulated
  unsigned long long* overflow;  // flat index >= n_cls^2 under a valid label (np.bincount would grow)
};

__device__ __forceinline__ int hist_label(const HistArgs& ha, int64_t i) {
  // -1 = not counted: the reference's mask (a >= 0) & (a < n) after the optional mapping of ids in [0, 256)
  unsigned long long v;
  if (ha.label_dtype == ASN_LABEL_U8) v = static_cast<const uint8_t*>(ha.label)[i];
  else if (ha.label_dtype == ASN_LABEL_I32) v = (unsigned long long)(long long)static_cast<const int32_t*>(ha.label)[i];
  else v = (unsigned long long)static_cast<const int64_t*>(ha.label)[i];
  if (ha.lut && v < 256ull) v = ha.lut[v];
  return v < (unsigned long long)ha.n_cls ? (int)v : -1;
}

__global__ void __launch_bounds__(256)
upsample2_argmax_kernel(const float* __restrict__ x, uint8_t* __restrict__ pred, int N, int C, int h, int w, int Hm,
                        int Wm, int H, int W, float s1h, float s1w, float s2h, float s2w, int MR, int MC,
                        const HistArgs ha) {
  extern __shared__ float mid[];  // [C][MR][MC] (+ n_cls^2 uint32 bins when the confusion matrix is fused)
  uint32_t* bins = reinterpret_cast<uint32_t*>(mid + C * MR * MC);
  const int nbins = ha.n_cls * ha.n_cls;
  if (ha.label)
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) bins[i] = 0;
  const int tiles_w = (W + TW2 - 1) / TW2, tiles_h = (H + TH2 - 1) / TH2;
  const int tx = blockIdx.x % tiles_w, ty = (blockIdx.x / tiles_w) % tiles_h, n = blockIdx.x / (tiles_w * tiles_h);
  const int Y0 = ty * TH2, X0 = tx * TW2;
  const int my0 = lerp_at(Y0, s2h, Hm).i0, mx0 = lerp_at(X0, s2w, Wm).i0;   // first intermediate row / column of the tile
  // ---- stage 1: intermediate pixels (my0 + r, mx0 + c) for all channels ----
  const float* xn = x + (int64_t)n * C * h * w;
  for (int i = threadIdx.x; i < MR * MC; i += blockDim.x) {
    const int r = i / MC, c = i - r * MC;
    const int my = min(my0 + r, Hm - 1), mx = min(mx0 + c, Wm - 1);
    const Lerp ly = lerp_at(my, s1h, h), lx = lerp_at(mx, s1w, w);
    const float* p00 = xn + (int64_t)ly.i0 * w + lx.i0;
    const int dx = lx.i1 - lx.i0, dy = (ly.i1 - ly.i0) * w;
    for (int ch = 0; ch < C; ++ch) {
      const float* p = p00 + (int64_t)ch * h * w;
      mid[(ch * MR + r) * MC + c] = lerp3(ly.l0, ly.l1, lerp3(lx.l0, lx.l1, __ldg(p), __ldg(p + dx)),
                                          lerp3(lx.l0, lx.l1, __ldg(p + dy), __ldg(p + dy + dx)));
    }
  }
  __syncthreads();
  // ---- stage 2: thread = 4 consecutive output pixels of one row ----
  const int Y = Y0 + (threadIdx.x >> 5), Xb = X0 + (threadIdx.x & 31) * 4;
  const bool active = Y < H && Xb < W;
  if (active) {
  const Lerp ly = lerp_at(min(Y, H - 1), s2h, Hm);
  const int r0 = (ly.i0 - my0) * MC, r1 = (ly.i1 - my0) * MC;
  int c0[4], c1[4];
  float l0[4], l1[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const Lerp lx = lerp_at(min(Xb + k, W - 1), s2w, Wm);
    c0[k] = lx.i0 - mx0; c1[k] = lx.i1 - mx0; l0[k] = lx.l0; l1[k] = lx.l1;
  }
  float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int arg[4] = {0, 0, 0, 0};
  for (int ch = 0; ch < C; ++ch) {
    const float* m0 = mid + ch * MR * MC + r0;
    const float* m1 = mid + ch * MR * MC + r1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float v = lerp3(ly.l0, ly.l1, lerp3(l0[k], l1[k], m0[c0[k]], m0[c1[k]]), lerp3(l0[k], l1[k], m1[c0[k]], m1[c1[k]]));
      if (v > best[k] || (v != v && best[k] == best[k])) {   // first maximum; the first NaN wins (numpy)
        best[k] = v;
        arg[k] = ch;
      }
    }
  }
  const int64_t px = ((int64_t)n * H + Y) * W + Xb;
  if (pred) {
    uint8_t* dst = pred + px;
    if ((W & 3) == 0) {
      *reinterpret_cast<uint32_t*>(dst) =
          (uint32_t)arg[0] | ((uint32_t)arg[1] << 8) | ((uint32_t)arg[2] << 16) | ((uint32_t)arg[3] << 24);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (Xb + k < W) dst[k] = (uint8_t)arg[k];
    }
  }
  if (ha.label) {   // hist[a][b] += 1 for valid labels; equal neighbours are merged before the shared-memory atomic
    int cur = -1;
    uint32_t cnt = 0, ovf = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (Xb + k >= W) break;
      const int a = hist_label(ha, px + k);
      const int idx = a >= 0 ? a * ha.n_cls + arg[k] : -1;
      if (idx == cur) { ++cnt; continue; }
      if (cur >= 0) { if (cur < nbins) atomicAdd(&bins[cur], cnt); else ovf += cnt; }
      cur = idx; cnt = 1;
    }
    if (cur >= 0) { if (cur < nbins) atomicAdd(&bins[cur], cnt); else ovf += cnt; }
    if (ovf) atomicAdd(ha.overflow, (unsigned long long)ovf);
  }
  }  // active
  if (ha.label) {
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += blockDim.x)
      if (bins[i]) atomicAdd(&ha.hist[i], (unsigned long long)bins[i]);
  }
}

}  // namespace asn

using namespace asn;

extern "C" int asn_upsample_bilinear_fwd(const float* x, int N, int C, int h, int w, float* y, int H,
                                         int W, void* stream) {
  ASN_CHECK_ARG(x && y, "asn_upsample_bilinear_fwd: null pointer");
  ASN_CHECK_ARG(N > 0 && C > 0 && h > 0 && w > 0 && H > 0 && W > 0, "asn_upsample_bilinear_fwd: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float sh = lerp_scale(h, H), sw = lerp_scale(w, W);
  bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
  prof::Scope ps("upsample_fwd", 0, 4.0 * N * C * ((double)H * W + (double)h * w), st);
  const int wv = vec ? W / 4 : W;
  // block = one row of column vectors (rounded up to warps, capped at 512 threads)
  int threads = (int)round_up(wv < 512 ? wv : 512, 32);
  if (wv > 512) threads = (int)round_up(cdiv(wv, cdiv(wv, 512)), 32);
  const int gx = cdiv(wv, threads);
  const int rows = N * C * H;
  int gy = (8 * sm_count() * 256) / (gx * threads);  // ~8 x 256 resident threads per SM
  if (gy > rows) gy = rows;
  if (gy < 1) gy = 1;
  if (vec) {
    upsample_fwd_kernel<4><<<dim3(gx, gy), threads, 0, st>>>(x, y, N * C, h, w, H, W, sh, sw);
  } else {
    upsample_fwd_kernel<1><<<dim3(gx, gy), threads, 0, st>>>(x, y, N * C, h, w, H, W, sh, sw);
  }
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

extern "C" size_t asn_upsample_bwd_workspace_bytes(int N, int C, int H, int W, int h, int w) {
  (void)W; (void)h;
  return (size_t)N * C * H * w * sizeof(float);
}

extern "C" int asn_upsample_bilinear_bwd(const float* dy, int N, int C, int H, int W, float* dx, int h,
                                         int w, void* workspace, size_t workspace_bytes, void* stream) {
  ASN_CHECK_ARG(dy && dx && workspace, "asn_upsample_bilinear_bwd: null pointer");
  ASN_CHECK_ARG(N > 0 && C > 0 && h > 0 && w > 0 && H > 0 && W > 0, "asn_upsample_bilinear_bwd: bad shape");
  if (workspace_bytes < asn_upsample_bwd_workspace_bytes(N, C, H, W, h, w)) {
    set_error("asn_upsample_bilinear_bwd: workspace too small");
    return ASN_EWORKSPACE;
  }
  ASN_CHECK_ARG((size_t)W * 4 * (UPB_ROWS + 2) <= 200 * 1024, "asn_upsample_bilinear_bwd: rows of %d floats exceed shared memory", W);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float sh = lerp_scale(h, H), sw = lerp_scale(w, W);
  float* T = static_cast<float*>(workspace);
  int vec_ok = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(dy) & 15) == 0);
  size_t smem = (size_t)W * 4 * (UPB_ROWS + 2);
  const int n_rows = N * C * H;
  const int groups_w = cdiv(n_rows, UPB_ROWS);
  const int grid_w = groups_w < 8 * sm_count() ? groups_w : 8 * sm_count();
  if (smem > 48 * 1024)
    ASN_CUDA(cudaFuncSetAttribute(upsample_bwd_w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    prof::Scope ps("upsample_bwd_w", 0, 4.0 * N * C * ((double)H * W + (double)H * w), st);
    const bool fast = sw > 0.f && w <= 1024 && (int)ceilf(2.f / sw) + 7 <= UPB_MAXT;
    if (fast) {
      const size_t smem_f = (size_t)UPB_ROWS * (W + (W >> 5) + 1) * 4;
      if (w <= 256) {
        if (smem_f > 48 * 1024)
          ASN_CUDA(cudaFuncSetAttribute(upsample_bwd_w_fast_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
        upsample_bwd_w_fast_kernel<256><<<grid_w, (unsigned)round_up(w, 32), smem_f, st>>>(dy, T, n_rows, w, W, sw, vec_ok);
      } else {
        if (smem_f > 48 * 1024)
          ASN_CUDA(cudaFuncSetAttribute(upsample_bwd_w_fast_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
        upsample_bwd_w_fast_kernel<1024><<<grid_w, (unsigned)round_up(w, 32), smem_f, st>>>(dy, T, n_rows, w, W, sw, vec_ok);
      }
    } else {
      upsample_bwd_w_kernel<<<grid_w, UP_THREADS, smem, st>>>(dy, T, n_rows, w, W, sw, vec_ok);
    }
    ASN_LAUNCH_CHECK();
  }
  int64_t items = (int64_t)N * C * h * w;
  {
    prof::Scope ps("upsample_bwd_h", 0, 4.0 * N * C * ((double)H * w + (double)h * w), st);
    upsample_bwd_h_kernel<<<full_grid(items, UP_THREADS), UP_THREADS, 0, st>>>(T, dx, N * C, h, w, H, sh);
    ASN_LAUNCH_CHECK();
  }
  return ASN_OK;
}

extern "C" int asn_upsample_argmax_u8(const float* x, int N, int C, int h, int w, uint8_t* pred, int H,
                                      int W, void* stream) {
  ASN_CHECK_ARG(x && pred, "asn_upsample_argmax_u8: null pointer");
  ASN_CHECK_ARG(N > 0 && C > 0 && C <= 256 && h > 0 && w > 0 && H > 0 && W > 0, "asn_upsample_argmax_u8: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float sh = lerp_scale(h, H), sw = lerp_scale(w, W);
  int64_t items = (int64_t)N * H * ((W + 3) / 4);
  prof::Scope ps("upsample_argmax", 0, 4.0 * N * C * h * w + (double)N * H * W, st);
  upsample_argmax_kernel<<<full_grid(items, 128), 128, 0, st>>>(x, pred, N, C, h, w, H, W, sh, sw);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

// rows / columns of the intermediate grid a TH2 x TW2 output tile can touch (host side of upsample2_argmax_kernel)
static int mid_extent(int tile, int n_mid, int n_out) {
  const float s = lerp_scale(n_mid, n_out);
  int e = (int)ceilf(s * (float)(tile - 1)) + 3;
  return e < n_mid + 1 ? e : n_mid + 1;
}

static int upsample2_impl(const float* x, int N, int C, int h, int w, int Hm, int Wm, uint8_t* pred, int H, int W,
                          const HistArgs& ha, cudaStream_t st) {
  const int MR = mid_extent(TH2, Hm, H), MC = mid_extent(TW2, Wm, W);
  const size_t smem = (size_t)C * MR * MC * sizeof(float) + (ha.label ? (size_t)ha.n_cls * ha.n_cls * 4 : 0);
  if (smem > 160 * 1024) {
    set_error("asn_upsample2_argmax: a %dx%d output tile needs %zu bytes of shared memory (strong minification in the "
              "second stage, or too many classes for the fused confusion matrix); use the unfused entry points", TH2, TW2,
              smem);
    return ASN_EUNSUPPORTED;
  }
  static PerDevice smem_set;
  if (smem > 48 * 1024 && smem_set.get() < (int)smem) {
    ASN_CUDA(cudaFuncSetAttribute(upsample2_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set.set((int)smem);
  }
  const long long tiles = (long long)N * cdiv(H, TH2) * cdiv(W, TW2);
  const double label_bytes = !ha.label ? 0.0 : (ha.label_dtype == ASN_LABEL_U8 ? 1.0 : ha.label_dtype == ASN_LABEL_I32 ? 4.0 : 8.0);
  prof::Scope ps(ha.label ? "upsample2_argmax_hist" : "upsample2_argmax", 0,
                 4.0 * N * C * h * w + (double)N * H * W * ((pred ? 1.0 : 0.0) + label_bytes), st);
  upsample2_argmax_kernel<<<(unsigned)tiles, 256, smem, st>>>(x, pred, N, C, h, w, Hm, Wm, H, W, lerp_scale(h, Hm),
                                                              lerp_scale(w, Wm), lerp_scale(Hm, H), lerp_scale(Wm, W), MR,
                                                              MC, ha);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

extern "C" int asn_upsample2_argmax_u8(const float* x, int N, int C, int h, int w, int Hm, int Wm, uint8_t* pred, int H,
                                       int W, void* stream) {
  ASN_CHECK_ARG(x && pred, "asn_upsample2_argmax_u8: null pointer");
  ASN_CHECK_ARG(N > 0 && C > 0 && C <= 256 && h > 0 && w > 0 && Hm > 0 && Wm > 0 && H > 0 && W > 0,
                "asn_upsample2_argmax_u8: bad shape");
  HistArgs ha{};
  return upsample2_impl(x, N, C, h, w, Hm, Wm, pred, H, W, ha, static_cast<cudaStream_t>(stream));
}

extern "C" int asn_upsample2_argmax_hist(const float* x, int N, int C, int h, int w, int Hm, int Wm, uint8_t* pred, int H,
                                         int W, const void* label, int label_dtype, const uint8_t* lut256, int n_cls,
                                         int64_t* hist, int64_t* overflow, void* stream) {
  ASN_CHECK_ARG(x && label && hist && overflow, "asn_upsample2_argmax_hist: null pointer");
  ASN_CHECK_ARG(N > 0 && C > 0 && C <= 256 && h > 0 && w > 0 && Hm > 0 && Wm > 0 && H > 0 && W > 0,
                "asn_upsample2_argmax_hist: bad shape");
  ASN_CHECK_ARG(n_cls >= 1 && n_cls <= 64, "asn_upsample2_argmax_hist: n_cls=%d outside [1,64]", n_cls);
  ASN_CHECK_ARG(label_dtype == ASN_LABEL_U8 || label_dtype == ASN_LABEL_I32 || label_dtype == ASN_LABEL_I64,
                "asn_upsample2_argmax_hist: unknown label_dtype %d", label_dtype);
  HistArgs ha{label, label_dtype, lut256, n_cls, reinterpret_cast<unsigned long long*>(hist),
              reinterpret_cast<unsigned long long*>(overflow)};
  return upsample2_impl(x, N, C, h, w, Hm, Wm, pred, H, W, ha, static_cast<cudaStream_t>(stream));
}
