// Tier-B "lazy upsample" kernels: the consumers of interp(logits) read the low-res logits and
// interpolate on the fly; the full-resolution fp32 logits / probabilities of the reference are never written.
//
//   asn_upsample_ce_fwd_bwd : nn.Upsample(bilinear, align_corners=True) -> CrossEntropyLoss(ignore_index) forward AND the
//                             gradient w.r.t. the low-res logits in one pass over the labels
//                             (model/deeplab_multi.py:188-189 + train_gta2cityscapes_multi.py:599-600 + autograd)
//   lazy::pack_input        : interp -> F.softmax -> discriminator input (train...:617-618, 645-646, 665-666)
//   lazy::unpack_dx         : the backward of that chain down to the low-res logits
//
// One skeleton ("strip kernel"): a CTA owns LZ_COLS full-res columns x SH (<= 8) full-res rows of one image; a thread
// owns one column.  The strip touches at most three low-res rows, whose x-interpolated values the thread keeps in
// registers (3 x C floats); walking down the rows it forms the C logits with one y-lerp each (the same two-stage lerp
// ATen performs), the channel softmax, and a per-pixel gradient g[c]:
//     CE   : g = w_y * (p - onehot(y)) for valid pixels     (+ loss / weight / count statistics)
//     DBWD : g = p * (dp - sum_c p*dp),  dp = dA0 (bf16 NHWC)
// g is folded immediately into the three low-res rows (y-transpose, registers), then through shared memory into the
// low-res columns of the block (x-transpose, gather form: no atomics).  Each CTA writes its [3][C][JB] partial block;
// a second, tiny kernel adds, for every low-res node, the <= 2 x 3 partials that can touch it in a FIXED order, so the
// result is deterministic.  Compulsory HBM traffic: labels (8 B/px) or dA0 (64 B/px) + a few MB, instead of 4-6
// passes over C*4 B/px.
#include "lazy_up.cuh"

#include "../../include/asn_b200.h"
#include "ce_common.cuh"

namespace asn {
namespace lazy {

constexpr int LZ_COLS = 128;   // full-res columns per CTA = threads per CTA
constexpr int LZ_MAX_SH = 9;   // full-res rows per strip (upper bound)
constexpr int LZ_R = 3;        // low-res rows a strip may touch

enum { F_CE = 0, F_PACK = 1, F_DBWD = 2 };

struct Geom {
  int N, h, w, H, W;
  float sy, sx;
  int SH, strips, xblocks, JB;
};

// `slots` = CTAs of the strip kernel the device keeps resident (0: unknown).  A strip costs about SH + 2.3 row-times
// (interpolation of the low-res rows and the x-transpose are per strip), the grid runs in ceil(CTAs / slots) rounds:
// SH is the admissible value with the cheapest total (e.g. 720 x 1280 on 444 slots: 9 rows = 2 rounds, 8 rows = 3).
static Geom make_geom(int N, int h, int w, int H, int W, int slots = 0) {
  Geom g;
  g.N = N; g.h = h; g.w = w; g.H = H; g.W = W;
  g.sy = lerp_scale(h, H);
  g.sx = lerp_scale(w, W);
  // rows Y0 .. Y0+SH-1 must not span more than two values of floor(sy*Y): sy * (SH - 1) <= 1
  int sh = g.sy > 0.f ? 1 + (int)floorf(1.f / g.sy) : LZ_MAX_SH;
  sh = sh > LZ_MAX_SH ? LZ_MAX_SH : (sh < 1 ? 1 : sh);
  while (sh > 1 && g.sy * (float)(sh - 1) > 1.f) --sh;
  g.xblocks = cdiv(W, LZ_COLS);
  g.SH = sh;
  if (slots > 0) {
    double best = 0.0;
    for (int cand = sh; cand >= (sh > 3 ? 3 : 1); --cand) {
      const long long ctas = (long long)N * cdiv(H, cand) * g.xblocks;
      const double cost = (double)cdiv(ctas, (long long)slots) * (cand + 2.3);
      if (cand == sh || cost < best) {
        best = cost;
        g.SH = cand;
      }
    }
  }
  g.strips = cdiv(H, g.SH);
  g.JB = (int)ceilf(g.sx * (float)(LZ_COLS - 1)) + 3;  // low-res columns a block of LZ_COLS columns may touch
  return g;
}

bool supported(int C, int h, int w, int H, int W) {
  return C == 19 && h >= 1 && w >= 1 && H >= h && W >= w;
}

// sized for the smallest strip height make_geom may choose (the choice depends on the device)
size_t partial_bytes(int N, int C, int h, int w, int H, int W) {
  const Geom g = make_geom(N, h, w, H, W);
  const int sh_min = g.SH > 3 ? 3 : 1;
  return (size_t)N * cdiv(H, sh_min) * g.xblocks * LZ_R * C * g.JB * sizeof(float);
}

struct Args {
  Geom g;
  const float* z;            // low-res logits [N][C][h][w]
  // CE
  const long long* y;        // labels [N][H][W]
  int ignore, mask_negative;
  const float* cw;           // nullable class weights
  CeStats* stats;
  // PACK / DBWD
  __nv_bfloat16* a0;         // PACK: out [N][H][W0p][32];  DBWD: in (dA0)
  int W0p;
  // CE / DBWD
  float* partial;            // [N][strips][xblocks][3*C][JB]
};

__device__ __forceinline__ void store_px32_bf16(__nv_bfloat16* dst, const float* v) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 hh = __floats2bfloat162_rn(v[q * 8 + 2 * j], v[q * 8 + 2 * j + 1]);
      pk[j] = *reinterpret_cast<uint32_t*>(&hh);
    }
    d[q] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

template <int C, int FUNC>
__global__ void __launch_bounds__(LZ_COLS, FUNC == F_PACK ? 4 : 3)
lazy_strip_kernel(const Args a) {
  static_assert(C <= 32, "a pixel is 32 bf16 channels in the discriminator input");
  // S[k*C + c][x]: the (<= 3) low-res rows of this strip interpolated to full-res column x; every thread fills and reads
  // only its own column (no barrier), and re-uses it at the end for its folded gradient (x-transpose input)
  __shared__ float S[LZ_R * C][LZ_COLS + 1];
  const Geom& g = a.g;
  const int xb = blockIdx.x, strip = blockIdx.y, n = blockIdx.z;
  const int tid = threadIdx.x;
  const int X0 = xb * LZ_COLS, Y0 = strip * g.SH;
  const int X = X0 + tid;
  const bool col_ok = X < g.W;
  const Lerp lx = lerp_at(min(X, g.W - 1), g.sx, g.w);
  const int ibase = lerp_at(Y0, g.sy, g.h).i0;
  const int rows = min(g.SH, g.H - Y0);

#pragma unroll
  for (int k = 0; k < LZ_R; ++k) {
    const int i = min(ibase + k, g.h - 1);
    const float* zr = a.z + ((int64_t)n * C * g.h + i) * g.w;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float* zc = zr + (int64_t)c * g.h * g.w;
      S[k * C + c][tid] = lx.l0 * __ldg(zc + lx.i0) + lx.l1 * __ldg(zc + lx.i1);
    }
  }

  float acc[LZ_R][C];
  if (FUNC != F_PACK) {
#pragma unroll
    for (int k = 0; k < LZ_R; ++k)
#pragma unroll
      for (int c = 0; c < C; ++c) acc[k][c] = 0.f;
  }
  double loss = 0.0, wsum = 0.0;
  long long nvalid = 0, nbad = 0;

  for (int r = 0; r < rows; ++r) {
    const int Y = Y0 + r;
    const Lerp ly = lerp_at(Y, g.sy, g.h);
    const int ka = ly.i0 - ibase, kb = ly.i1 - ibase;  // 0..2, uniform over the CTA
    const float* sa = &S[ka * C][tid];
    const float* sb = &S[kb * C][tid];
    float v[C];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      v[c] = ly.l0 * sa[c * (LZ_COLS + 1)] + ly.l1 * sb[c * (LZ_COLS + 1)];
      m = fmaxf(m, v[c]);
    }
    // exp(v - m) as one FFMA + ex2 (2 ulp; the kernels' tolerance is 1e-5 norm-wise)
    constexpr float LOG2E = 1.4426950408889634f;
    const float mb = m * LOG2E;
    float s = 0.f;
    if (FUNC == F_CE) {
      const long long lab = col_ok ? __ldg(a.y + ((int64_t)n * g.H + Y) * g.W + X) : (long long)a.ignore;
      const int cls = classify_label(lab, C, a.ignore, a.mask_negative);
      const int yi = cls == 1 ? (int)lab : 0;
      const float wgt = cls == 1 ? (a.cw ? __ldg(a.cw + yi) : 1.f) : 0.f;
      // the logit of the labelled class, re-interpolated with a dynamic class index (shared memory, not registers)
      const float zy = ly.l0 * sa[yi * (LZ_COLS + 1)] + ly.l1 * sb[yi * (LZ_COLS + 1)];
#pragma unroll
      for (int c = 0; c < C; ++c) { v[c] = exp2f(fmaf(v[c], LOG2E, -mb)); s += v[c]; }
      if (cls == 1) {
        loss += (double)(wgt * ((m + logf(s)) - zy));
        wsum += (double)wgt;
        ++nvalid;
      } else if (cls < 0 && col_ok) {
        ++nbad;
      }
      const float winv = wgt / s;  // 0 for ignored pixels: their gradient vanishes without a branch
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = fmaf(v[c], winv, c == yi ? -wgt : 0.f);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) { v[c] = exp2f(fmaf(v[c], LOG2E, -mb)); s += v[c]; }
      const float inv = 1.f / s;
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] *= inv;  // probabilities
      if (FUNC == F_PACK) {
        __nv_bfloat16* dst_row = a.a0 + ((int64_t)n * g.H + Y) * (int64_t)a.W0p * 32;
        float px[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) px[c] = c < C ? v[c] : 0.f;
        if (col_ok) store_px32_bf16(dst_row + (int64_t)(X + 1) * 32, px);
#pragma unroll
        for (int c = 0; c < 32; ++c) px[c] = 0.f;
        if (X == 0) store_px32_bf16(dst_row, px);  // the zero padding columns of this row
        if (X == g.W - 1)
          for (int wp = g.W + 1; wp < a.W0p; ++wp) store_px32_bf16(dst_row + (int64_t)wp * 32, px);
      } else {
        // g = p * (dp - sum_c p * dp)
        float dp[32];
        const uint4* src = reinterpret_cast<const uint4*>(
            a.a0 + (((int64_t)n * g.H + Y) * (int64_t)a.W0p + min(X, g.W - 1) + 1) * 32);
#pragma unroll
        for (int q = 0; q < (C + 7) / 8; ++q) {
          const uint4 u = __ldg(src + q);
          const uint32_t pk[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(&pk[j]);
            dp[q * 8 + 2 * j] = __low2float(hh);
            dp[q * 8 + 2 * j + 1] = __high2float(hh);
          }
        }
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) dot += v[c] * dp[c];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = col_ok ? v[c] * (dp[c] - dot) : 0.f;
      }
    }
    if (FUNC != F_PACK) {
      // y-transpose: this row's gradient goes to low-res rows ka / kb with its two lerp weights
      const float w0 = (ka == 0 ? ly.l0 : 0.f) + (kb == 0 ? ly.l1 : 0.f);
      const float w1 = (ka == 1 ? ly.l0 : 0.f) + (kb == 1 ? ly.l1 : 0.f);
      const float w2 = (ka == 2 ? ly.l0 : 0.f) + (kb == 2 ? ly.l1 : 0.f);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        acc[0][c] += w0 * v[c];
        acc[1][c] += w1 * v[c];
        acc[2][c] += w2 * v[c];
      }
    }
  }
  if constexpr (FUNC != F_PACK) {
    // x-transpose through shared memory: out[jl][k*C + c] = sum_x wx(j0 + jl, x) * acc_x[k][c].  floor(sx * x) is
    // monotone in x, so the columns whose left neighbour is low-res column j form one contiguous range
    // [xs[jl], xs[jl + 1]); node j collects l0 from that range and l1 from the range of j - 1.
    __shared__ int s_li[LZ_COLS];
    __shared__ int xs[LZ_COLS + 8];
    __shared__ float s_l0[LZ_COLS], s_l1[LZ_COLS];
    const int j0 = lerp_at(min(X0, g.W - 1), g.sx, g.w).i0;
    const int li = lx.i0 - j0;
#pragma unroll
    for (int k = 0; k < LZ_R; ++k)
#pragma unroll
      for (int c = 0; c < C; ++c) S[k * C + c][tid] = acc[k][c];
    s_li[tid] = li;
    const bool clamped = lx.i1 == lx.i0;  // last low-res column: both weights go to the same node
    s_l0[tid] = col_ok ? (clamped ? lx.l0 + lx.l1 : lx.l0) : 0.f;
    s_l1[tid] = col_ok && !clamped ? lx.l1 : 0.f;
    __syncthreads();
    {
      const int prev = tid > 0 ? s_li[tid - 1] : -1;
      for (int jl = prev + 1; jl <= li; ++jl) xs[jl] = tid;
      if (tid == LZ_COLS - 1)
        for (int jl = li + 1; jl <= g.JB; ++jl) xs[jl] = LZ_COLS;
    }
    __syncthreads();
    // partial block layout [k*C + c][jl] (jl fastest: what the reduce kernel's neighbouring threads read); staged in
    // shared memory when it is small (the 8x upsampling of the training shapes: JB = 19) so that it leaves coalesced
    constexpr int STAGE_JB = 24;
    __shared__ float outs[LZ_R * C * STAGE_JB];
    float* part = a.partial + (((int64_t)n * g.strips + strip) * g.xblocks + xb) * (int64_t)(LZ_R * C * g.JB);
    const bool staged = g.JB <= STAGE_JB;
    for (int item = tid; item < LZ_R * C * g.JB; item += LZ_COLS) {
      const int ck = item % (LZ_R * C), jl = item / (LZ_R * C);  // a warp walks one x-range over 32 rows of S
      float sum = 0.f;
      for (int x = xs[jl]; x < xs[jl + 1]; ++x) sum += s_l0[x] * S[ck][x];
      if (jl > 0)
        for (int x = xs[jl - 1]; x < xs[jl]; ++x) sum += s_l1[x] * S[ck][x];
      if (staged) outs[ck * g.JB + jl] = sum; else part[ck * g.JB + jl] = sum;
    }
    if (staged) {
      __syncthreads();
      for (int e = tid; e < LZ_R * C * g.JB; e += LZ_COLS) part[e] = outs[e];
    }

    if (FUNC == F_CE) {
      // block statistics (same accumulation as pointwise.cu)
      __shared__ double sa[LZ_COLS / 32], sb[LZ_COLS / 32];
      __shared__ long long sc[LZ_COLS / 32], sd[LZ_COLS / 32];
      loss = warp_sum(loss); wsum = warp_sum(wsum); nvalid = warp_sum(nvalid); nbad = warp_sum(nbad);
      const int wid = tid >> 5, lane = tid & 31;
      if (lane == 0) { sa[wid] = loss; sb[wid] = wsum; sc[wid] = nvalid; sd[wid] = nbad; }
      __syncthreads();
      if (tid == 0) {
        double A = 0.0, B = 0.0;
        long long Cc = 0, D = 0;
        for (int i = 0; i < LZ_COLS / 32; ++i) { A += sa[i]; B += sb[i]; Cc += sc[i]; D += sd[i]; }
        if (A != 0.0 || A != A) atomicAdd(&a.stats->loss_sum, A);
        if (B != 0.0) atomicAdd(&a.stats->weight_sum, B);
        if (Cc) atomicAdd(reinterpret_cast<unsigned long long*>(&a.stats->n_valid), (unsigned long long)Cc);
        if (D) atomicAdd(reinterpret_cast<unsigned long long*>(&a.stats->n_bad), (unsigned long long)D);
      }
    }
  }
}

// resident CTAs of a strip kernel (queried once per kernel)
template <int C, int FUNC>
static int strip_slots() {
  static PerDevice cache;   // per device: occupancy depends on the device the kernel will run on
  int slots = cache.get();
  if (!slots) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lazy_strip_kernel<C, FUNC>, LZ_COLS, 0) != cudaSuccess ||
        per_sm < 1) {
      cudaGetLastError();
      per_sm = 2;
    }
    slots = per_sm * sm_count();
    cache.set(slots);
  }
  return slots;
}

// dz_low[n][c][i][j] = scale * sum over the strips / column blocks whose partial block contains node (i, j), in a fixed
// order.  CE: scale = 1 / weight_sum when size_average (read from the statistics the strip kernel produced), and
// thread 0 finalises the loss.
template <int C>
__global__ void __launch_bounds__(256)
lazy_reduce_kernel(const Geom g, const float* __restrict__ partial, float* __restrict__ dz, const CeStats* stats,
                   int size_average, float* loss) {
  const int64_t total = (int64_t)g.N * C * g.h * g.w;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (stats && loss && idx == 0) {
    double v = size_average ? stats->loss_sum / stats->weight_sum : stats->loss_sum;
    if (stats->n_bad) v = __longlong_as_double(0x7ff8000000000000LL);  // out-of-bounds target, as pointwise.cu
    *loss = (float)v;
  }
  if (idx >= total) return;
  const int j = (int)(idx % g.w);
  const int i = (int)((idx / g.w) % g.h);
  const int c = (int)((idx / ((int64_t)g.w * g.h)) % C);
  const int n = (int)(idx / ((int64_t)g.w * g.h * C));
  int s_lo = 0, s_hi = g.strips - 1, b_lo = 0, b_hi = g.xblocks - 1;
  if (g.sy > 0.f) {
    const float inv = 1.f / g.sy;
    s_lo = max(0, (int)floorf((float)(i - 1) * inv) - 1) / g.SH;
    s_hi = min(g.H - 1, (int)ceilf((float)(i + 1) * inv) + 1) / g.SH;
  }
  if (g.sx > 0.f) {
    const float inv = 1.f / g.sx;
    b_lo = max(0, (int)floorf((float)(j - 1) * inv) - 1) / LZ_COLS;
    b_hi = min(g.W - 1, (int)ceilf((float)(j + 1) * inv) + 1) / LZ_COLS;
  }
  float sum = 0.f;
  for (int s = s_lo; s <= s_hi; ++s) {
    const int k = i - lerp_at(s * g.SH, g.sy, g.h).i0;
    if (k < 0 || k >= LZ_R) continue;
    for (int b = b_lo; b <= b_hi; ++b) {
      const int jl = j - lerp_at(min(b * LZ_COLS, g.W - 1), g.sx, g.w).i0;
      if (jl < 0 || jl >= g.JB) continue;
      sum += __ldg(partial + (((int64_t)n * g.strips + s) * g.xblocks + b) * (int64_t)(LZ_R * C * g.JB) +
                   (int64_t)(k * C + c) * g.JB + jl);
    }
  }
  float scale = 1.f;
  if (stats && size_average) scale = (float)(1.0 / stats->weight_sum);
  dz[idx] = sum * scale;
}

int pack_input(const float* z_low, __nv_bfloat16* a0, int N, int C, int h, int w, int H, int W, int W0p,
               cudaStream_t st) {
  ASN_CHECK_ARG(supported(C, h, w, H, W), "lazy::pack_input: unsupported shape C=%d %dx%d -> %dx%d", C, h, w, H, W);
  Args a;
  memset(&a, 0, sizeof(a));
  a.g = make_geom(N, h, w, H, W, strip_slots<19, F_PACK>());
  a.z = z_low;
  a.a0 = a0;
  a.W0p = W0p;
  prof::Scope ps("lazy_up_softmax_pack", 0, (double)N * H * W * 64.0 + 4.0 * N * C * h * w, st);
  lazy_strip_kernel<19, F_PACK><<<dim3(a.g.xblocks, a.g.strips, N), LZ_COLS, 0, st>>>(a);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

int unpack_dx(const __nv_bfloat16* da0, const float* z_low, float* dz_low, int N, int C, int h, int w, int H, int W,
              int W0p, void* partial, size_t partial_size, cudaStream_t st) {
  ASN_CHECK_ARG(supported(C, h, w, H, W), "lazy::unpack_dx: unsupported shape C=%d %dx%d -> %dx%d", C, h, w, H, W);
  if (partial_size < partial_bytes(N, C, h, w, H, W)) {
    set_error("lazy::unpack_dx: workspace %zu < %zu", partial_size, partial_bytes(N, C, h, w, H, W));
    return ASN_EWORKSPACE;
  }
  Args a;
  memset(&a, 0, sizeof(a));
  a.g = make_geom(N, h, w, H, W, strip_slots<19, F_DBWD>());
  a.z = z_low;
  a.a0 = const_cast<__nv_bfloat16*>(da0);
  a.W0p = W0p;
  a.partial = static_cast<float*>(partial);
  {
    prof::Scope ps("lazy_up_softmax_bwd", 0, (double)N * H * W * 64.0 + 8.0 * N * C * h * w, st);
    lazy_strip_kernel<19, F_DBWD><<<dim3(a.g.xblocks, a.g.strips, N), LZ_COLS, 0, st>>>(a);
    ASN_LAUNCH_CHECK();
  }
  prof::Scope ps("lazy_up_reduce", 0, 4.0 * N * C * h * w * 7.0, st);
  lazy_reduce_kernel<19><<<full_grid((int64_t)N * C * h * w, 256), 256, 0, st>>>(a.g, a.partial, dz_low, nullptr, 0,
                                                                                 nullptr);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

}  // namespace lazy
}  // namespace asn

using namespace asn;

extern "C" int asn_upsample_ce_supported(int C, int h, int w, int H, int W) {
  return lazy::supported(C, h, w, H, W) ? 1 : 0;
}

extern "C" size_t asn_upsample_ce_workspace_bytes(int N, int C, int h, int w, int H, int W) {
  return lazy::supported(C, h, w, H, W) ? lazy::partial_bytes(N, C, h, w, H, W) : 0;
}

extern "C" int asn_upsample_ce_fwd_bwd(const float* z_low, const int64_t* y, int N, int C, int h, int w, int H, int W,
                                       int ignore_label, int mask_negative, const float* class_weight,
                                       int size_average, void* stats, float* loss, float* dz_low, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  ASN_CHECK_ARG(z_low && y && stats && loss && dz_low && workspace, "asn_upsample_ce_fwd_bwd: null pointer");
  ASN_CHECK_ARG(N > 0 && h > 0 && w > 0 && H > 0 && W > 0, "asn_upsample_ce_fwd_bwd: bad shape");
  if (!lazy::supported(C, h, w, H, W)) {
    set_error("asn_upsample_ce_fwd_bwd: unsupported case C=%d %dx%d -> %dx%d (use the unfused kernels)", C, h, w, H, W);
    return ASN_EUNSUPPORTED;
  }
  if (workspace_bytes < lazy::partial_bytes(N, C, h, w, H, W)) {
    set_error("asn_upsample_ce_fwd_bwd: workspace %zu < %zu", workspace_bytes, lazy::partial_bytes(N, C, h, w, H, W));
    return ASN_EWORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ASN_CUDA(cudaMemsetAsync(stats, 0, sizeof(CeStats), st));
  lazy::Args a;
  memset(&a, 0, sizeof(a));
  a.g = lazy::make_geom(N, h, w, H, W, lazy::strip_slots<19, lazy::F_CE>());
  a.z = z_low;
  a.y = reinterpret_cast<const long long*>(y);
  a.ignore = ignore_label;
  a.mask_negative = mask_negative;
  a.cw = class_weight;
  a.stats = static_cast<CeStats*>(stats);
  a.partial = static_cast<float*>(workspace);
  {
    prof::Scope ps("lazy_up_ce", 0, (double)N * H * W * 8.0 + 8.0 * N * C * h * w, st);
    lazy::lazy_strip_kernel<19, lazy::F_CE><<<dim3(a.g.xblocks, a.g.strips, N), lazy::LZ_COLS, 0, st>>>(a);
    ASN_LAUNCH_CHECK();
  }
  prof::Scope ps("lazy_up_reduce", 0, 4.0 * N * C * h * w * 7.0, st);
  lazy::lazy_reduce_kernel<19><<<full_grid((int64_t)N * C * h * w, 256), 256, 0, st>>>(
      a.g, a.partial, dz_low, static_cast<const CeStats*>(stats), size_average, loss);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}
