// K5 / K5b / K8: FCDiscriminator (model/discriminator.py:5-34) on tcgen05.
//
// Activations are bf16 NHWC.  A 4x4 stride-2 pad-1 convolution reads input pixel
// (2*oh - 1 + kh, 2*ow - 1 + kw); writing kh - 1 = 2*a + r (r = row parity, a in {-1,0,0,1})
// turns every tap into a unit-stride box over one of four "parity views" of the input
// (base + (r_h*W + r_w)*C, strides doubled), so one TMA box per (tap, 64-channel chunk) is the
// im2col tile of a 128-pixel output box, with the zero padding supplied by TMA's out-of-bounds
// fill.  conv1 (19 input channels) pads channels to 32 and W by one zero column on each side so
// that two horizontally adjacent taps form one 128-byte row (8 k-steps of 64).
//   forward  conv l : MODE_CONV, epilogue = bias + LeakyReLU(0.2) -> bf16 NHWC (input of l+1)
//   dgrad    conv l : 4 output-parity classes (blockIdx.z), each a 2x2-tap stride-1 conv over
//                     dPre_l; epilogue multiplies by the LeakyReLU mask of A_{l-1}
//   wgrad    conv l : MODE_WGRAD (pixels are the reduction), one CTA column per tap, split-K
//   classifier (512 -> 1) and its gradients: CUDA-core reductions (GEMV-shaped).
#include "lazy_up.cuh"
#include "umma_host.cuh"

namespace asn {

int channel_sum_nchw(const float* src, float* out, int N, int O, int P, cudaStream_t st);  // aspp.cu

constexpr float FCD_SLOPE = 0.2f;  // model/discriminator.py:16
constexpr int CLS_SLICES = 128;    // pixel slices of the classifier's weight gradient (one partial [16][C] each)

struct FcdPlan {
  int N, n_cls, ndf;
  int H[6], W[6], C[6];  // level 0 = packed input (C = 32), 1..4 conv outputs, 5 = classifier output
  int W0p;               // padded width of the packed input (W + 2 rounded up to even)
  // acts (bf16 elements offsets in bytes)
  size_t act_off[5], acts_total;
  // wpack
  size_t wf_off[5], wd_off[5], bias_off[5], wc_off, bc_off, wpack_total;  // index 1..4
  // workspace
  size_t dpre_off[5], da0_off, part_off[5], dbpart_off, clspart_off, ws_total;
  int split[5];          // wgrad split-K per layer
  size_t part_bytes;
};

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

// Tile shape of a layer's weight-gradient kernel (M = Cout, N = nn input channels per tap, K = pixels):
//   pair: CTA pairs on 256-row M tiles where Cout has an even number of 128-row tiles and nn % 128 == 0 (conv3, conv4);
//   nt  : otherwise, taps computed per CTA from ONE dY tile (A is re-used, 4 x 64 or 2 x 128 accumulator columns) --
//         the thin layers (conv1, conv2) move half the operand bytes per MAC this way.  ASN_WGRAD_TAPS=1 disables it.
static void wgrad_cfg(int Cout, int nn, int* bn, int* nt, bool* pair) {
  static const int taps_env = getenv("ASN_WGRAD_TAPS") ? atoi(getenv("ASN_WGRAD_TAPS")) : 0;
  *bn = nn >= 256 ? 256 : nn;
  *pair = umma::cluster_size(umma::MODE_WGRAD, *bn) == 2 && cdiv(Cout, 128) % 2 == 0;
  *nt = 1;
  if (!*pair && taps_env != 1) {
    if (*bn == 64) *nt = taps_env == 2 ? 2 : 4;
    else if (*bn == 128) *nt = 2;
  }
}

static void pick_tile(int OH, int OW, int px, int* th, int* tw) {
  int64_t best = -1;
  for (int h = 1; h <= px; h *= 2) {
    int w = px / h;
    if (w > 256 || h > 256) continue;
    int64_t area = (int64_t)cdiv(OH, h) * h * cdiv(OW, w) * w;
    if (best < 0 || area < best || (area == best && w > *tw)) {
      best = area;
      *th = h;
      *tw = w;
    }
  }
}

static int wgrad_taps(int l) { return l == 1 ? 8 : 16; }
static int wgrad_n(const FcdPlan& p, int l) { return l == 1 ? 64 : p.C[l - 1]; }

static int make_plan(FcdPlan& p, int N, int n_cls, int ndf, int H, int W) {
  ASN_CHECK_ARG(N > 0 && n_cls >= 1 && n_cls <= 32, "fcd: n_cls=%d outside [1,32]", n_cls);
  ASN_CHECK_ARG(ndf >= 64 && ndf % 64 == 0 && ndf <= 256, "fcd: ndf=%d must be a multiple of 64 (<= 256)", ndf);
  ASN_CHECK_ARG(H >= 32 && W >= 32, "fcd: input %dx%d smaller than the 32x32 receptive stride", H, W);
  p.N = N; p.n_cls = n_cls; p.ndf = ndf;
  p.H[0] = H; p.W[0] = W; p.C[0] = 32;
  for (int l = 1; l <= 5; ++l) {
    p.H[l] = (p.H[l - 1] + 2 - 4) / 2 + 1;
    p.W[l] = (p.W[l - 1] + 2 - 4) / 2 + 1;
    p.C[l] = l == 5 ? 1 : ndf << (l - 1);
  }
  p.W0p = (int)round_up(W + 2, 2);
  size_t off = 0;
  p.act_off[0] = off; off += align256((size_t)N * H * p.W0p * 32 * 2);
  for (int l = 1; l <= 4; ++l) { p.act_off[l] = off; off += align256((size_t)N * p.H[l] * p.W[l] * p.C[l] * 2); }
  p.acts_total = off;
  off = 0;
  for (int l = 1; l <= 4; ++l) {
    const size_t kf = l == 1 ? 8 * 64 : (size_t)16 * p.C[l - 1];           // forward K
    p.wf_off[l] = off; off += align256((size_t)p.C[l] * kf * 2);
    const size_t rows = l == 1 ? 32 : p.C[l - 1];                           // dgrad rows per parity class
    p.wd_off[l] = off; off += align256((size_t)4 * rows * 4 * p.C[l] * 2);
    p.bias_off[l] = off; off += align256((size_t)p.C[l] * 4);
  }
  p.wc_off = off; off += align256((size_t)16 * p.C[4] * 4);
  p.bc_off = off; off += 256;
  p.wpack_total = off;
  off = 0;
  for (int l = 1; l <= 4; ++l) { p.dpre_off[l] = off; off += align256((size_t)N * p.H[l] * p.W[l] * p.C[l] * 2); }
  p.da0_off = off; off += align256((size_t)N * H * p.W0p * 32 * 2);
  p.part_bytes = 0;
  for (int l = 1; l <= 4; ++l) {
    int th = 1, tw = 64;
    pick_tile(p.H[l], p.W[l], 64, &th, &tw);
    const int k_steps = N * cdiv(p.H[l], th) * cdiv(p.W[l], tw);
    const int nn = wgrad_n(p, l);
    int bn, nt;
    bool pair;
    wgrad_cfg(p.C[l], nn, &bn, &nt, &pair);
    const int ctas = cdiv(p.C[l], 128) * cdiv(nn, bn) * (wgrad_taps(l) / nt);
    // one wave of the persistent grid: CTA pairs (one per TPC), multi-tap CTAs (one per SM) or plain single CTAs
    // (two per SM when BLOCK_N <= 128)
    const int slots = pair ? sm_count() / 2 : sm_count() * (nt == 1 && bn <= 128 ? 2 : 1);
    int S = pair ? slots / (ctas / 2) : (slots + ctas - 1) / ctas;
    if (S > k_steps / 2) S = k_steps / 2;
    if (S < 1) S = 1;
    const int sps = cdiv(k_steps, S);
    p.split[l] = cdiv(k_steps, sps);
    const size_t bytes = (size_t)p.split[l] * wgrad_taps(l) * p.C[l] * nn * 4;
    p.part_off[l] = off; off += align256(bytes);   // every layer keeps its partials: one merged reduce at the end
    p.part_bytes += bytes;
  }
  p.dbpart_off = off; off += align256((size_t)4 * 256 * 2048 * 4);
  p.clspart_off = off; off += align256((size_t)CLS_SLICES * 16 * p.C[4] * 4);   // classifier weight-gradient partials
  p.ws_total = off;
  return ASN_OK;
}

// kh (or kw) - 1 = 2*a + r
__host__ __device__ inline int tap_a(int k) { return k == 0 ? -1 : (k == 3 ? 1 : 0); }
__host__ __device__ inline int tap_r(int k) { return (k == 0 || k == 2) ? 1 : 0; }

// ---- input pack: fp32 NCHW (probabilities or logits) -> bf16 [N][H][W0p][32] ------------------
// One thread = PV consecutive pixels of a row (PV = 4 with 16-byte loads per channel plane when W % 4 == 0):
// the C values per pixel stay in registers, the optional channel softmax (K4) is fused, and each pixel's
// 32 bf16 (64 bytes) are written with 16-byte stores.  Pixel w lands at padded column w + 1; the zero columns
// (0 and >= W + 1) are written by the thread that owns the neighbouring pixels.
__device__ __forceinline__ void store_px32(__nv_bfloat16* dst, const float* v) {
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

constexpr float LOG2E_F = 1.4426950408889634f;  // exp(x - m) = exp2(x * log2e - m * log2e): one FFMA + ex2

template <int PV>
__global__ void __launch_bounds__(128)
fcd_pack_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ a0, int N, int C, int H, int W,
                      int W0p, int softmax) {
  const int wv = (W + PV - 1) / PV;
  const int64_t total = (int64_t)N * H * wv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xv = (int)(i % wv);
    const int64_t row = i / wv;               // n * H + h
    const int h = (int)(row % H);
    const int n = (int)(row / H);
    const int w0 = xv * PV;
    float v[PV][32];
#pragma unroll
    for (int k = 0; k < PV; ++k)
#pragma unroll
      for (int c = 0; c < 32; ++c) v[k][c] = 0.f;
    const float* src = x + ((int64_t)n * C * H + h) * W + w0;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      if (c < C) {
        if (PV == 4) {
          const float4 t = ld_stream(reinterpret_cast<const float4*>(src + (int64_t)c * H * W));
          v[0][c] = t.x; v[1][c] = t.y; v[2][c] = t.z; v[3][c] = t.w;
        } else {
          v[0][c] = __ldg(src + (int64_t)c * H * W);
        }
      }
    }
    if (softmax) {
#pragma unroll
      for (int k = 0; k < PV; ++k) {
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c < C) m = fmaxf(m, v[k][c]);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c < C) { v[k][c] = exp2f(fmaf(v[k][c], LOG2E_F, -m * LOG2E_F)); s += v[k][c]; }  // as lazy_up.cu
        const float inv = 1.f / s;
#pragma unroll
        for (int c = 0; c < 32; ++c) v[k][c] *= inv;
      }
    }
    __nv_bfloat16* dst_row = a0 + row * (int64_t)W0p * 32;
#pragma unroll
    for (int k = 0; k < PV; ++k)
      if (w0 + k < W) store_px32(dst_row + (int64_t)(w0 + k + 1) * 32, v[k]);
    // zero padding columns of this row
    float z[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) z[c] = 0.f;
    if (xv == 0) store_px32(dst_row, z);
    if (xv == wv - 1)
      for (int wp = W + 1; wp < W0p; ++wp) store_px32(dst_row + (int64_t)wp * 32, z);
  }
}

// dA0 bf16 [N][H][W0p][32] -> dx fp32 NCHW; with logits given, the channel-softmax backward is fused:
// dz = p * (g - sum_c p*g), p = softmax(x).  PV pixels per thread (PV = 2: 8-byte accesses per channel plane;
// 4 pixels would need 256 live floats and spill).
template <int PV>
__global__ void __launch_bounds__(128)
fcd_unpack_dx_kernel(const __nv_bfloat16* __restrict__ da0, const float* __restrict__ logits,
                     float* __restrict__ dx, int N, int C, int H, int W, int W0p) {
  const int wv = (W + PV - 1) / PV;
  const int64_t total = (int64_t)N * H * wv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xv = (int)(i % wv);
    const int64_t row = i / wv;
    const int h = (int)(row % H);
    const int n = (int)(row / H);
    const int w0 = xv * PV;
    float g[PV][32];
#pragma unroll
    for (int k = 0; k < PV; ++k) {
      const uint4* src = reinterpret_cast<const uint4*>(da0 + (row * (int64_t)W0p + min(w0 + k, W - 1) + 1) * 32);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 u = __ldg(src + q);
        const uint32_t pk[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(&pk[j]);
          g[k][q * 8 + 2 * j] = __low2float(hh);
          g[k][q * 8 + 2 * j + 1] = __high2float(hh);
        }
      }
    }
    const int64_t base = ((int64_t)n * C * H + h) * W + w0;
    if (logits) {
      float z[PV][32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        if (c < C) {
          if (PV == 2) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(logits + base + (int64_t)c * H * W));
            z[0][c] = t.x; z[1][c] = t.y;
          } else {
            z[0][c] = __ldg(logits + base + (int64_t)c * H * W);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < PV; ++k) {
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c < C) m = fmaxf(m, z[k][c]);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c < C) { z[k][c] = exp2f(fmaf(z[k][c], LOG2E_F, -m * LOG2E_F)); s += z[k][c]; }
        const float inv = 1.f / s;
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c < C) { z[k][c] *= inv; dot += z[k][c] * g[k][c]; }
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c < C) g[k][c] = z[k][c] * (g[k][c] - dot);
      }
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      if (c < C) {
        if (PV == 2) {
          *reinterpret_cast<float2*>(dx + base + (int64_t)c * H * W) = make_float2(g[0][c], g[1][c]);
        } else {
          dx[base + (int64_t)c * H * W] = g[0][c];
        }
      }
    }
  }
}

// ---- weight packing ---------------------------------------------------------------------------
struct FcdParamPtrs {
  const float* w[5];
  const float* b[5];
};

// forward pack  Wf_l[co][k]:  l >= 2: k = (kh*4+kw)*Cin + ci ; l == 1: k = (kh*2+pw)*64 + kwl*32 + c, kw = 2*pw + kwl
// dgrad pack    Wd_l[z=(rh,rw)][ci][k], k = (th*2+tw)*Cout + co with kh = kh(rh, th), kw = kw(rw, tw):
//               rh = 0 -> kh in {1,3}, rh = 1 -> kh in {0,2}
struct PackArgs {  // blockIdx.y = 0..3: conv1..conv4, 4: classifier -- one launch packs the whole discriminator
  const float* w[5];
  const float* b[5];
  __nv_bfloat16* wf[4];
  __nv_bfloat16* wd[4];
  float* bias[4];
  int Cout[4], Cin_real[4], Cin_rows[4];
  float* wc;
  float* bc;
  int C4;
};

// One CTA per (32 output channels x 16 input channels) block of a layer: its 32 x 16 x 16 fp32 weights are read as
// 1 KB contiguous runs into shared memory and leave as 32- / 64-byte runs of the two bf16 layouts (ci fastest in Wf, co
// fastest in Wd) -- every global access of the pack is a full sector.
constexpr int PK_CO = 32, PK_CI = 16;
__global__ void __launch_bounds__(256) fcd_pack_all_kernel(const PackArgs a) {
  if (blockIdx.y == 4) {  // classifier weights [1][C][4][4] -> fp32 [16][C]
    for (int i = blockIdx.x * 256 + threadIdx.x; i < 16 * a.C4; i += gridDim.x * 256) {
      const int t = i / a.C4, c = i % a.C4;
      a.wc[i] = __ldg(a.w[4] + (int64_t)c * 16 + t);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) a.bc[0] = __ldg(a.b[4]);
    return;
  }
  const int li = blockIdx.y, l = li + 1;
  const float* __restrict__ w = a.w[li];
  __nv_bfloat16* __restrict__ wf = a.wf[li];
  __nv_bfloat16* __restrict__ wd = a.wd[li];
  const int Cout = a.Cout[li], Cin_real = a.Cin_real[li], Cin_rows = a.Cin_rows[li];
  const int ci_blocks = (Cin_rows + PK_CI - 1) / PK_CI, co_blocks = Cout / PK_CO;  // Cout is a multiple of 64
  if ((int)blockIdx.x >= ci_blocks * co_blocks) return;
  const int co0 = (blockIdx.x / ci_blocks) * PK_CO, ci0 = (blockIdx.x % ci_blocks) * PK_CI;
  __shared__ float t[PK_CO][PK_CI * 16 + 1];  // [co][ci * 16 + kh * 4 + kw]
  for (int e = threadIdx.x; e < PK_CO * PK_CI * 16; e += 256) {
    const int col = e % (PK_CI * 16), co = e / (PK_CI * 16);
    const int ci = ci0 + col / 16;
    t[co][col] = ci < Cin_real ? __ldg(w + ((int64_t)(co0 + co) * Cin_real + ci0) * 16 + col) : 0.f;  // padded channels: 0
  }
  if (blockIdx.x % ci_blocks == 0 && threadIdx.x < PK_CO) a.bias[li][co0 + threadIdx.x] = __ldg(a.b[li] + co0 + threadIdx.x);
  __syncthreads();
  // forward pack Wf[co][k]: conv1 k = (kh*2 + kw/2)*64 + (kw%2)*32 + ci (512 per row); else k = (kh*4 + kw)*Cin + ci
  const int Kf = l == 1 ? 512 : 16 * Cin_real;
  for (int e = threadIdx.x; e < PK_CO * 16 * PK_CI; e += 256) {
    const int cil = e % PK_CI, tap = (e / PK_CI) % 16, co = e / (PK_CI * 16);
    const int kh = tap >> 2, kw = tap & 3;
    const int ci = ci0 + cil;
    if (ci >= Cin_rows) continue;
    const int k = l == 1 ? (kh * 2 + (kw >> 1)) * 64 + (kw & 1) * 32 + ci : tap * Cin_real + ci;
    wf[(int64_t)(co0 + co) * Kf + k] = __float2bfloat16(t[co][cil * 16 + tap]);
  }
  // dgrad pack Wd[z = (rh,rw)][ci][(th*2 + tw)*Cout + co], kh = kh(rh, th): rh = 0 -> {1,3}, rh = 1 -> {0,2}
  const int Kd = 4 * Cout;
  for (int e = threadIdx.x; e < 16 * PK_CI * PK_CO; e += 256) {
    const int col = e % PK_CO, cil = (e / PK_CO) % PK_CI, zt = e / (PK_CO * PK_CI);   // zt = z*4 + t
    const int z = zt >> 2, tt = zt & 3;
    const int rh = z >> 1, rw = z & 1, th = tt >> 1, tw = tt & 1;
    const int kh = rh == 0 ? (th == 0 ? 1 : 3) : (th == 0 ? 0 : 2);
    const int kw = rw == 0 ? (tw == 0 ? 1 : 3) : (tw == 0 ? 0 : 2);
    const int ci = ci0 + cil;
    if (ci >= Cin_rows) continue;
    wd[((int64_t)z * Cin_rows + ci) * Kd + tt * Cout + co0 + col] = __float2bfloat16(t[col][cil * 16 + kh * 4 + kw]);
  }
}

// Round-2 form of the pack: a flat block table (no empty CTAs) and 16-byte accesses on both sides for conv2..conv4.
// A CTA owns the same (32 output channels x 16 input channels) block: its 8192 fp32 weights arrive as eight independent
// float4 loads per thread, and leave as 16-byte stores -- 8 consecutive input channels of one (co, tap) of Wf, 8 consecutive
// output channels of one (parity class, tap, ci) of Wd -- gathered from shared memory (pitch 260 floats: the gathers of a
// warp hit 32 different banks).  conv1 (19 -> 32 padded channels, 64 x 19 x 16 weights) and the classifier keep the scalar code.
constexpr int PK_PITCH = PK_CI * 16 + 4;
struct PackTable {
  int seg_end[5];   // conv1..conv4 blocks, classifier blocks
};
__global__ void __launch_bounds__(256) fcd_pack_vec_kernel(const PackArgs a, const PackTable tb) {
  int li = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) li += (int)blockIdx.x >= tb.seg_end[i] ? 1 : 0;
  int blk = blockIdx.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) blk = (li == i + 1) ? (int)blockIdx.x - tb.seg_end[i] : blk;
  if (li == 4) {  // classifier weights [1][C][4][4] -> fp32 [16][C]
    const int i = blk * 256 + threadIdx.x;
    if (i < 16 * a.C4) {
      const int t = i / a.C4, c = i % a.C4;
      a.wc[i] = __ldg(a.w[4] + (int64_t)c * 16 + t);
    }
    if (blk == 0 && threadIdx.x == 0) a.bc[0] = __ldg(a.b[4]);
    return;
  }
  const int l = li + 1;
  const float* __restrict__ w = a.w[li];
  __nv_bfloat16* __restrict__ wf = a.wf[li];
  __nv_bfloat16* __restrict__ wd = a.wd[li];
  const int Cout = a.Cout[li], Cin_real = a.Cin_real[li], Cin_rows = a.Cin_rows[li];
  const int ci_blocks = (Cin_rows + PK_CI - 1) / PK_CI;
  const int co0 = (blk / ci_blocks) * PK_CO, ci0 = (blk % ci_blocks) * PK_CI;
  __shared__ float t[PK_CO][PK_PITCH];  // [co][ci * 16 + kh * 4 + kw]
  if (blk % ci_blocks == 0 && threadIdx.x < PK_CO) a.bias[li][co0 + threadIdx.x] = __ldg(a.b[li] + co0 + threadIdx.x);
  if (l == 1) {   // scalar path (padded channels: 0)
    for (int e = threadIdx.x; e < PK_CO * PK_CI * 16; e += 256) {
      const int col = e % (PK_CI * 16), co = e / (PK_CI * 16);
      const int ci = ci0 + col / 16;
      t[co][col] = ci < Cin_real ? __ldg(w + ((int64_t)(co0 + co) * Cin_real + ci0) * 16 + col) : 0.f;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < PK_CO * 16 * PK_CI; e += 256) {
      const int cil = e % PK_CI, tap = (e / PK_CI) % 16, co = e / (PK_CI * 16);
      const int kh = tap >> 2, kw = tap & 3;
      const int ci = ci0 + cil;
      if (ci >= Cin_rows) continue;
      wf[(int64_t)(co0 + co) * 512 + (kh * 2 + (kw >> 1)) * 64 + (kw & 1) * 32 + ci] = __float2bfloat16(t[co][cil * 16 + tap]);
    }
    const int Kd1 = 4 * Cout;
    for (int e = threadIdx.x; e < 16 * PK_CI * PK_CO; e += 256) {
      const int col = e % PK_CO, cil = (e / PK_CO) % PK_CI, zt = e / (PK_CO * PK_CI);
      const int z = zt >> 2, tt = zt & 3;
      const int rh = z >> 1, rw = z & 1, th = tt >> 1, tw = tt & 1;
      const int kh = rh == 0 ? (th == 0 ? 1 : 3) : (th == 0 ? 0 : 2);
      const int kw = rw == 0 ? (tw == 0 ? 1 : 3) : (tw == 0 ? 0 : 2);
      const int ci = ci0 + cil;
      if (ci >= Cin_rows) continue;
      wd[((int64_t)z * Cin_rows + ci) * Kd1 + tt * Cout + co0 + col] = __float2bfloat16(t[col][cil * 16 + kh * 4 + kw]);
    }
    return;
  }
  // conv2..conv4: Cin_real == Cin_rows, a multiple of 16; every block is full
  {
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int e4 = threadIdx.x + 256 * k;                 // float4 index in the [32][256] block
      const int co = e4 >> 6, col = (e4 & 63) * 4;
      v[k] = ld_stream(reinterpret_cast<const float4*>(w + ((int64_t)(co0 + co) * Cin_real + ci0) * 16 + col));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int e4 = threadIdx.x + 256 * k;
      const int co = e4 >> 6, col = (e4 & 63) * 4;
      *reinterpret_cast<float4*>(&t[co][col]) = v[k];
    }
  }
  __syncthreads();
  auto pack8 = [](const float (&f)[8]) {
    uint32_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      r[i] = *reinterpret_cast<const uint32_t*>(&b);
    }
    return make_uint4(r[0], r[1], r[2], r[3]);
  };
  // forward pack Wf[co][tap * Cin + ci]: thread = (co, tap, half of the 16 input channels)
  const int Kf = 16 * Cin_real;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = threadIdx.x + 256 * k;
    const int half = e & 1, tap = (e >> 1) & 15, co = e >> 5;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = t[co][(half * 8 + i) * 16 + tap];
    *reinterpret_cast<uint4*>(wf + (int64_t)(co0 + co) * Kf + tap * Cin_real + ci0 + half * 8) = pack8(f);
  }
  // dgrad pack Wd[z = (rh,rw)][ci][(th*2 + tw)*Cout + co]: thread = (kh*4+kw, ci, 8 output channels)
  const int Kd = 4 * Cout;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = threadIdx.x + 256 * k;
    const int khkw = e & 15, cil = (e >> 4) & 15, g = e >> 8;
    const int kh = khkw >> 2, kw = khkw & 3;
    // kh = 1,3 belong to row parity rh = 0 (th = 0,1); kh = 0,2 to rh = 1 (th = 0,1); the same for kw
    const int rh = (kh & 1) ? 0 : 1, th = kh >> 1, rw = (kw & 1) ? 0 : 1, tw = kw >> 1;
    const int z = rh * 2 + rw, tt = th * 2 + tw;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = t[g * 8 + i][cil * 16 + khkw];
    *reinterpret_cast<uint4*>(wd + ((int64_t)z * Cin_rows + ci0 + cil) * Kd + tt * Cout + co0 + g * 8) = pack8(f);
  }
}

// ---- classifier (N = 1 output channel): CUDA-core reductions ------------------------------------
// out[n][oh][ow] = bc + sum_{kh,kw,c} A4[n][2oh-1+kh][2ow-1+kw][c] * wc[kh*4+kw][c].
// One 128-thread CTA per output pixel: warp = kernel row kh, lanes stride the channels with 16-byte loads.
__global__ void __launch_bounds__(128)
fcd_cls_fwd_kernel(const __nv_bfloat16* __restrict__ a4, const float* __restrict__ wc, const float* __restrict__ bc,
                   float* __restrict__ out, int N, int H4, int W4, int C, int H5, int W5) {
  __shared__ float red[4];
  const int o = blockIdx.x;
  const int kh = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ow = o % W5, oh = (o / W5) % H5, n = o / (W5 * H5);
  float acc = 0.f;
  const int ih = 2 * oh - 1 + kh;
  if ((unsigned)ih < (unsigned)H4) {
    for (int kw = 0; kw < 4; ++kw) {
      const int iw = 2 * ow - 1 + kw;
      if ((unsigned)iw >= (unsigned)W4) continue;
      const __nv_bfloat16* src = a4 + (((int64_t)n * H4 + ih) * W4 + iw) * C;
      const float* wt = wc + (kh * 4 + kw) * C;
      for (int c = lane * 8; c < C; c += 256) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + c));
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wt + c));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wt + c + 4));
        const uint32_t pk[4] = {u.x, u.y, u.z, u.w};
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&pk[j]);
          acc = fmaf(__low2float(v), wv[2 * j], acc);
          acc = fmaf(__high2float(v), wv[2 * j + 1], acc);
        }
      }
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) red[kh] = acc;
  __syncthreads();
  if (threadIdx.x == 0) out[o] = red[0] + red[1] + red[2] + red[3] + __ldg(bc);
}

// dPre4[n][ih][iw][c] = mask(A4) * sum_{kh,kw valid} dout[n][oh][ow] * wc[kh*4+kw][c]
__global__ void __launch_bounds__(256)
fcd_cls_dgrad_kernel(const float* __restrict__ dout, const float* __restrict__ wc,
                     const __nv_bfloat16* __restrict__ a4, __nv_bfloat16* __restrict__ dpre4, int N, int H4, int W4,
                     int C, int H5, int W5) {
  const int64_t total = (int64_t)N * H4 * W4 * (C / 2);
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int c = (int)(i % (C / 2)) * 2;
    const int64_t px = i / (C / 2);
    const int iw = (int)(px % W4), ih = (int)((px / W4) % H4), n = (int)(px / ((int64_t)W4 * H4));
    float g0 = 0.f, g1 = 0.f;
    for (int kh = (ih + 1) & 1; kh < 4; kh += 2) {
      const int oh = (ih + 1 - kh) / 2;
      if (ih + 1 - kh < 0 || oh >= H5) continue;
      for (int kw = (iw + 1) & 1; kw < 4; kw += 2) {
        const int ow = (iw + 1 - kw) / 2;
        if (iw + 1 - kw < 0 || ow >= W5) continue;
        const float d = __ldg(dout + ((int64_t)n * H5 + oh) * W5 + ow);
        g0 = fmaf(d, __ldg(wc + (kh * 4 + kw) * C + c), g0);
        g1 = fmaf(d, __ldg(wc + (kh * 4 + kw) * C + c + 1), g1);
      }
    }
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(a4 + px * C + c);
    g0 *= __low2float(a) > 0.f ? 1.f : FCD_SLOPE;
    g1 *= __high2float(a) > 0.f ? 1.f : FCD_SLOPE;
    *reinterpret_cast<__nv_bfloat162*>(dpre4 + px * C + c) = __floats2bfloat162_rn(g0, g1);
  }
}

// Same gradient, thread = 8 consecutive channels of one pixel (16-byte accesses of A4, dPre4 and the fp32 tap weights; the pixel's
// index arithmetic and its <= 4 dout values are shared by 8 channels instead of 2).  Same fmaf order per channel: bit-identical.
__global__ void __launch_bounds__(256)
fcd_cls_dgrad8_kernel(const float* __restrict__ dout, const float* __restrict__ wc,
                      const __nv_bfloat16* __restrict__ a4, __nv_bfloat16* __restrict__ dpre4, int N, int H4, int W4,
                      int C, int H5, int W5) {
  const int groups = C / 8;
  const int64_t total = (int64_t)N * H4 * W4 * groups;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % groups) * 8;
  const int64_t px = i / groups;
  const int iw = (int)(px % W4), ih = (int)((px / W4) % H4), n = (int)(px / ((int64_t)W4 * H4));
  const uint4 av = __ldg(reinterpret_cast<const uint4*>(a4 + px * C + c));
  float g[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) g[e] = 0.f;
  for (int kh = (ih + 1) & 1; kh < 4; kh += 2) {
    const int oh = (ih + 1 - kh) / 2;
    if (ih + 1 - kh < 0 || oh >= H5) continue;
    for (int kw = (iw + 1) & 1; kw < 4; kw += 2) {
      const int ow = (iw + 1 - kw) / 2;
      if (iw + 1 - kw < 0 || ow >= W5) continue;
      const float d = __ldg(dout + ((int64_t)n * H5 + oh) * W5 + ow);
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wc + (kh * 4 + kw) * C + c));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wc + (kh * 4 + kw) * C + c + 4));
      g[0] = fmaf(d, w0.x, g[0]); g[1] = fmaf(d, w0.y, g[1]); g[2] = fmaf(d, w0.z, g[2]); g[3] = fmaf(d, w0.w, g[3]);
      g[4] = fmaf(d, w1.x, g[4]); g[5] = fmaf(d, w1.y, g[5]); g[6] = fmaf(d, w1.z, g[6]); g[7] = fmaf(d, w1.w, g[7]);
    }
  }
  const uint32_t aw[4] = {av.x, av.y, av.z, av.w};
  uint32_t ow4[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&aw[j]);
    const float g0 = g[2 * j] * (__low2float(a) > 0.f ? 1.f : FCD_SLOPE);
    const float g1 = g[2 * j + 1] * (__high2float(a) > 0.f ? 1.f : FCD_SLOPE);
    const __nv_bfloat162 o = __floats2bfloat162_rn(g0, g1);
    ow4[j] = *reinterpret_cast<const uint32_t*>(&o);
  }
  *reinterpret_cast<uint4*>(dpre4 + px * C + c) = make_uint4(ow4[0], ow4[1], ow4[2], ow4[3]);
}

// dwc[0][c][kh][kw] = sum_{n,oh,ow} dout * A4[n][2oh-1+kh][2ow-1+kw][c].
// grid (CLS_SLICES, C/64): thread = (channel, 1 of 4 pixel phases).  Every INPUT pixel of the slice is read once
// (coalesced 128-byte rows along the channels) and feeds the <= 4 taps whose stride-2 footprint contains it -- which
// taps is a matter of the pixel's row/column parity, uniform over the CTA, so the 16 accumulators stay in registers.
// Shared-memory reduce over the phases, then ONE plain store per (slice, tap, channel) into `part`; the slices are summed in
// a fixed order by the merged end-of-backward reduce (fcd_wgrad_reduce_kernel, blockIdx.y == 4) -- no atomics, bit-reproducible (round 1 used fp32 atomics here).

template <int KH, int KW>
__device__ __forceinline__ void cls_tap(float (&acc)[16], float v, const float* __restrict__ dout, int n, int ih, int iw,
                                        int H5, int W5) {
  // output pixel that sees input (ih, iw) through tap (KH, KW): ih = 2*oh - 1 + KH
  const int oh2 = ih + 1 - KH, ow2 = iw + 1 - KW;
  if (oh2 < 0 || ow2 < 0) return;
  const int oh = oh2 >> 1, ow = ow2 >> 1;
  if (oh >= H5 || ow >= W5) return;
  acc[KH * 4 + KW] = fmaf(__ldg(dout + ((int64_t)n * H5 + oh) * W5 + ow), v, acc[KH * 4 + KW]);
}

__global__ void __launch_bounds__(256)
fcd_cls_wgrad_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ a4, float* __restrict__ part,
                     int N, int H4, int W4, int C, int H5, int W5) {
  __shared__ float red[4][16][64];
  const int cl = threadIdx.x & 63, ph = threadIdx.x >> 6;
  const int c = blockIdx.y * 64 + cl;
  const int n_in = N * H4 * W4;
  float acc[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) acc[t] = 0.f;
  if (c < C) {
    // the activations of up to CLS_UNR pixels of this thread are loaded first (independent 2-byte loads, one DRAM round trip
    // instead of one per pixel), then folded into the taps in pixel order
    constexpr int CLS_UNR = 8;
    for (int i0 = blockIdx.x * 4 + ph; i0 < n_in; i0 += 4 * CLS_SLICES * CLS_UNR) {
      float vv[CLS_UNR];
#pragma unroll
      for (int u = 0; u < CLS_UNR; ++u) {
        const int i = i0 + u * 4 * CLS_SLICES;
        vv[u] = i < n_in ? __bfloat162float(a4[(int64_t)i * C + c]) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < CLS_UNR; ++u) {
        const int i = i0 + u * 4 * CLS_SLICES;
        if (i >= n_in) break;
        const int iw = i % W4, ih = (i / W4) % H4, n = i / (W4 * H4);
        const float v = vv[u];
        // kh = ih + 1 (mod 2), kw = iw + 1 (mod 2)
        if (ih & 1) {
          if (iw & 1) {
            cls_tap<0, 0>(acc, v, dout, n, ih, iw, H5, W5); cls_tap<0, 2>(acc, v, dout, n, ih, iw, H5, W5);
            cls_tap<2, 0>(acc, v, dout, n, ih, iw, H5, W5); cls_tap<2, 2>(acc, v, dout, n, ih, iw, H5, W5);
          } else {
            cls_tap<0, 1>(acc, v, dout, n, ih, iw, H5, W5); cls_tap<0, 3>(acc, v, dout, n, ih, iw, H5, W5);
            cls_tap<2, 1>(acc, v, dout, n, ih, iw, H5, W5); cls_tap<2, 3>(acc, v, dout, n, ih, iw, H5, W5);
          }
        } else {
          if (iw & 1) {
            cls_tap<1, 0>(acc, v, dout, n, ih, iw, H5, W5); cls_tap<1, 2>(acc, v, dout, n, ih, iw, H5, W5);
            cls_tap<3, 0>(acc, v, dout, n, ih, iw, H5, W5); cls_tap<3, 2>(acc, v, dout, n, ih, iw, H5, W5);
          } else {
            cls_tap<1, 1>(acc, v, dout, n, ih, iw, H5, W5); cls_tap<1, 3>(acc, v, dout, n, ih, iw, H5, W5);
            cls_tap<3, 1>(acc, v, dout, n, ih, iw, H5, W5); cls_tap<3, 3>(acc, v, dout, n, ih, iw, H5, W5);
          }
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 16; ++t) red[ph][t][cl] = acc[t];
  __syncthreads();
  // 64 channels x 16 taps = 1024 sums over the 4 phases; part[slice][tap][channel]: consecutive threads -> consecutive channels
  for (int e = threadIdx.x; e < 64 * 16; e += 256) {
    const int t = e >> 6, ch = e & 63;
    if (blockIdx.y * 64 + ch < C)
      part[((int64_t)blockIdx.x * 16 + t) * C + blockIdx.y * 64 + ch] = red[0][t][ch] + red[1][t][ch] + red[2][t][ch] + red[3][t][ch];
  }
}

// Round-2 form: thread = (8 consecutive channels, ONE parity class of input pixels).  The four taps an input pixel feeds are a
// function of its (row, column) parity alone, so a thread that only visits pixels of one class keeps 4 x 8 accumulators
// (not 16 per channel), reads 16 bytes per pixel and amortises the index arithmetic over 8 channels (the per-channel form
// above spends ~170 instructions per pixel and channel on it: instruction-bound at 21 us).  No shared memory: the thread
// owns its (slice, 4 taps, 8 channels) outputs.  blockDim = C/8 groups x 4 classes; grid = CLS_SLICES.
__global__ void __launch_bounds__(256)
fcd_cls_wgrad_par_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ a4, float* __restrict__ part,
                         int N, int H4, int W4, int C, int H5, int W5) {
  const int groups = C / 8;
  const int g = threadIdx.x % groups, pc = threadIdx.x / groups;     // pc = (ih & 1) * 2 + (iw & 1)
  const int pa = pc >> 1, pb = pc & 1;
  const int nh = (H4 - pa + 1) / 2, nw = (W4 - pb + 1) / 2;           // pixels of this class: ih = 2*mh + pa, iw = 2*mw + pb
  const int total = N * nh * nw;
  float acc[4][8];
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
  constexpr int UNR = 8;
  for (int q0 = blockIdx.x; q0 < total; q0 += CLS_SLICES * UNR) {
    uint4 u[UNR];
    int on[UNR], om[UNR], ow_[UNR];
#pragma unroll
    for (int k = 0; k < UNR; ++k) {
      const int q = q0 + k * CLS_SLICES;
      u[k] = make_uint4(0u, 0u, 0u, 0u);
      on[k] = -1;
      if (q < total) {
        const int n = q / (nh * nw), rem = q - n * (nh * nw);
        const int mh = rem / nw, mw = rem - mh * nw;
        on[k] = n; om[k] = mh; ow_[k] = mw;
        u[k] = __ldg(reinterpret_cast<const uint4*>(a4 + (((int64_t)n * H4 + 2 * mh + pa) * W4 + 2 * mw + pb) * C) + g);
      }
    }
#pragma unroll
    for (int k = 0; k < UNR; ++k) {
      if (on[k] < 0) break;
      float v[8];
      const uint32_t pk[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(&pk[j]);
        v[2 * j] = __low2float(hh);
        v[2 * j + 1] = __high2float(hh);
      }
      // tap (kh, kw) = (1 - pa + 2i, 1 - pb + 2j) sees this pixel from output (mh + pa - i, mw + pb - j)
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int oh = om[k] + pa - i, ow = ow_[k] + pb - j;
          const bool ok = oh >= 0 && oh < H5 && ow >= 0 && ow < W5;
          const float d = ok ? __ldg(dout + ((int64_t)on[k] * H5 + oh) * W5 + ow) : 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[i * 2 + j][e] = fmaf(d, v[e], acc[i * 2 + j][e]);
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int tap = (1 - pa + 2 * i) * 4 + (1 - pb + 2 * j);
      float4* dst = reinterpret_cast<float4*>(part + ((int64_t)blockIdx.x * 16 + tap) * C + g * 8);
      dst[0] = make_float4(acc[i * 2 + j][0], acc[i * 2 + j][1], acc[i * 2 + j][2], acc[i * 2 + j][3]);
      dst[1] = make_float4(acc[i * 2 + j][4], acc[i * 2 + j][5], acc[i * 2 + j][6], acc[i * 2 + j][7]);
    }
}

// ---- reductions at the end of the backward: one launch each for the four conv layers (blockIdx.y = layer) ----
struct LayerReduce {
  // bias gradient: db[c] = sum over rows of dpre[row][c]  (bf16 [P][C])
  const __nv_bfloat16* dpre[4];
  float* db[4];
  float* db_partial[4];
  long long P[4];
  int C[4], rows_per_cta[4], ctas[4];
  // weight gradient: dW[co][ci][kh][kw] = sum_z part[z][tap][co][col]
  const float* part[4];
  float* dw[4];
  int S[4], Cout[4], Cin_real[4], Ncols[4];
  // classifier weight gradient: dwc[c][tap] = sum over the CLS_SLICES partials [slice][tap][c] (blockIdx.y == 4)
  const float* cls_part;
  float* cls_dw;
  int cls_C;
  // classifier bias gradient: cls_db[0] = sum of dout (cls_n values); merged launch only
  const float* cls_dout;
  float* cls_db;
  int cls_n;
  // merged launch (fcd_reduce_tail_kernel): block ranges  [0..3] final bias sums of conv1..4, [4] classifier bias gradient,
  // [5..8] weight gradient of conv1..4, [9] classifier weight gradient; seg_end[i] = first block past segment i
  int seg_end[10];
};

// partial[cta][c] = sum of this CTA's row slice (fixed order -> deterministic).  Thread = 8 consecutive channels (one 16-byte
// load per row); the 256 threads split into C/8 channel groups x (2048/C) row phases, each phase striding the rows of the
// slice with four independent loads in flight; the phases are folded through shared memory in a fixed order.
__global__ void __launch_bounds__(256) fcd_colsum_partial_kernel(LayerReduce R) {
  const int l = blockIdx.y;
  if ((int)blockIdx.x >= R.ctas[l]) return;
  const __nv_bfloat16* src = R.dpre[l];
  const int C = R.C[l];                         // 64 .. 512, a multiple of 64
  const long long r0 = (long long)blockIdx.x * R.rows_per_cta[l];
  const long long r1 = min(R.P[l], r0 + R.rows_per_cta[l]);
  const int groups = C / 8;                     // 8 .. 64 threads cover one row
  const int nsub = 256 / groups;                // row phases
  const int g = threadIdx.x % groups, sub = threadIdx.x / groups;
  __shared__ float sm[256 * 8];                 // [sub][channel]
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  auto add = [&](const uint4& u) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      acc[2 * i] += __low2float(v);
      acc[2 * i + 1] += __high2float(v);
    }
  };
  long long r = r0 + sub;
  for (; r + 3LL * nsub < r1; r += 4LL * nsub) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = ld_stream(reinterpret_cast<const uint4*>(src + (r + (long long)k * nsub) * C) + g);
#pragma unroll
    for (int k = 0; k < 4; ++k) add(u[k]);
  }
  for (; r < r1; r += nsub) add(ld_stream(reinterpret_cast<const uint4*>(src + r * C) + g));
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[(sub * groups + g) * 8 + i] = acc[i];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float t = 0.f;
    for (int k = 0; k < nsub; ++k) t += sm[k * C + c];     // (sub k, channel c) lives at (k * groups + c / 8) * 8 + c % 8
    R.db_partial[l][(long long)blockIdx.x * C + c] = t;
  }
}
// db[c] = sum_r partial[r][c]: 32 channels per CTA (lane = channel), 8 warps split the partial rows
__global__ void __launch_bounds__(256) fcd_colsum_final_kernel(LayerReduce R) {
  __shared__ float red[8][32];
  const int l = blockIdx.y;
  const int C = R.C[l];
  if ((int)blockIdx.x * 32 >= C) return;
  const int rows = R.ctas[l];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (c < C)
    for (int r = warp; r < rows; r += 8) acc += R.db_partial[l][(long long)r * C + c];
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][lane];
    R.db[l][c] = t;
  }
}

// dW_l[co][ci][kh][kw] = sum_z part[z][tap][co][col].  One CTA per (co, chunk of 64 columns): the partials are read as
// 256-byte row segments (col fastest), summed over the splits, transposed through shared memory and written as one
// contiguous run of dW (tap fastest) -- both sides coalesced.
__global__ void __launch_bounds__(256) fcd_wgrad_reduce_kernel(LayerReduce R) {
  if (blockIdx.y == 4) {              // classifier: thread = (tap, channel), slices summed in order
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= 16 * R.cls_C) return;
    const int t = i / R.cls_C, c = i - t * R.cls_C;
    float acc = 0.f;
#pragma unroll 8
    for (int s = 0; s < CLS_SLICES; ++s) acc += __ldg(R.cls_part + (int64_t)s * 16 * R.cls_C + i);
    R.cls_dw[(int64_t)c * 16 + t] = acc;
    return;
  }
  const int l = blockIdx.y;           // 0..3 <-> conv1..conv4
  const int Cout = R.Cout[l], Cin_real = R.Cin_real[l], Ncols = R.Ncols[l], S = R.S[l];
  const int taps = l == 0 ? 8 : 16;
  const int chunks = (Ncols + 63) / 64;
  if ((int)blockIdx.x >= Cout * chunks) return;
  const int co = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  __shared__ float t[16][65];
  const float* part = R.part[l];
  for (int e = threadIdx.x; e < taps * 64; e += 256) {
    const int tap = e >> 6, cl = e & 63;
    const int col = chunk * 64 + cl;
    float acc = 0.f;
    if (col < Ncols) {
      const float* src = part + ((long long)tap * Cout + co) * Ncols + col;
      const long long zs = (long long)taps * Cout * Ncols;  // elements between split-K partials
#pragma unroll 8
      for (int s = 0; s < S; ++s) acc += __ldg(src + s * zs);
    }
    t[tap][cl] = acc;
  }
  __syncthreads();
  if (l == 0) {
    // conv1: a k-step holds two horizontally adjacent taps x 32 (padded) channels: tap = kh*2 + kw/2, col = (kw%2)*32 + ci
    for (int e = threadIdx.x; e < Cin_real * 16; e += 256) {
      const int ci = e >> 4, kh = (e >> 2) & 3, kw = e & 3;
      R.dw[l][((long long)co * Cin_real + ci) * 16 + (e & 15)] = t[kh * 2 + (kw >> 1)][(kw & 1) * 32 + ci];
    }
  } else {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int cl = e >> 4, k = e & 15;
      const int ci = chunk * 64 + cl;
      if (ci < Cin_real) R.dw[l][((long long)co * Cin_real + ci) * 16 + k] = t[k][cl];
    }
  }
}

// Round-2 tail of the backward pass: ONE launch with a flat block table (no empty CTAs) that (a) sums the split-K partials of
// the four weight gradients with 16-byte loads -- a CTA owns (co, 64 columns): thread = (tap, four columns), all S splits of
// its float4 in flight at once, summed in split order, transposed through shared memory into one contiguous 4 KB run of dW --
// (b) sums the classifier's slices, (c) finishes the four bias gradients from their per-CTA partials and (d) sums dout for the
// classifier's bias gradient.  Same summation orders as the separate kernels above (bit-identical results).
__global__ void __launch_bounds__(256) fcd_reduce_tail_kernel(const LayerReduce R) {
  int seg = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) seg += (int)blockIdx.x >= R.seg_end[i] ? 1 : 0;
  int blk = blockIdx.x;
#pragma unroll
  for (int i = 0; i < 9; ++i) blk = (seg == i + 1) ? (int)blockIdx.x - R.seg_end[i] : blk;
  // the short latency-bound segments come first so that they run beside the bandwidth-bound ones instead of after them
  if (seg == 4) {
    block_channel_sum(R.cls_dout, R.cls_db, 1, 1, R.cls_n, 0);
    return;
  }
  if (seg < 4) {                       // bias gradient of conv layer l: 32 channels per CTA, 8 warps split the partial rows
    __shared__ float red[8][32];
    const int l = seg;
    const int C = R.C[l], rows = R.ctas[l];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blk * 32 + lane;
    float acc = 0.f;
    if (c < C) {
#pragma unroll 8
      for (int r = warp; r < rows; r += 8) acc += R.db_partial[l][(long long)r * C + c];
    }
    red[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && c < C) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[k][lane];
      R.db[l][c] = t;
    }
    return;
  }
  if (seg == 9) {                      // classifier: thread = (tap, channel), slices summed in order
    const int i = blk * 256 + threadIdx.x;
    if (i >= 16 * R.cls_C) return;
    const int t = i / R.cls_C, c = i - t * R.cls_C;
    float acc = 0.f;
#pragma unroll 16
    for (int s = 0; s < CLS_SLICES; ++s) acc += __ldg(R.cls_part + (int64_t)s * 16 * R.cls_C + i);
    R.cls_dw[(int64_t)c * 16 + t] = acc;
    return;
  }
  const int l = seg - 5;               // 0..3 <-> conv1..conv4
  const int Cout = R.Cout[l], Cin_real = R.Cin_real[l], Ncols = R.Ncols[l], S = R.S[l];
  const int taps = l == 0 ? 8 : 16;
  const int chunks = (Ncols + 63) / 64;
  const int co = blk / chunks, chunk = blk - co * chunks;
  __shared__ float t[16][65];
  if ((int)threadIdx.x < taps * 16) {
    const int tap = threadIdx.x >> 4, c4 = (threadIdx.x & 15) * 4;
    const int col = chunk * 64 + c4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < Ncols) {                 // Ncols is a multiple of 4 (64 .. 256)
      const float4* src = reinterpret_cast<const float4*>(R.part[l] + ((long long)tap * Cout + co) * Ncols + col);
      const long long zs4 = (long long)taps * Cout * Ncols / 4;  // float4s between split-K partials
#pragma unroll 8
      for (int s = 0; s < S; ++s) {
        const float4 v = ld_stream(src + s * zs4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    t[tap][c4] = acc.x; t[tap][c4 + 1] = acc.y; t[tap][c4 + 2] = acc.z; t[tap][c4 + 3] = acc.w;
  }
  __syncthreads();
  if (l == 0) {
    for (int e = threadIdx.x; e < Cin_real * 16; e += 256) {
      const int ci = e >> 4, kh = (e >> 2) & 3, kw = e & 3;
      R.dw[l][((long long)co * Cin_real + ci) * 16 + (e & 15)] = t[kh * 2 + (kw >> 1)][(kw & 1) * 32 + ci];
    }
  } else {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int cl = e >> 4, k = e & 15;
      const int ci = chunk * 64 + cl;
      if (ci < Cin_real) R.dw[l][((long long)co * Cin_real + ci) * 16 + k] = t[k][cl];
    }
  }
}

// algorithmic flops of conv l (forward = dgrad = wgrad): 2 * pixels_out * Cout * 16 * Cin_real
static double layer_flops(const FcdPlan& p, int l) {
  const double cin = l == 1 ? p.n_cls : p.C[l - 1];
  return 2.0 * p.N * p.H[l] * p.W[l] * p.C[l] * 16.0 * cin;
}

// algorithmic HBM bytes of conv l: which = 0 -> input A_{l-1} + output-sized tensor of level l (bf16, as stored),
// which = 1 -> A_{l-1} alone
static double layer_bytes(const FcdPlan& p, int l, int which) {
  const double in_b = l == 1 ? 2.0 * p.N * p.H[0] * p.W0p * 32 : 2.0 * p.N * p.H[l - 1] * p.W[l - 1] * p.C[l - 1];
  const double out_b = 2.0 * p.N * p.H[l] * p.W[l] * p.C[l];
  return which == 1 ? in_b : in_b + out_b;
}

// ---- tcgen05 launches ---------------------------------------------------------------------------
// largest BLOCK_N dividing n that still yields enough tiles to occupy most SMs of the persistent grid
static int block_n_for(int n, long long m_tiles = 1 << 30) {
  const int cand[4] = {256, 128, 64, 32};
  int best = 32;
  for (int i = 3; i >= 0; --i)
    if (n % cand[i] == 0) {
      best = cand[i];
      break;
    }
  for (int i = 0; i < 4; ++i) {
    if (n % cand[i]) continue;
    if (m_tiles * (n / cand[i]) * 4 >= 3LL * sm_count() || cand[i] == best) return cand[i];
  }
  return best;
}

// parity view (rh, rw) of an NHWC tensor [N][Hin][Win][C]
static int encode_parity(CUtensorMap* m, const __nv_bfloat16* base, int N, int Hin, int Win, int C, int rh, int rw,
                         int th, int tw) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)((Win - rw + 1) / 2), (uint64_t)((Hin - rh + 1) / 2), (uint64_t)N};
  uint64_t str[3] = {(uint64_t)2 * C * 2, (uint64_t)2 * Win * C * 2, (uint64_t)Hin * Win * C * 2};
  return umma::encode_4d(m, base + ((int64_t)rh * Win + rw) * C, dims, str, tw, th);
}
// pair view (row parity rh) of the packed input [N][H][W0p][32] seen as [N][H][W0p/2][64]
static int encode_pair(CUtensorMap* m, const __nv_bfloat16* base, int N, int H, int W0p, int rh, int th, int tw) {
  uint64_t dims[4] = {64, (uint64_t)(W0p / 2), (uint64_t)((H - rh + 1) / 2), (uint64_t)N};
  uint64_t str[3] = {(uint64_t)64 * 2, (uint64_t)2 * W0p * 32 * 2, (uint64_t)H * W0p * 32 * 2};
  return umma::encode_4d(m, base + (int64_t)rh * W0p * 32, dims, str, tw, th);
}
static int encode_plain(CUtensorMap* m, const __nv_bfloat16* base, int N, int H, int W, int C, int th, int tw) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  return umma::encode_4d(m, base, dims, str, tw, th);
}

namespace halo {  // halo.cu: conv1 forward from shared-memory halo tiles (ASN_HALO=0 falls back to the ring kernel)
int conv2_dgrad(const __nv_bfloat16* dpre, const __nv_bfloat16* wd, const __nv_bfloat16* mask_src, __nv_bfloat16* out,
                int N, int Hin, int Win, int H2, int W2, float mask_slope, double flops, double bytes, cudaStream_t st);
int conv1_dgrad(const __nv_bfloat16* dpre, const __nv_bfloat16* wd, __nv_bfloat16* out, int N, int Hin, int Win, int W0p,
                int H1, int W1, double flops, double bytes, cudaStream_t st);
int conv1_fwd(const __nv_bfloat16* in, const __nv_bfloat16* wf, const float* bias, __nv_bfloat16* out, int N, int H0,
              int W0p, int OH, int OW, float slope, double flops, double bytes, cudaStream_t st);
}

static int conv_fwd(const FcdPlan& p, int l, const __nv_bfloat16* in, const __nv_bfloat16* wf, const float* bias,
                    __nv_bfloat16* out, cudaStream_t st) {
  using namespace umma;
  const int OH = p.H[l], OW = p.W[l], Cout = p.C[l];
  static const bool use_halo = !(getenv("ASN_HALO") != nullptr && getenv("ASN_HALO")[0] == '0');
  if (use_halo && l == 1 && Cout == 64)
    return halo::conv1_fwd(in, wf, bias, out, p.N, p.H[0], p.W0p, OH, OW, FCD_SLOPE, layer_flops(p, l),
                           layer_bytes(p, l, 0) + 2.0 * Cout * 512, st);
  int th = 1, tw = 128;
  pick_tile(OH, OW, 128, &th, &tw);
  const int bn_guess = block_n_for(Cout, (long long)p.N * cdiv(OH, th) * cdiv(OW, tw));
  const int rows = rows_per_cta(MODE_CONV, (long long)p.N * cdiv(OH, th) * cdiv(OW, tw), cdiv(Cout, bn_guess));
  if (rows == 256) pick_tile(OH, OW, 256, &th, &tw);
  CUtensorMap maps[5];
  Params P;
  memset(&P, 0, sizeof(P));
  int rc;
  if (l == 1) {
    for (int r = 0; r < 2; ++r)
      if ((rc = encode_pair(&maps[r], in, p.N, p.H[0], p.W0p, r, th, tw))) return rc;
    maps[2] = maps[3] = maps[0];
    P.c_chunks = 1;
    P.taps = 8;
    for (int kh = 0; kh < 4; ++kh)
      for (int pw = 0; pw < 2; ++pw) {
        P.tap_map[kh * 2 + pw] = tap_r(kh);
        P.tap_dh[kh * 2 + pw] = tap_a(kh);
        P.tap_dw[kh * 2 + pw] = pw;
      }
  } else {
    const int Cin = p.C[l - 1];
    for (int rh = 0; rh < 2; ++rh)
      for (int rw = 0; rw < 2; ++rw)
        if ((rc = encode_parity(&maps[rh * 2 + rw], in, p.N, p.H[l - 1], p.W[l - 1], Cin, rh, rw, th, tw))) return rc;
    P.c_chunks = Cin / 64;
    P.taps = 16;
    for (int kh = 0; kh < 4; ++kh)
      for (int kw = 0; kw < 4; ++kw) {
        P.tap_map[kh * 4 + kw] = tap_r(kh) * 2 + tap_r(kw);
        P.tap_dh[kh * 4 + kw] = tap_a(kh);
        P.tap_dw[kh * 4 + kw] = tap_a(kw);
      }
  }
  const int K = P.taps * P.c_chunks * 64;
  const int bn = bn_guess;
  if ((rc = encode_2d(&maps[4], wf, (uint64_t)K, (uint64_t)Cout, (uint64_t)K * 2,
                      (uint32_t)(bn / (rows == 128 ? cluster_size(MODE_CONV, bn) : 1)))))
    return rc;
  P.M = 0;
  P.N = Cout;
  P.k_steps = P.taps * P.c_chunks;
  P.steps_per_split = P.k_steps;
  P.tw = tw; P.th = th;
  P.tiles_w = cdiv(OW, tw); P.tiles_h = cdiv(OH, th);
  P.b_rows_per_z = 0;
  P.epi = EPI_BF16;
  P.out = out;
  P.ld_out = Cout;
  P.oh_ext[0] = OH; P.ow_ext[0] = OW;
  P.out_h = OH; P.out_w = OW;
  P.sy = P.sx = 1;
  P.bias = bias;
  P.slope = FCD_SLOPE;
  P.mask_src = nullptr;
  P.mask_slope = 1.f;
  dim3 grid(p.N * P.tiles_h * P.tiles_w, cdiv(Cout, bn), 1);
  static const char* names[5] = {"", "fcd_conv1_fwd", "fcd_conv2_fwd", "fcd_conv3_fwd", "fcd_conv4_fwd"};
  // the output tile leaves through TMA stores (ASN_TMA_STORE=0: register -> shared memory -> st.global epilogue)
  static const bool tma_store = !(getenv("ASN_TMA_STORE") != nullptr && getenv("ASN_TMA_STORE")[0] == '0');
  CUtensorMap omaps[4];
  if (tma_store && rows == 128 && Cout % 64 == 0) {
    if ((rc = encode_plain(&omaps[0], out, p.N, OH, OW, Cout, th, tw))) return rc;
    omaps[1] = omaps[2] = omaps[3] = omaps[0];
    P.tma_store = 1;
  }
  return launch(MODE_CONV, bn, maps, P, grid, st, names[l], layer_flops(p, l),
                layer_bytes(p, l, 0) + 2.0 * Cout * K, rows, P.tma_store ? omaps : nullptr);
}

// dIn (= dPre_{l-1} after the LeakyReLU mask, or dA0 for l == 1) from dPre_l
static int conv_dgrad(const FcdPlan& p, int l, const __nv_bfloat16* dpre, const __nv_bfloat16* wd,
                      const __nv_bfloat16* act_in, __nv_bfloat16* din, cudaStream_t st) {
  using namespace umma;
  const int Hin = p.H[l - 1], Win = p.W[l - 1], Hout = p.H[l], Wout = p.W[l], Cout = p.C[l];
  const int rows = l == 1 ? 32 : p.C[l - 1];
  const int eh = (Hin + 1) / 2, ew = (Win + 1) / 2;  // largest parity class
  int th = 1, tw = 128;
  pick_tile(eh, ew, 128, &th, &tw);
  const int bn = block_n_for(rows, 4LL * p.N * cdiv(eh, th) * cdiv(ew, tw));
  const int tile_rows = rows_per_cta(MODE_CONV, (long long)p.N * cdiv(eh, th) * cdiv(ew, tw), 4LL * cdiv(rows, bn));
  if (tile_rows == 256) pick_tile(eh, ew, 256, &th, &tw);
  // conv2 (128 -> 64 channels): halo-tile kernel (halo.cu); ASN_HALO2=0 falls back to the ring kernel.  Default since
  // round 2: the whole GPU suite (124 tests) ran green with it, 0.234 -> 0.187 ms per step.
  static const bool use_halo2 = !(getenv("ASN_HALO2") != nullptr && getenv("ASN_HALO2")[0] == '0');
  if (use_halo2 && l == 2 && rows == 64 && Cout == 128)
    return halo::conv2_dgrad(dpre, wd, act_in, din, p.N, Hin, Win, Hout, Wout, FCD_SLOPE, layer_flops(p, l),
                             layer_bytes(p, l, 0) + layer_bytes(p, l, 1) + 2.0 * 4 * rows * 4 * Cout, st);
  // conv1 (64 -> 19 (32) channels, writes the packed-input layout): halo-tile kernel, ASN_HALO1D=0 for the ring kernel
  static const bool use_halo1d = !(getenv("ASN_HALO1D") != nullptr && getenv("ASN_HALO1D")[0] == '0');
  if (use_halo1d && l == 1 && rows == 32 && Cout == 64)
    return halo::conv1_dgrad(dpre, wd, din, p.N, Hin, Win, p.W0p, Hout, Wout, layer_flops(p, l),
                             layer_bytes(p, l, 0) + 2.0 * 4 * rows * 4 * Cout, st);
  CUtensorMap maps[5];
  int rc;
  if ((rc = encode_plain(&maps[0], dpre, p.N, Hout, Wout, Cout, th, tw))) return rc;
  maps[1] = maps[2] = maps[3] = maps[0];
  const int K = 4 * Cout;
  if ((rc = encode_2d(&maps[4], wd, (uint64_t)K, (uint64_t)4 * rows, (uint64_t)K * 2,
                      (uint32_t)(bn / (tile_rows == 128 ? cluster_size(MODE_CONV, bn) : 1)))))
    return rc;
  Params P;
  memset(&P, 0, sizeof(P));
  P.N = rows;
  P.c_chunks = Cout / 64;
  P.taps = 4;
  P.k_steps = 4 * P.c_chunks;
  P.steps_per_split = P.k_steps;
  P.tw = tw; P.th = th;
  P.tiles_w = cdiv(ew, tw); P.tiles_h = cdiv(eh, th);
  for (int z = 0; z < 4; ++z) {
    const int rh = z / 2, rw = z % 2;
    for (int t = 0; t < 4; ++t) {
      const int thh = t / 2, tww = t % 2;
      // rh = 0: kh in {1,3} -> oh = i, i-1 ; rh = 1: kh in {0,2} -> oh = i+1, i
      P.tap_map[z * 4 + t] = 0;
      P.tap_dh[z * 4 + t] = rh == 0 ? (thh == 0 ? 0 : -1) : (thh == 0 ? 1 : 0);
      P.tap_dw[z * 4 + t] = rw == 0 ? (tww == 0 ? 0 : -1) : (tww == 0 ? 1 : 0);
    }
    P.oh_ext[z] = (Hin - rh + 1) / 2;
    P.ow_ext[z] = (Win - rw + 1) / 2;
    P.oy[z] = rh;
    P.ox[z] = rw + (l == 1 ? 1 : 0);  // packed input: pixel w lives at padded column w + 1
  }
  P.b_rows_per_z = rows;
  // Thin layers (<= 64 input channels), opt-in with ASN_DGRAD4=1: all four parity classes in one CTA (MODE_DGRAD4).
  // The 16 (class, tap) products read only NINE distinct shifts of dPre; each shifted tile is fetched once and feeds
  // every class that uses it (-29 % operand bytes into the SM).  Parity-tested, but measured 2x SLOWER on B200
  // (conv2 dgrad 0.24 -> 0.47 ms/step): the stage grows to 48 KB, only three fit, and these layers run at
  // bytes-in-flight / latency -- see profiles/README.md.  Off by default.
  static const bool d4_on = getenv("ASN_DGRAD4") != nullptr && getenv("ASN_DGRAD4")[0] == '1';
  const bool d4 = d4_on && tile_rows == 128 && rows == bn && (bn == 32 || bn == 64);
  if (d4) {
    for (int sh = 0; sh < 9; ++sh) {
      const int dh = sh / 3 - 1, dw = sh % 3 - 1;
      int n = 0;
      for (int z = 0; z < 4; ++z)
        for (int t = 0; t < 4; ++t)
          if (P.tap_dh[z * 4 + t] == dh && P.tap_dw[z * 4 + t] == dw) {
            P.d4_z[sh][n] = z;
            P.d4_t[sh][n] = t;
            ++n;
          }
      P.d4_n[sh] = n;
    }
    P.k_steps = 9 * P.c_chunks;
    P.steps_per_split = P.k_steps;
  }
  P.epi = EPI_BF16;
  P.out = din;
  P.ld_out = rows;
  P.out_h = Hin;
  P.out_w = l == 1 ? p.W0p : Win;
  P.sy = P.sx = 2;
  P.bias = nullptr;
  P.slope = 1.f;
  P.mask_src = l == 1 ? nullptr : act_in;
  P.mask_slope = FCD_SLOPE;
  dim3 grid(p.N * P.tiles_h * P.tiles_w, cdiv(rows, bn), 4);
  static const char* names[5] = {"", "fcd_conv1_dgrad", "fcd_conv2_dgrad", "fcd_conv3_dgrad", "fcd_conv4_dgrad"};
  // reads dPre_l and the mask source A_{l-1}, writes dIn (same size as A_{l-1})
  const double bytes = layer_bytes(p, l, 0) + (l > 1 ? layer_bytes(p, l, 1) : 0.0) + 2.0 * 4 * rows * K;
  if (d4) {
    dim3 grid4(p.N * P.tiles_h * P.tiles_w, 1, 1);
    return launch(MODE_DGRAD4, bn, maps, P, grid4, st, names[l], layer_flops(p, l), bytes, 4 * 128);
  }
  // data gradients: the TMA-store epilogue is parity-tested but measured 3-5 % slower here than the transpose epilogue (the
  // LeakyReLU-mask loads become per-thread rows) -> opt-in with ASN_TMA_STORE_DGRAD=1; the forward convolutions use it
  static const bool tma_store = getenv("ASN_TMA_STORE_DGRAD") != nullptr && getenv("ASN_TMA_STORE_DGRAD")[0] == '1';
  CUtensorMap omaps[4];
  if (tma_store && tile_rows == 128 && rows % 64 == 0 && l > 1) {
    for (int z = 0; z < 4; ++z)   // output-parity class z writes the stride-2 view of dIn that starts at (rh, rw)
      if ((rc = encode_parity(&omaps[z], din, p.N, Hin, Win, rows, z / 2, z % 2, th, tw))) return rc;
    P.tma_store = 1;
  }
  return launch(MODE_CONV, bn, maps, P, grid, st, names[l], layer_flops(p, l), bytes, tile_rows,
                P.tma_store ? omaps : nullptr);
}

static int conv_wgrad(const FcdPlan& p, int l, const __nv_bfloat16* dpre, const __nv_bfloat16* act_in, float* part,
                      int* S_out, cudaStream_t st) {
  using namespace umma;
  const int Hout = p.H[l], Wout = p.W[l], Cout = p.C[l];
  int th = 1, tw = 64;
  pick_tile(Hout, Wout, 64, &th, &tw);
  CUtensorMap maps[5];
  int rc;
  if ((rc = encode_plain(&maps[0], dpre, p.N, Hout, Wout, Cout, th, tw))) return rc;
  Params P;
  memset(&P, 0, sizeof(P));
  const int nn = wgrad_n(p, l);
  if (l == 1) {
    if ((rc = encode_pair(&maps[4], act_in, p.N, p.H[0], p.W0p, 0, th, tw))) return rc;
    if ((rc = encode_pair(&maps[1], act_in, p.N, p.H[0], p.W0p, 1, th, tw))) return rc;
    maps[2] = maps[3] = maps[1];
    P.taps = 8;
    for (int kh = 0; kh < 4; ++kh)
      for (int pw = 0; pw < 2; ++pw) {
        P.tap_map[kh * 2 + pw] = tap_r(kh);
        P.tap_dh[kh * 2 + pw] = tap_a(kh);
        P.tap_dw[kh * 2 + pw] = pw;
      }
  } else {
    const int Cin = p.C[l - 1];
    CUtensorMap pm[4];
    for (int rh = 0; rh < 2; ++rh)
      for (int rw = 0; rw < 2; ++rw)
        if ((rc = encode_parity(&pm[rh * 2 + rw], act_in, p.N, p.H[l - 1], p.W[l - 1], Cin, rh, rw, th, tw))) return rc;
    maps[4] = pm[0]; maps[1] = pm[1]; maps[2] = pm[2]; maps[3] = pm[3];
    P.taps = 16;
    for (int kh = 0; kh < 4; ++kh)
      for (int kw = 0; kw < 4; ++kw) {
        P.tap_map[kh * 4 + kw] = tap_r(kh) * 2 + tap_r(kw);
        P.tap_dh[kh * 4 + kw] = tap_a(kh);
        P.tap_dw[kh * 4 + kw] = tap_a(kw);
      }
  }
  int bn, nt;
  bool pair;
  wgrad_cfg(Cout, nn, &bn, &nt, &pair);
  P.M = Cout;
  P.N = nn;
  P.tw = tw; P.th = th;
  P.tiles_w = cdiv(Wout, tw); P.tiles_h = cdiv(Hout, th);
  P.k_steps = p.N * P.tiles_h * P.tiles_w;
  P.steps_per_split = cdiv(P.k_steps, p.split[l]);
  const int S = cdiv(P.k_steps, P.steps_per_split);
  P.a_boxes = Cout <= 64 ? 1 : 2;
  P.epi = EPI_F32;
  P.out = part;
  P.ld_out = nn;
  P.tap_stride_out = (long long)Cout * nn;
  P.z_stride_out = (long long)P.taps * Cout * nn;
  P.slope = 1.f;
  dim3 grid(cdiv(Cout, 128) * cdiv(nn, bn), P.taps / nt, S);
  static const char* names[5] = {"", "fcd_conv1_wgrad", "fcd_conv2_wgrad", "fcd_conv3_wgrad", "fcd_conv4_wgrad"};
  if ((rc = launch(MODE_WGRAD, bn, maps, P, grid, st, names[l], layer_flops(p, l),
                   layer_bytes(p, l, 0) + 4.0 * S * P.taps * Cout * nn, nt * 128)))
    return rc;
  *S_out = S;
  return ASN_OK;
}

// merged end-of-backward reductions: bias gradients (2 launches) and split-K weight gradients (1 launch)
static int reduce_all(const FcdPlan& p, LayerReduce& R, cudaStream_t st) {
  int max_ctas = 1, max_c = 1;
  double db_bytes = 0, dw_bytes = 0;
  for (int i = 0; i < 4; ++i) {
    const int l = i + 1;
    const int C = p.C[l];
    const long long P = (long long)p.N * p.H[l] * p.W[l];
    const int tpr = C / 2 < 256 ? C / 2 : 256;
    const int sub = 256 / tpr;
    long long ctas = 2LL * sm_count();
    if (ctas > (P + 63) / 64) ctas = (P + 63) / 64;
    const long long cap = (long long)256 * 2048 / ((long long)sub * C);
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    R.rows_per_cta[i] = (int)((P + ctas - 1) / ctas);
    R.ctas[i] = (int)((P + R.rows_per_cta[i] - 1) / R.rows_per_cta[i]);
    R.P[i] = P;
    R.C[i] = C;
    if (R.ctas[i] > max_ctas) max_ctas = R.ctas[i];
    if (C > max_c) max_c = C;
    db_bytes += 2.0 * P * C;
    dw_bytes += 4.0 * (R.S[i] + 1.0) * (i == 0 ? 8 : 16) * R.Cout[i] * R.Ncols[i];
  }
  {
    prof::Scope ps("fcd_bias_grad_partial", 0, db_bytes, st);
    fcd_colsum_partial_kernel<<<dim3(max_ctas, 4), 256, 0, st>>>(R);
    ASN_LAUNCH_CHECK();
  }
  static const bool merged = !(getenv("ASN_GLUE") != nullptr && getenv("ASN_GLUE")[0] == '0');
  bool aligned = true;
  for (int i = 0; i < 4; ++i) aligned = aligned && (reinterpret_cast<uintptr_t>(R.part[i]) & 15) == 0 && R.Ncols[i] % 4 == 0;
  if (merged && aligned) {
    int end = 0;
    for (int i = 0; i < 4; ++i) { end += cdiv(R.C[i], 32); R.seg_end[i] = end; }
    end += R.cls_dout ? 1 : 0; R.seg_end[4] = end;
    for (int i = 0; i < 4; ++i) { end += R.Cout[i] * cdiv(R.Ncols[i], 64); R.seg_end[5 + i] = end; }
    end += R.cls_part ? cdiv(16 * R.cls_C, 256) : 0; R.seg_end[9] = end;
    prof::Scope ps("fcd_wgrad_reduce", 0, dw_bytes, st);
    fcd_reduce_tail_kernel<<<end, 256, 0, st>>>(R);
    ASN_LAUNCH_CHECK();
    return ASN_OK;
  }
  {
    prof::Scope ps("fcd_bias_grad_final", 0, 0, st);
    fcd_colsum_final_kernel<<<dim3(cdiv(max_c, 32), 4), 256, 0, st>>>(R);
    ASN_LAUNCH_CHECK();
  }
  {
    prof::Scope ps("fcd_wgrad_reduce", 0, dw_bytes, st);
    int blocks = 1;
    for (int i = 0; i < 4; ++i) blocks = max(blocks, R.Cout[i] * cdiv(R.Ncols[i], 64));
    blocks = max(blocks, cdiv(16 * R.cls_C, 256));
    fcd_wgrad_reduce_kernel<<<dim3((unsigned)blocks, R.cls_part ? 5 : 4), 256, 0, st>>>(R);
    ASN_LAUNCH_CHECK();
  }
  if (R.cls_dout) return channel_sum_nchw(R.cls_dout, R.cls_db, 1, 1, R.cls_n, st);
  return ASN_OK;
}

}  // namespace asn

using namespace asn;

extern "C" size_t asn_fcd_wpack_bytes(int n_cls, int ndf) {
  FcdPlan p;
  if (make_plan(p, 1, n_cls, ndf, 64, 64)) return 0;
  return p.wpack_total;
}
extern "C" size_t asn_fcd_acts_bytes(int N, int n_cls, int ndf, int H, int W) {
  FcdPlan p;
  if (make_plan(p, N, n_cls, ndf, H, W)) return 0;
  return p.acts_total;
}
extern "C" size_t asn_fcd_workspace_bytes(int N, int n_cls, int ndf, int H, int W) {
  FcdPlan p;
  if (make_plan(p, N, n_cls, ndf, H, W)) return 0;
  return p.ws_total;
}

extern "C" int asn_fcd_act_layout(int N, int n_cls, int ndf, int H, int W, int64_t* out_host) {
  ASN_CHECK_ARG(out_host, "asn_fcd_act_layout: null pointer");
  FcdPlan p;
  int rc = make_plan(p, N, n_cls, ndf, H, W);
  if (rc) return rc;
  for (int l = 0; l <= 4; ++l) {
    out_host[4 * l + 0] = (int64_t)p.act_off[l];
    out_host[4 * l + 1] = p.H[l];
    out_host[4 * l + 2] = l == 0 ? p.W0p : p.W[l];
    out_host[4 * l + 3] = p.C[l];
  }
  return ASN_OK;
}

extern "C" int asn_fcd_pack_weights(const float* const* params_host, int n_cls, int ndf, void* wpack, void* stream) {
  ASN_CHECK_ARG(params_host && wpack, "asn_fcd_pack_weights: null pointer");
  FcdPlan p;
  int rc = make_plan(p, 1, n_cls, ndf, 64, 64);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* base = static_cast<uint8_t*>(wpack);
  PackArgs a;
  memset(&a, 0, sizeof(a));
  int64_t max_total = 16 * (int64_t)p.C[4];
  double bytes = 0;
  for (int l = 1; l <= 4; ++l) {
    const float* w = params_host[2 * (l - 1)];
    const float* b = params_host[2 * (l - 1) + 1];
    ASN_CHECK_ARG(w && b, "asn_fcd_pack_weights: null parameter %d", l);
    const int cin_real = l == 1 ? n_cls : p.C[l - 1];
    const int rows = l == 1 ? 32 : p.C[l - 1];
    const int64_t total = (int64_t)p.C[l] * (l == 1 ? 512 : 16 * cin_real) + (int64_t)16 * rows * p.C[l] + p.C[l];
    a.w[l - 1] = w;
    a.b[l - 1] = b;
    a.wf[l - 1] = reinterpret_cast<__nv_bfloat16*>(base + p.wf_off[l]);
    a.wd[l - 1] = reinterpret_cast<__nv_bfloat16*>(base + p.wd_off[l]);
    a.bias[l - 1] = reinterpret_cast<float*>(base + p.bias_off[l]);
    a.Cout[l - 1] = p.C[l];
    a.Cin_real[l - 1] = cin_real;
    a.Cin_rows[l - 1] = rows;
    if (total > max_total) max_total = total;
    bytes += 4.0 * p.C[l] * cin_real * 16 + 2.0 * total;
  }
  ASN_CHECK_ARG(params_host[8] && params_host[9], "asn_fcd_pack_weights: null classifier parameter");
  a.w[4] = params_host[8];
  a.b[4] = params_host[9];
  a.wc = reinterpret_cast<float*>(base + p.wc_off);
  a.bc = reinterpret_cast<float*>(base + p.bc_off);
  a.C4 = p.C[4];
  prof::Scope ps("fcd_pack_weights", 0, bytes, st);
  (void)max_total;
  int blocks = 1;
  for (int i = 0; i < 4; ++i) blocks = max(blocks, (a.Cout[i] / PK_CO) * cdiv(a.Cin_rows[i], PK_CI));
  static const bool vec = !(getenv("ASN_GLUE") != nullptr && getenv("ASN_GLUE")[0] == '0');
  bool ok = vec;
  for (int i = 1; i < 4; ++i)   // conv2..4: whole 16-channel blocks, 16-byte aligned rows
    ok = ok && a.Cin_real[i] == a.Cin_rows[i] && a.Cin_rows[i] % PK_CI == 0 && a.Cout[i] % PK_CO == 0 &&
         ((reinterpret_cast<uintptr_t>(a.w[i]) | reinterpret_cast<uintptr_t>(a.wf[i]) | reinterpret_cast<uintptr_t>(a.wd[i])) & 15) == 0;
  if (ok) {
    PackTable tb;
    int end = 0;
    for (int i = 0; i < 4; ++i) { end += (a.Cout[i] / PK_CO) * cdiv(a.Cin_rows[i], PK_CI); tb.seg_end[i] = end; }
    end += cdiv(16 * a.C4, 256); tb.seg_end[4] = end;
    fcd_pack_vec_kernel<<<end, 256, 0, st>>>(a, tb);
  } else {
    fcd_pack_all_kernel<<<dim3((unsigned)blocks, 5), 256, 0, st>>>(a);
  }
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

// x_h x x_w = resolution of x: equal to H x W (x is what the reference hands to the discriminator), or lower -- then x
// holds low-res LOGITS and bilinear upsample + softmax happen inside the input pack (lazy_up.cu).
static int fcd_fwd_impl(const float* x_nchw, int x_is_logits, int x_h, int x_w, const void* wpack, void* acts,
                        float* out, int N, int n_cls, int ndf, int H, int W, void* stream) {
  ASN_CHECK_ARG(x_nchw && wpack && acts && out, "asn_fcd_fwd: null pointer");
  const bool lowres = x_h != H || x_w != W;
  ASN_CHECK_ARG(!lowres || (x_is_logits && lazy::supported(n_cls, x_h, x_w, H, W)),
                "asn_fcd_fwd_lowres: needs logits of %d classes at a resolution <= %dx%d (got %dx%d)", 19, H, W, x_h, x_w);
  FcdPlan p;
  int rc = make_plan(p, N, n_cls, ndf, H, W);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint8_t* wb = static_cast<const uint8_t*>(wpack);
  uint8_t* ab = static_cast<uint8_t*>(acts);
  __nv_bfloat16* A[5];
  for (int l = 0; l <= 4; ++l) A[l] = reinterpret_cast<__nv_bfloat16*>(ab + p.act_off[l]);
  if (lowres) {
    if ((rc = lazy::pack_input(x_nchw, A[0], N, n_cls, x_h, x_w, H, W, p.W0p, st))) return rc;
  } else {
    prof::Scope ps("fcd_pack_input", 0, (double)N * H * W * (4.0 * n_cls + 64.0), st);
    // (4 pixels per thread measured slower on B200: 178 registers halve the occupancy)
    if (false && W % 4 == 0 && (reinterpret_cast<uintptr_t>(x_nchw) & 15) == 0)
      fcd_pack_input_kernel<4><<<full_grid((int64_t)N * H * (W / 4), 128), 128, 0, st>>>(x_nchw, A[0], N, n_cls, H, W,
                                                                                         p.W0p, x_is_logits);
    else
      fcd_pack_input_kernel<1><<<full_grid((int64_t)N * H * W, 128), 128, 0, st>>>(x_nchw, A[0], N, n_cls, H, W,
                                                                                   p.W0p, x_is_logits);
    ASN_LAUNCH_CHECK();
  }
  for (int l = 1; l <= 4; ++l) {
    rc = conv_fwd(p, l, A[l - 1], reinterpret_cast<const __nv_bfloat16*>(wb + p.wf_off[l]),
                  reinterpret_cast<const float*>(wb + p.bias_off[l]), A[l], st);
    if (rc) return rc;
  }
  const int n_out = N * p.H[5] * p.W[5];
  prof::Scope ps("fcd_classifier_fwd", 2.0 * n_out * 16 * p.C[4], 0, st);
  fcd_cls_fwd_kernel<<<n_out, 128, 0, st>>>(
      A[4], reinterpret_cast<const float*>(wb + p.wc_off), reinterpret_cast<const float*>(wb + p.bc_off), out, N, p.H[4],
      p.W[4], p.C[4], p.H[5], p.W[5]);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

extern "C" int asn_fcd_fwd(const float* x_nchw, int x_is_logits, const void* wpack, void* acts, float* out, int N,
                           int n_cls, int ndf, int H, int W, void* workspace, size_t workspace_bytes, void* stream) {
  (void)workspace; (void)workspace_bytes;
  return fcd_fwd_impl(x_nchw, x_is_logits, H, W, wpack, acts, out, N, n_cls, ndf, H, W, stream);
}

extern "C" int asn_fcd_fwd_lowres(const float* z_low, int x_h, int x_w, const void* wpack, void* acts, float* out,
                                  int N, int n_cls, int ndf, int H, int W, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  (void)workspace; (void)workspace_bytes;
  return fcd_fwd_impl(z_low, 1, x_h, x_w, wpack, acts, out, N, n_cls, ndf, H, W, stream);
}

extern "C" size_t asn_fcd_workspace_bytes_lowres(int N, int n_cls, int ndf, int H, int W, int x_h, int x_w) {
  FcdPlan p;
  if (make_plan(p, N, n_cls, ndf, H, W) || !lazy::supported(n_cls, x_h, x_w, H, W)) return 0;
  return p.ws_total + lazy::partial_bytes(N, n_cls, x_h, x_w, H, W);
}

static int fcd_bwd_impl(const float* dout, const float* x_logits, int x_h, int x_w, const void* wpack,
                        const void* acts, float* dx_nchw, float* const* dparams_host, int N, int n_cls, int ndf, int H,
                        int W, void* workspace, size_t workspace_bytes, void* stream) {
  ASN_CHECK_ARG(dout && wpack && acts && workspace, "asn_fcd_bwd: null pointer");
  const bool lowres = x_h != H || x_w != W;
  ASN_CHECK_ARG(!lowres || !dx_nchw || (x_logits && lazy::supported(n_cls, x_h, x_w, H, W)),
                "asn_fcd_bwd_lowres: needs the low-res logits the forward saw");
  FcdPlan p;
  int rc = make_plan(p, N, n_cls, ndf, H, W);
  if (rc) return rc;
  const size_t need = p.ws_total + (lowres && dx_nchw ? lazy::partial_bytes(N, n_cls, x_h, x_w, H, W) : 0);
  if (workspace_bytes < need) {
    set_error("asn_fcd_bwd: workspace %zu < %zu", workspace_bytes, need);
    return ASN_EWORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint8_t* wb = static_cast<const uint8_t*>(wpack);
  const uint8_t* ab = static_cast<const uint8_t*>(acts);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const __nv_bfloat16* A[5];
  for (int l = 0; l <= 4; ++l) A[l] = reinterpret_cast<const __nv_bfloat16*>(ab + p.act_off[l]);
  __nv_bfloat16* dPre[5];
  for (int l = 1; l <= 4; ++l) dPre[l] = reinterpret_cast<__nv_bfloat16*>(ws + p.dpre_off[l]);
  __nv_bfloat16* dA0 = reinterpret_cast<__nv_bfloat16*>(ws + p.da0_off);
  float* dbpart = reinterpret_cast<float*>(ws + p.dbpart_off);
  LayerReduce R;
  memset(&R, 0, sizeof(R));
  const float* wc = reinterpret_cast<const float*>(wb + p.wc_off);

  // classifier
  static const bool cls8 = !(getenv("ASN_GLUE") != nullptr && getenv("ASN_GLUE")[0] == '0');
  {
    prof::Scope ps_cls("fcd_classifier_dgrad", 0, 4.0 * N * p.H[4] * p.W[4] * p.C[4], st);
    if (cls8 && p.C[4] % 8 == 0 && ((reinterpret_cast<uintptr_t>(A[4]) | reinterpret_cast<uintptr_t>(dPre[4]) |
                                      reinterpret_cast<uintptr_t>(wc)) & 15) == 0)
      fcd_cls_dgrad8_kernel<<<full_grid((int64_t)N * p.H[4] * p.W[4] * (p.C[4] / 8), 256), 256, 0, st>>>(
          dout, wc, A[4], dPre[4], N, p.H[4], p.W[4], p.C[4], p.H[5], p.W[5]);
    else
      fcd_cls_dgrad_kernel<<<full_grid((int64_t)N * p.H[4] * p.W[4] * (p.C[4] / 2), 256), 256, 0, st>>>(
          dout, wc, A[4], dPre[4], N, p.H[4], p.W[4], p.C[4], p.H[5], p.W[5]);
    ASN_LAUNCH_CHECK();
  }
  if (dparams_host) {
    ASN_CHECK_ARG(dparams_host[8] && dparams_host[9], "asn_fcd_bwd: null classifier gradient");
    {
      prof::Scope ps("fcd_classifier_wgrad", 2.0 * N * p.H[5] * p.W[5] * 16 * p.C[4], 0, st);
      float* clspart = reinterpret_cast<float*>(ws + p.clspart_off);
      static const bool par = !(getenv("ASN_GLUE") != nullptr && getenv("ASN_GLUE")[0] == '0');
      if (par && p.C[4] % 8 == 0 && p.C[4] / 2 <= 256 && (reinterpret_cast<uintptr_t>(clspart) & 15) == 0)
        fcd_cls_wgrad_par_kernel<<<CLS_SLICES, p.C[4] / 2, 0, st>>>(dout, A[4], clspart, N, p.H[4], p.W[4], p.C[4], p.H[5],
                                                                    p.W[5]);
      else
        fcd_cls_wgrad_kernel<<<dim3(CLS_SLICES, cdiv(p.C[4], 64)), 256, 0, st>>>(dout, A[4], clspart, N, p.H[4], p.W[4],
                                                                                     p.C[4], p.H[5], p.W[5]);
      ASN_LAUNCH_CHECK();
      R.cls_part = clspart;          // summed over the slices by the merged reduce at the end of the backward
      R.cls_dw = dparams_host[8];
      R.cls_C = p.C[4];
    }
    R.cls_dout = dout;               // db = sum(dout): part of the merged tail launch
    R.cls_db = dparams_host[9];
    R.cls_n = N * p.H[5] * p.W[5];
  }
  for (int l = 4; l >= 1; --l) {
    if (dparams_host) {
      float* dw = dparams_host[2 * (l - 1)];
      float* db = dparams_host[2 * (l - 1) + 1];
      ASN_CHECK_ARG(dw && db, "asn_fcd_bwd: null gradient pointer for layer %d", l);
      float* part = reinterpret_cast<float*>(ws + p.part_off[l]);
      if ((rc = conv_wgrad(p, l, dPre[l], A[l - 1], part, &R.S[l - 1], st))) return rc;
      R.part[l - 1] = part;
      R.dw[l - 1] = dw;
      R.Cout[l - 1] = p.C[l];
      R.Cin_real[l - 1] = l == 1 ? n_cls : p.C[l - 1];
      R.Ncols[l - 1] = wgrad_n(p, l);
      R.dpre[l - 1] = dPre[l];
      R.db[l - 1] = db;
      R.db_partial[l - 1] = dbpart + (size_t)(l - 1) * 256 * 2048;
    }
    if (l > 1 || dx_nchw) {
      rc = conv_dgrad(p, l, dPre[l], reinterpret_cast<const __nv_bfloat16*>(wb + p.wd_off[l]), A[l - 1],
                      l == 1 ? dA0 : dPre[l - 1], st);
      if (rc) return rc;
    }
  }
  if (dparams_host && (rc = reduce_all(p, R, st))) return rc;
  if (dx_nchw && lowres) {
    rc = lazy::unpack_dx(dA0, x_logits, dx_nchw, N, n_cls, x_h, x_w, H, W, p.W0p, ws + p.ws_total,
                         workspace_bytes - p.ws_total, st);
    if (rc) return rc;
  } else if (dx_nchw) {
    prof::Scope ps("fcd_unpack_dx", 0, (double)N * H * W * (64.0 + 4.0 * n_cls * (x_logits ? 2 : 1)), st);
    // (2 pixels per thread measured slower on B200: 211 registers)
    if (false && W % 2 == 0 && ((reinterpret_cast<uintptr_t>(dx_nchw) | reinterpret_cast<uintptr_t>(x_logits)) & 7) == 0)
      fcd_unpack_dx_kernel<2><<<full_grid((int64_t)N * H * (W / 2), 128), 128, 0, st>>>(dA0, x_logits, dx_nchw, N, n_cls,
                                                                                        H, W, p.W0p);
    else
      fcd_unpack_dx_kernel<1><<<full_grid((int64_t)N * H * W, 128), 128, 0, st>>>(dA0, x_logits, dx_nchw, N, n_cls, H, W,
                                                                                  p.W0p);
    ASN_LAUNCH_CHECK();
  }
  return ASN_OK;
}

extern "C" int asn_fcd_bwd(const float* dout, const float* x_logits, const void* wpack, const void* acts,
                           float* dx_nchw, float* const* dparams_host, int N, int n_cls, int ndf, int H, int W,
                           void* workspace, size_t workspace_bytes, void* stream) {
  return fcd_bwd_impl(dout, x_logits, H, W, wpack, acts, dx_nchw, dparams_host, N, n_cls, ndf, H, W, workspace,
                      workspace_bytes, stream);
}

extern "C" int asn_fcd_bwd_lowres(const float* dout, const float* z_low, int x_h, int x_w, const void* wpack,
                                  const void* acts, float* dz_low, float* const* dparams_host, int N, int n_cls,
                                  int ndf, int H, int W, void* workspace, size_t workspace_bytes, void* stream) {
  return fcd_bwd_impl(dout, z_low, x_h, x_w, wpack, acts, dz_low, dparams_host, N, n_cls, ndf, H, W, workspace,
                      workspace_bytes, stream);
}
