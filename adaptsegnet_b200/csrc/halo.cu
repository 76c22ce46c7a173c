// conv1 of the discriminator (19 -> 64 channels, k4 s2 p1, + bias + LeakyReLU) with SHARED-MEMORY HALO TILES: the first
// kernel of the round-2 plan; ASN_HALO=0 falls back to the ring kernel.  Measured on B200:
// 0.161 -> 0.096 ms per step (four launches), 1.7 -> 2.9 TB/s of algorithmic traffic.
//
// The ring kernel (umma.cuh, MODE_CONV) fetches one 128-pixel activation box per tap: 8 boxes of 16 KB per output tile,
// every input byte enters the SM 4 times, and the 64 KB of weights are re-read for every tile.  Here
//   * the 64 KB of packed weights are loaded ONCE per CTA and stay resident,
//   * per output tile (7 x 16 pixels) the two row-parity views of the packed input are loaded once as (7+1) x (16+1)
//     halo tiles (136 rows of 128 bytes = two adjacent input pixels x 32 channels, SWIZZLE_128B),
//   * the 8 k-steps (kh, kw-pair) are tcgen05 MMAs whose A descriptor points at a 128-row WINDOW of a halo tile starting
//     at row (a - a_min) * 17 + pw -- exact for any row offset because the 128-byte swizzle is a function of the absolute
//     shared-memory address (tools/probes/halo_probe.cu, profiles/r01_halo_probe.json).  Window row m is output pixel
//     (m / 17, m % 17); rows with m % 17 == 16 or m / 17 == 7 are not outputs and are masked on the way out.
// Per tile 34 KB enter the SM instead of 192 KB.  Warp roles as in umma.cuh: TMA producer, MMA issuer (owns TMEM, two
// accumulators), four epilogue warps (TMEM -> registers -> +bias, LeakyReLU -> bf16 -> shared-memory transpose ->
// 512 contiguous bytes per store instruction).
#include "umma_host.cuh"

#include "../../include/asn_b200.h"

namespace asn {
namespace halo {

using namespace umma;

constexpr int TH = 7, TW = 16, PITCH = TW + 1;        // output tile and halo-tile row pitch (pixels)
constexpr int VIEW_ROWS = (TH + 1) * PITCH;           // 136
constexpr int VIEW_BYTES = VIEW_ROWS * 128;           // 17408 = 17 KB (1024-byte multiple)
constexpr int A_BYTES = 2 * VIEW_BYTES;               // both row-parity views of one tile
constexpr int W_BYTES = 8 * 64 * 128;                 // 8 k-steps x 64 output channels x 64 bf16
constexpr int NBUF = 2;
constexpr int THREADS = 192;
constexpr int STG_PITCH = 128;                        // bytes per staged output row (64 bf16), XOR-swizzled 16-byte chunks

struct Layout {
  static constexpr int A_OFF = 0;
  static constexpr int W_OFF = NBUF * A_BYTES;
  static constexpr int BAR_OFF = W_OFF + W_BYTES;
  static constexpr int BIAS_OFF = BAR_OFF + 128;
  static constexpr int STG_OFF = BIAS_OFF + 256;
  static constexpr int ROW_OFF = STG_OFF + 4 * 32 * STG_PITCH;
  static constexpr int TOTAL = ROW_OFF + 4 * 32 * 8 + 1024;
};

struct Args {
  int N, OH, OW, tiles_h, tiles_w;
  const float* bias;
  float slope;
  __nv_bfloat16* out;  // [N][OH][OW][64]
};

__global__ void __launch_bounds__(THREADS, 1)
conv1_fwd_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                 const __grid_constant__ CUtensorMap map_w, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + Layout::BAR_OFF;
  const uint32_t w_bar = bar;
  auto full_bar = [&](int s) { return bar + 8u * (1 + s); };
  auto empty_bar = [&](int s) { return bar + 8u * (3 + s); };
  auto tfull_bar = [&](int s) { return bar + 8u * (5 + s); };
  auto tempty_bar = [&](int s) { return bar + 8u * (7 + s); };
  const uint32_t holder = bar + 8u * 9;
  volatile uint32_t* holder_ptr = reinterpret_cast<volatile uint32_t*>(gen + Layout::BAR_OFF + 72);
  float* bias_s = reinterpret_cast<float*>(gen + Layout::BIAS_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = a.N * a.tiles_h * a.tiles_w;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_a0);
    prefetch_tmap(&map_a1);
    prefetch_tmap(&map_w);
    mbar_init(w_bar, 1);
    for (int s = 0; s < NBUF; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (threadIdx.x < 64) bias_s[threadIdx.x] = a.bias ? __ldg(a.bias + threadIdx.x) : 0.f;
  if (warp == 1) {
    tmem_alloc<1>(holder, 128);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *holder_ptr;
  const uint32_t sw = base + Layout::W_OFF;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_bar, W_BYTES);  // the packed weights [64][512]: one box of 64 rows per k-step, loaded once
      for (int t = 0; t < 8; ++t) tma_load_2d(&map_w, sw + t * 8192, w_bar, t * 64, 0);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const int tx = tile % a.tiles_w, ty = (tile / a.tiles_w) % a.tiles_h, img = tile / (a.tiles_w * a.tiles_h);
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t sa = base + Layout::A_OFF + s * A_BYTES;
        mbar_arrive_expect_tx(full_bar(s), A_BYTES);
        // row parity 0 (kh = 1, 3): view rows oh0 .. oh0+7; row parity 1 (kh = 0, 2): view rows oh0-1 .. oh0+6
        tma_load_4d(&map_a0, sa, full_bar(s), 0, tx * TW, ty * TH, img);
        tma_load_4d(&map_a1, sa + VIEW_BYTES, full_bar(s), 0, tx * TW, ty * TH - 1, img);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, 64, 0, 0);
      mbar_wait(w_bar, 0);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(tempty_bar(s), ph ^ 1);
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = base + Layout::A_OFF + s * A_BYTES;
        const uint32_t td = tmem + s * 64;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int kh = t >> 1, pw = t & 1;
          const int r = (kh == 0 || kh == 2) ? 1 : 0;            // kh - 1 = 2a + r
          const int av = kh == 0 ? -1 : (kh == 3 ? 1 : 0);
          const int off = (av - (r ? -1 : 0)) * PITCH + pw;       // window start (rows) inside the view's halo tile
          const uint32_t ab = sa + r * VIEW_BYTES + off * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_f16_ss(td, make_smem_desc(ab + k * 32, 16, 1024), make_smem_desc(sw + t * 8192 + k * 32, 16, 1024), idesc,
                       (t > 0 || k > 0) ? 1u : 0u);
        }
        mma_commit(empty_bar(s));   // the halo tiles of this buffer have been read
        mma_commit(tfull_bar(s));   // the accumulator is complete
      }
    }
  } else {
    const int ew = warp - 2, q = warp & 3;  // TMEM lane quarter this warp may read
    uint8_t* stg = gen + Layout::STG_OFF + ew * (32 * STG_PITCH);
    long long* row_tab = reinterpret_cast<long long*>(gen + Layout::ROW_OFF) + ew * 32;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int tx = tile % a.tiles_w, ty = (tile / a.tiles_w) % a.tiles_h, img = tile / (a.tiles_w * a.tiles_h);
      {
        const int m = q * 32 + lane;
        const int dy = m / PITCH, dx = m - dy * PITCH;
        const int oh = ty * TH + dy, ow = tx * TW + dx;
        const bool ok = dy < TH && dx < TW && oh < a.OH && ow < a.OW;
        __syncwarp();
        row_tab[lane] = ok ? (((long long)img * a.OH + oh) * a.OW + ow) * 64 : -1;
      }
      mbar_wait(tfull_bar(s), ph);
      tc_fence_after();
      float v[64];
      const uint32_t ta = tmem + s * 64 + ((uint32_t)(q * 32) << 16);
      tmem_ld32(ta, v);
      tmem_ld32(ta + 32, v + 32);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(s));  // the MMAs of the tile after next may overwrite this accumulator
      // + bias, LeakyReLU, bf16; stage this thread's row (128 bytes) with its 16-byte chunks XOR-swizzled by the row
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float x0 = v[c8 * 8 + 2 * j] + bias_s[c8 * 8 + 2 * j];
          float x1 = v[c8 * 8 + 2 * j + 1] + bias_s[c8 * 8 + 2 * j + 1];
          x0 = x0 > 0.f ? x0 : x0 * a.slope;
          x1 = x1 > 0.f ? x1 : x1 * a.slope;
          __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
          pk[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(stg + lane * STG_PITCH + ((c8 ^ (lane & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      __syncwarp();
      // write out: a warp instruction covers 4 rows x 128 bytes (consecutive output pixels are contiguous in NHWC)
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int row = jj * 4 + (lane >> 3), ch = lane & 7;
        const long long off = row_tab[row];
        if (off >= 0) {
          const uint4 val = *reinterpret_cast<const uint4*>(stg + row * STG_PITCH + ((ch ^ (row & 7)) << 4));
          *reinterpret_cast<uint4*>(a.out + off + ch * 8) = val;
        }
      }
      __syncwarp();  // the staging rows are free for the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<1>(tmem, 128);
  }
}

// in: packed discriminator input A0 [N][H0][W0p][32] bf16; wf: conv1's forward pack [64][512]; out: A1 [N][OH][OW][64]
int conv1_fwd(const __nv_bfloat16* in, const __nv_bfloat16* wf, const float* bias, __nv_bfloat16* out, int N, int H0,
              int W0p, int OH, int OW, float slope, double flops, double bytes, cudaStream_t st) {
  CUtensorMap maps[3];
  int rc;
  for (int r = 0; r < 2; ++r) {
    // row-parity view r of the packed input as rows of 128 bytes = 2 adjacent pixels x 32 channels
    uint64_t dims[4] = {64, (uint64_t)(W0p / 2), (uint64_t)((H0 - r + 1) / 2), (uint64_t)N};
    uint64_t str[3] = {(uint64_t)64 * 2, (uint64_t)2 * W0p * 32 * 2, (uint64_t)H0 * W0p * 32 * 2};
    if ((rc = encode_4d(&maps[r], in + (int64_t)r * W0p * 32, dims, str, PITCH, TH + 1))) return rc;
  }
  if ((rc = encode_2d(&maps[2], wf, 512, 64, 512 * 2, 64))) return rc;
  Args a;
  a.N = N; a.OH = OH; a.OW = OW;
  a.tiles_h = cdiv(OH, TH); a.tiles_w = cdiv(OW, TW);
  a.bias = bias; a.slope = slope; a.out = out;
  static PerDevice configured;
  if (!configured.get()) {
    ASN_CUDA(cudaFuncSetAttribute(conv1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Layout::TOTAL));
    configured.set(1);
  }
  const long long tiles = (long long)N * a.tiles_h * a.tiles_w;
  const int ctas = (int)(tiles < sm_count() ? tiles : sm_count());
  prof::Scope ps("fcd_conv1_fwd", flops, bytes, st);
  conv1_fwd_kernel<<<ctas, THREADS, Layout::TOTAL, st>>>(maps[0], maps[1], maps[2], a);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// conv2 data gradient (dPre2 [H2][W2][128] -> dPre1 [H1][W1][64], x LeakyReLU mask of A1) with the same construction.
// The transposed 4x4/s2 convolution splits into four output-parity classes (rh, rw) of 2x2 taps each; the 16
// (class, tap) products read only the nine shifts (dh, dw) in {-1,0,1}^2 of dPre2.  Per base tile (7 x 16 positions
// (i, j); class (rh, rw) writes pixel (2i+rh, 2j+rw)) ONE (7+2) x (16+2) halo tile per 64-channel chunk is loaded and
// every product is an MMA over the 128-row window starting at row (dh+1)*18 + (dw+1).  A CTA owns the two classes of one
// row parity rh (blockIdx.x & 1): their 2 x 4 taps x 2 chunks of weights (128 KB) stay resident in shared memory, the
// halo chunks stream through a 3-stage ring, the two class accumulators are double buffered in TMEM (256 columns).
// Per base tile 2 x 41 KB enter the SMs instead of 672 KB.
constexpr int D_TH = 7, D_TW = 16, D_PITCH = D_TW + 2;   // 18
constexpr int D_BOX_H = D_TH + 2;                        // 9
constexpr int D_BOX_BYTES = D_BOX_H * D_PITCH * 128;     // 162 rows x 128 B = 20736
constexpr int D_STAGE = 21504;                           // 168 rows: the last window (row 38 + 127 = 165) stays inside
constexpr int D_NSTAGE = 3;
constexpr int D_THREADS = 320;                           // producer, MMA issuer, 8 epilogue warps

struct DLayout {
  static constexpr int A_OFF = 0;
  static constexpr int W_OFF = D_NSTAGE * D_STAGE;       // 64512 = 63 KB
  static constexpr int W_BYTES = 16 * 8192;              // 2 classes x 4 taps x 2 chunks x [64 rows][64 k]
  static constexpr int BAR_OFF = W_OFF + W_BYTES;
  static constexpr int STG_OFF = BAR_OFF + 128;
  static constexpr int ROW_OFF = STG_OFF + 8 * 32 * 128;   // per epilogue warp: 32 rows x 128 bytes (bf16, two pixels x 32 ch)
  static constexpr int TOTAL = ROW_OFF + 8 * 32 * 8 + 1024;
};
static_assert(DLayout::W_OFF % 1024 == 0 && D_STAGE % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte aligned bases");
static_assert(DLayout::TOTAL <= 227 * 1024, "shared memory budget");

struct DArgs {
  int N, Hin, Win, tiles_h, tiles_w;   // Hin x Win = size of dPre1 (= A1); tiles over the base grid ceil(Hin/2) x ceil(Win/2)
  const __nv_bfloat16* mask_src;       // A1 (post-LeakyReLU activation): gradient x (A1 > 0 ? 1 : mask_slope)
  float mask_slope;
  __nv_bfloat16* out;                  // dPre1 [N][Hin][Win][64]
};

__global__ void __launch_bounds__(D_THREADS, 1)
conv2_dgrad_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const DArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + DLayout::BAR_OFF;
  const uint32_t w_bar = bar;
  auto full_bar = [&](int s) { return bar + 8u * (1 + s); };
  auto empty_bar = [&](int s) { return bar + 8u * (4 + s); };
  auto tfull_bar = [&](int s) { return bar + 8u * (7 + s); };
  auto tempty_bar = [&](int s) { return bar + 8u * (9 + s); };
  const uint32_t holder = bar + 8u * 11;
  volatile uint32_t* holder_ptr = reinterpret_cast<volatile uint32_t*>(gen + DLayout::BAR_OFF + 88);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rh = blockIdx.x & 1;                       // the row parity whose two classes (rw = 0, 1) this CTA computes
  const int first_tile = blockIdx.x >> 1, tile_stride = gridDim.x >> 1;
  const int total_tiles = a.N * a.tiles_h * a.tiles_w;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_w);
    mbar_init(w_bar, 1);
    for (int s = 0; s < D_NSTAGE; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc<1>(holder, 256);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *holder_ptr;
  const uint32_t sw = base + DLayout::W_OFF;

  if (warp == 0) {
    if (lane == 0) {
      // resident weights: Wd rows (z*64 + ci), K = t*128 + co; tile ((zi*4 + t)*2 + cc) = [64 ci][64 co of chunk cc]
      mbar_arrive_expect_tx(w_bar, DLayout::W_BYTES);
      for (int zi = 0; zi < 2; ++zi)
        for (int t = 0; t < 4; ++t)
          for (int cc = 0; cc < 2; ++cc)
            tma_load_2d(&map_w, sw + ((zi * 4 + t) * 2 + cc) * 8192, w_bar, t * 128 + cc * 64, (rh * 2 + zi) * 64);
      uint32_t kit = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
        const int tx = tile % a.tiles_w, ty = (tile / a.tiles_w) % a.tiles_h, img = tile / (a.tiles_w * a.tiles_h);
        for (int cc = 0; cc < 2; ++cc, ++kit) {
          const int s = kit % D_NSTAGE;
          const uint32_t ph = (kit / D_NSTAGE) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          mbar_arrive_expect_tx(full_bar(s), D_BOX_BYTES);
          // halo tile of chunk cc: base positions (ty*7 - 1 .. +7, tx*16 - 1 .. +16); outside dPre2 = zero fill
          tma_load_4d(&map_a, base + DLayout::A_OFF + s * D_STAGE, full_bar(s), cc * 64, tx * D_TW - 1, ty * D_TH - 1, img);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, 64, 0, 0);
      mbar_wait(w_bar, 0);
      uint32_t kit = 0, tit = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++tit) {
        const int acc = tit & 1;
        mbar_wait(tempty_bar(acc), ((tit >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int cc = 0; cc < 2; ++cc, ++kit) {
          const int s = kit % D_NSTAGE;
          mbar_wait(full_bar(s), (kit / D_NSTAGE) & 1);
          tc_fence_after();
          const uint32_t sa = base + DLayout::A_OFF + s * D_STAGE;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int thh = t >> 1, tww = t & 1;
            // rh = 0: kh in {1,3} -> dh = 0, -1;  rh = 1: kh in {0,2} -> dh = +1, 0   (same for the columns with rw = zi)
            const int dh = rh == 0 ? (thh == 0 ? 0 : -1) : (thh == 0 ? 1 : 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
              for (int zi = 0; zi < 2; ++zi) {   // alternate the two class accumulators: no back-to-back dependent MMAs
                const int dw = zi == 0 ? (tww == 0 ? 0 : -1) : (tww == 0 ? 1 : 0);
                const uint32_t ab = sa + ((dh + 1) * D_PITCH + (dw + 1)) * 128;
                const uint32_t wb = sw + ((zi * 4 + t) * 2 + cc) * 8192;
                mma_f16_ss(tmem + acc * 128 + zi * 64, make_smem_desc(ab + k * 32, 16, 1024),
                           make_smem_desc(wb + k * 32, 16, 1024), idesc, (cc > 0 || t > 0 || k > 0) ? 1u : 0u);
              }
            }
          }
          mma_commit(empty_bar(s));
        }
        mma_commit(tfull_bar(acc));
      }
    }
  } else {
    // Epilogue.  A thread owns base position (i, j) = accumulator row m of BOTH column classes rw = 0, 1 of this CTA's
    // row parity and one half of the 64 channels (warps 2..5: channels 0..31, warps 6..9: 32..63): pixels (2i+rh, 2j) and
    // (2i+rh, 2j+1) are neighbours in dPre1, so its 2 x 32 values are two 64-byte runs 128 bytes apart.  The LeakyReLU
    // masks of those runs (A1 at the same offsets) are fetched BEFORE the wait for the accumulator -- their latency hides
    // behind the MMAs -- and applied in fp32; the bf16 rows are staged XOR-swizzled in shared memory and leave as
    // 16-byte chunks with consecutive lanes on consecutive chunks (full 32-byte sectors, no partial writes).
    const int ew = warp - 2, q = warp & 3;   // TMEM lane quarter this warp may read
    const int half = ew >> 2;                // channel half
    uint8_t* stg = gen + DLayout::STG_OFF + ew * (32 * 128);
    long long* row_tab = reinterpret_cast<long long*>(gen + DLayout::ROW_OFF) + ew * 32;
    uint32_t tit = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++tit) {
      const int acc = tit & 1;
      const int tx = tile % a.tiles_w, ty = (tile / a.tiles_w) % a.tiles_h, img = tile / (a.tiles_w * a.tiles_h);
      const int m = q * 32 + lane;
      const int dy = m / D_PITCH, dx = m - dy * D_PITCH;
      const int ih = 2 * (ty * D_TH + dy) + rh, iw = 2 * (tx * D_TW + dx);
      const bool ok0 = dy < D_TH && dx < D_TW && ih < a.Hin && iw < a.Win, ok1 = ok0 && iw + 1 < a.Win;
      // element offset of this thread's first run: pixel (ih, iw), channels [32 * half, +32); the second run is + 64
      const long long off = (((long long)img * a.Hin + ih) * a.Win + iw) * 64 + half * 32;
      __syncwarp();
      row_tab[lane] = ok0 ? off * 2 + (ok1 ? 1 : 0) : -1;
      uint4 mk[2][4];
#pragma unroll
      for (int z = 0; z < 2; ++z)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          mk[z][c] = (z == 0 ? ok0 : ok1) ? __ldg(reinterpret_cast<const uint4*>(a.mask_src + off + z * 64) + c)
                                          : make_uint4(0u, 0u, 0u, 0u);
      mbar_wait(tfull_bar(acc), (tit >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int z = 0; z < 2; ++z) {
        float v[32];
        tmem_ld32(tmem + acc * 128 + z * 64 + half * 32 + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t mw[4] = {mk[z][c].x, mk[z][c].y, mk[z][c].z, mk[z][c].w};
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {   // bf16 > 0  <=>  sign bit clear and not zero
            const uint32_t lo = mw[i] & 0xffffu, hi = mw[i] >> 16;
            const float x0 = v[c * 8 + 2 * i] * ((lo != 0u && lo < 0x8000u) ? 1.f : a.mask_slope);
            const float x1 = v[c * 8 + 2 * i + 1] * ((hi != 0u && hi < 0x8000u) ? 1.f : a.mask_slope);
            __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
            pk[i] = *reinterpret_cast<uint32_t*>(&h);
          }
          const int ch = z * 4 + c;
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((ch ^ (lane & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));   // the accumulator set is free for the tile after next
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int row = jj * 4 + (lane >> 3), ch = lane & 7;   // chunks 0..3: pixel iw, chunks 4..7: pixel iw + 1
        const long long e = row_tab[row];
        if (e >= 0 && (ch < 4 || (e & 1))) {
          const uint4 val = *reinterpret_cast<const uint4*>(stg + row * 128 + ((ch ^ (row & 7)) << 4));
          *reinterpret_cast<uint4*>(a.out + (e >> 1) + (ch < 4 ? ch * 8 : 64 + (ch - 4) * 8)) = val;
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<1>(tmem, 256);
  }
}

// dpre: dPre2 [N][H2][W2][128]; wd: conv2's dgrad pack [4*64][512]; mask_src: A1 [N][Hin][Win][64]; out: dPre1
int conv2_dgrad(const __nv_bfloat16* dpre, const __nv_bfloat16* wd, const __nv_bfloat16* mask_src, __nv_bfloat16* out,
                int N, int Hin, int Win, int H2, int W2, float mask_slope, double flops, double bytes, cudaStream_t st) {
  CUtensorMap map_a, map_w;
  int rc;
  uint64_t dims[4] = {128, (uint64_t)W2, (uint64_t)H2, (uint64_t)N};
  uint64_t str[3] = {(uint64_t)128 * 2, (uint64_t)W2 * 128 * 2, (uint64_t)H2 * W2 * 128 * 2};
  if ((rc = encode_4d(&map_a, dpre, dims, str, D_PITCH, D_BOX_H))) return rc;
  if ((rc = encode_2d(&map_w, wd, 512, 4 * 64, 512 * 2, 64))) return rc;
  DArgs a;
  a.N = N; a.Hin = Hin; a.Win = Win;
  a.tiles_h = cdiv(cdiv(Hin, 2), D_TH);
  a.tiles_w = cdiv(cdiv(Win, 2), D_TW);
  a.mask_src = mask_src; a.mask_slope = mask_slope; a.out = out;
  static PerDevice configured;
  if (!configured.get()) {
    ASN_CUDA(cudaFuncSetAttribute(conv2_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DLayout::TOTAL));
    configured.set(1);
  }
  const long long tiles = (long long)N * a.tiles_h * a.tiles_w;
  long long ctas = 2 * tiles < sm_count() ? 2 * tiles : sm_count();
  ctas &= ~1LL;  // pairs of CTAs: row parity 0 / 1
  if (ctas < 2) ctas = 2;
  prof::Scope ps("fcd_conv2_dgrad", flops, bytes, st);
  conv2_dgrad_kernel<<<(unsigned)ctas, D_THREADS, DLayout::TOTAL, st>>>(map_a, map_w, a);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}


// ------------------------------------------------------------------------------------------------------------------
// conv1 data gradient (dPre1 [H1][W1][64] -> dA0 [H0][W0p][32], the packed-input layout: pixel w at padded column w + 1;
// no LeakyReLU mask below the first layer).  Same construction as conv2_dgrad, one size smaller: Cout = 64 is ONE
// 64-channel chunk, so a base tile is a single (7+2) x (16+2) halo box (20 KB); the 16 (class, tap) weight tiles are
// [32 ci][64 co] = 4 KB each, 64 KB for all four parity classes -- resident, and every CTA computes all four classes of
// its base tiles (4 x 32 accumulator columns, double buffered in 256 TMEM columns).  Per base tile 20 KB enter the SM
// instead of 16 boxes x 16 KB + weights.  Epilogue: a thread owns base position (i, j) of one ROW parity rh and reads
// the 64 accumulator columns of classes (rh, 0) and (rh, 1): 2 x 32 channels = the 128 contiguous bytes of pixels
// (2i+rh, 2j) and (2i+rh, 2j+1); rows are staged XOR-swizzled in shared memory and leave as 16-byte chunks with
// consecutive lanes on consecutive chunks.
constexpr int E_THREADS = 320;
struct ELayout {
  static constexpr int A_OFF = 0;
  static constexpr int W_OFF = D_NSTAGE * D_STAGE;       // 63 KB
  static constexpr int W_BYTES = 16 * 4096;              // 4 classes x 4 taps x [32 ci][64 co]
  static constexpr int BAR_OFF = W_OFF + W_BYTES;
  static constexpr int STG_OFF = BAR_OFF + 128;
  static constexpr int ROW_OFF = STG_OFF + 8 * 32 * 128;  // per epilogue warp: 32 rows x 128 bytes
  static constexpr int TOTAL = ROW_OFF + 8 * 32 * 8 + 1024;
};
static_assert(ELayout::W_OFF % 1024 == 0 && ELayout::TOTAL <= 227 * 1024, "conv1 dgrad shared-memory layout");

struct EArgs {
  int N, Hin, Win, W0p, tiles_h, tiles_w;   // Hin x Win: the discriminator's input resolution; W0p: padded row pitch of dA0
  __nv_bfloat16* out;                       // dA0 [N][Hin][W0p][32]
};

__global__ void __launch_bounds__(E_THREADS, 1)
conv1_dgrad_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const EArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + ELayout::BAR_OFF;
  const uint32_t w_bar = bar;
  auto full_bar = [&](int s) { return bar + 8u * (1 + s); };
  auto empty_bar = [&](int s) { return bar + 8u * (4 + s); };
  auto tfull_bar = [&](int s) { return bar + 8u * (7 + s); };
  auto tempty_bar = [&](int s) { return bar + 8u * (9 + s); };
  const uint32_t holder = bar + 8u * 11;
  volatile uint32_t* holder_ptr = reinterpret_cast<volatile uint32_t*>(gen + ELayout::BAR_OFF + 88);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = a.N * a.tiles_h * a.tiles_w;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_w);
    mbar_init(w_bar, 1);
    for (int s = 0; s < D_NSTAGE; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc<1>(holder, 256);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *holder_ptr;
  const uint32_t sw = base + ELayout::W_OFF;

  if (warp == 0) {
    if (lane == 0) {
      // resident weights: Wd rows z*32 + ci, K = t*64 + co; tile (z*4 + t) = [32 ci][64 co]
      mbar_arrive_expect_tx(w_bar, ELayout::W_BYTES);
      for (int z = 0; z < 4; ++z)
        for (int t = 0; t < 4; ++t) tma_load_2d(&map_w, sw + (z * 4 + t) * 4096, w_bar, t * 64, z * 32);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int tx = tile % a.tiles_w, ty = (tile / a.tiles_w) % a.tiles_h, img = tile / (a.tiles_w * a.tiles_h);
        const int s = it % D_NSTAGE;
        mbar_wait(empty_bar(s), ((it / D_NSTAGE) & 1) ^ 1);
        mbar_arrive_expect_tx(full_bar(s), D_BOX_BYTES);
        tma_load_4d(&map_a, base + ELayout::A_OFF + s * D_STAGE, full_bar(s), 0, tx * D_TW - 1, ty * D_TH - 1, img);
      }
    }
  } else if (warp == 1) {
    // the whole warp walks the tile loop (converged); one elected lane issues the MMAs (see elect_one)
    constexpr uint32_t idesc = make_idesc(128, 32, 0, 0);
    mbar_wait(w_bar, 0);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1, s = it % D_NSTAGE;
      mbar_wait(tempty_bar(acc), ((it >> 1) & 1) ^ 1);
      mbar_wait(full_bar(s), (it / D_NSTAGE) & 1);
      tc_fence_after();
      const uint32_t sa = base + ELayout::A_OFF + s * D_STAGE;
      if (elect_one()) {
        // consecutive MMAs go to different accumulators (class innermost)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int thh = t >> 1, tww = t & 1;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int z = 0; z < 4; ++z) {
              const int rh = z >> 1, rw = z & 1;
              const int dh = rh == 0 ? (thh == 0 ? 0 : -1) : (thh == 0 ? 1 : 0);
              const int dw = rw == 0 ? (tww == 0 ? 0 : -1) : (tww == 0 ? 1 : 0);
              const uint32_t ab = sa + ((dh + 1) * D_PITCH + (dw + 1)) * 128;
              const uint32_t wb = sw + (z * 4 + t) * 4096;
              mma_f16_ss(tmem + acc * 128 + z * 32, make_smem_desc(ab + k * 32, 16, 1024),
                         make_smem_desc(wb + k * 32, 16, 1024), idesc, (t > 0 || k > 0) ? 1u : 0u);
            }
          }
        }
        mma_commit(empty_bar(s));
        mma_commit(tfull_bar(acc));
      }
      __syncwarp();
    }
  } else {
    const int ew = warp - 2, q = warp & 3;   // TMEM lane quarter this warp may read
    const int rh = ew >> 2;                  // warps 2..5: row parity 0, warps 6..9: row parity 1
    uint8_t* stg = gen + ELayout::STG_OFF + ew * (32 * 128);
    long long* row_tab = reinterpret_cast<long long*>(gen + ELayout::ROW_OFF) + ew * 32;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int tx = tile % a.tiles_w, ty = (tile / a.tiles_w) % a.tiles_h, img = tile / (a.tiles_w * a.tiles_h);
      {
        const int m = q * 32 + lane;
        const int dy = m / D_PITCH, dx = m - dy * D_PITCH;
        const int ih = 2 * (ty * D_TH + dy) + rh, iw = 2 * (tx * D_TW + dx);
        const bool ok = dy < D_TH && dx < D_TW && ih < a.Hin && iw < a.Win;
        __syncwarp();
        // element offset of pixel (ih, iw) in the packed layout (padded column iw + 1); low bit: pixel iw + 1 exists too
        row_tab[lane] = ok ? ((((long long)img * a.Hin + ih) * a.W0p + iw + 1) * 32) * 2 + (iw + 1 < a.Win ? 1 : 0) : -1;
      }
      mbar_wait(tfull_bar(acc), (it >> 1) & 1);
      tc_fence_after();
      float v[64];
      const uint32_t ta = tmem + acc * 128 + rh * 64 + ((uint32_t)(q * 32) << 16);
      tmem_ld32(ta, v);
      tmem_ld32(ta + 32, v + 32);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162 h = __floats2bfloat162_rn(v[c8 * 8 + 2 * j], v[c8 * 8 + 2 * j + 1]);
          pk[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(stg + lane * 128 + ((c8 ^ (lane & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      __syncwarp();
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int row = jj * 4 + (lane >> 3), ch = lane & 7;   // chunks 0..3: pixel iw, chunks 4..7: pixel iw + 1
        const long long e = row_tab[row];
        if (e >= 0 && (ch < 4 || (e & 1))) {
          const uint4 val = *reinterpret_cast<const uint4*>(stg + row * 128 + ((ch ^ (row & 7)) << 4));
          *reinterpret_cast<uint4*>(a.out + (e >> 1) + ch * 8) = val;
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<1>(tmem, 256);
  }
}

// dpre: dPre1 [N][H1][W1][64]; wd: conv1's dgrad pack [4*32][256]; out: dA0 [N][Hin][W0p][32]
int conv1_dgrad(const __nv_bfloat16* dpre, const __nv_bfloat16* wd, __nv_bfloat16* out, int N, int Hin, int Win, int W0p,
                int H1, int W1, double flops, double bytes, cudaStream_t st) {
  CUtensorMap map_a, map_w;
  int rc;
  uint64_t dims[4] = {64, (uint64_t)W1, (uint64_t)H1, (uint64_t)N};
  uint64_t str[3] = {(uint64_t)64 * 2, (uint64_t)W1 * 64 * 2, (uint64_t)H1 * W1 * 64 * 2};
  if ((rc = encode_4d(&map_a, dpre, dims, str, D_PITCH, D_BOX_H))) return rc;
  if ((rc = encode_2d(&map_w, wd, 256, 4 * 32, 256 * 2, 32))) return rc;
  EArgs a;
  a.N = N; a.Hin = Hin; a.Win = Win; a.W0p = W0p;
  a.tiles_h = cdiv(cdiv(Hin, 2), D_TH);
  a.tiles_w = cdiv(cdiv(Win, 2), D_TW);
  a.out = out;
  static PerDevice configured;
  if (!configured.get()) {
    ASN_CUDA(cudaFuncSetAttribute(conv1_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ELayout::TOTAL));
    configured.set(1);
  }
  const long long tiles = (long long)N * a.tiles_h * a.tiles_w;
  const int ctas = (int)(tiles < sm_count() ? tiles : sm_count());
  prof::Scope ps("fcd_conv1_dgrad", flops, bytes, st);
  conv1_dgrad_kernel<<<ctas, E_THREADS, ELayout::TOTAL, st>>>(map_a, map_w, a);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

}  // namespace halo
}  // namespace asn
