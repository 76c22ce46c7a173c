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
  static bool configured = false;
  if (!configured) {
    ASN_CUDA(cudaFuncSetAttribute(conv1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Layout::TOTAL));
    configured = true;
  }
  const long long tiles = (long long)N * a.tiles_h * a.tiles_w;
  const int ctas = (int)(tiles < sm_count() ? tiles : sm_count());
  prof::Scope ps("fcd_conv1_fwd", flops, bytes, st);
  conv1_fwd_kernel<<<ctas, THREADS, Layout::TOTAL, st>>>(maps[0], maps[1], maps[2], a);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

}  // namespace halo
}  // namespace asn
