// tcgen05 / TMEM / TMA primitives and the one warp-specialised implicit-GEMM kernel every
// tensor-core path of this library runs on (sm_100a only).
//
//   D[128 x BLOCK_N] (fp32, TMEM) += A[128 x 64] (bf16, smem) . B[BLOCK_N x 64]^T (bf16, smem)
//
// per k-step; A/B stages are filled by TMA (SWIZZLE_128B boxes whose inner extent is 64 bf16 =
// 128 bytes), consumed by tcgen05.mma issued from one thread, released by tcgen05.commit.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2.. = epilogue (tcgen05.ld ->
// registers -> shared-memory transpose -> coalesced global stores).  The kernel is PERSISTENT: each CTA walks the
// tile list (tile = blockIdx.x + i * gridDim.x); the shared-memory ring (3-8 stages) keeps streaming across tile
// boundaries and the accumulator is double buffered in TMEM, so the epilogue of tile i overlaps the MMAs of tile
// i+1 and the fixed costs (TMEM allocation, barrier init, descriptor prefetch, pipeline fill) are paid once per SM
// instead of once per tile.  Two launch shapes (EpiCfg): one 320-thread CTA per SM with eight epilogue warps, or two
// 192-thread CTAs per SM with four each; umma.cu picks per (mode, BLOCK_N) from measurements.
// CL = 2 runs the tile on a CTA PAIR (tcgen05 cta_group::2): the two CTAs of a cluster sit on the two SMs of a TPC, each
// loads its own 128 rows of A and HALF of the B tile, and one thread of the leader CTA issues 256 x BLOCK_N MMAs that
// read both shared memories and write both TMEMs -- per output element only (256 + BLOCK_N) / (2 * (128 + BLOCK_N)) of
// the operand bytes enter the SMs, which is what bounds these kernels (profiles/README.md).
// Optional and OFF by default: 256-row CTA tiles with two accumulators per B stage (MT = 2, ASN_MT2=1).
//
// Three operand-fetch programs share the skeleton:
//   GEMM  : A and B are plain row-major [rows][K] matrices (K-major operands).
//   CONV  : A rows are the pixels of a th x tw output box; k-steps walk (tap, 64-channel
//           chunk); each step is ONE 4-D TMA box over an NHWC tensor (or a stride-2 parity
//           view of it) at a tap-dependent offset; out-of-range pixels are zero-filled by
//           TMA, which is the convolution's zero padding.  B = packed weights (K-major).
//   WGRAD : the reduction runs over pixels, so both operands arrive "MN-major": each stage
//           holds 64 pixels x {128 | BLOCK_N} channels as 64-channel TMA boxes.
//   GEMM_MN: the same operand fetch over plain 2-D [K][M] / [K][N] matrices (head weight gradient:
//           X^T . dYcol straight from the NHWC activations, no transposed copies).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace asn {
namespace umma {

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint32_t dst, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint32_t dst, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// TMA store: a SWIZZLE_128B box in shared memory -> global (out-of-range parts of the box are clipped), bulk-group tracked
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {   // at most N of this thread's bulk groups still READ shared memory
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// generic-proxy writes to shared memory (st.shared) become visible to the async proxy (TMA) of this CTA
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// cta_group::2 flavours: the box lands in THIS CTA's shared memory, the bytes are counted on an mbarrier that may live
// in the peer CTA (`bar` is a shared::cluster address, see mapa_rank)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint32_t dst, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* m, uint32_t dst, uint32_t bar, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// true in exactly one lane of a CONVERGED warp.  Issuing the single-thread tcgen05 / TMA instructions under this predicate
// from an otherwise converged warp (instead of inside `if (lane == 0)`) keeps their uniform-register operands out of
// divergent control flow, where ptxas wraps every such instruction in an ELECT / BRA.U.ANY loop (~14 instructions and
// ~70 cycles per MMA: more than a 128 x 32 or 128 x 64 MMA takes to execute).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// CG = 2: executed by the same warp of BOTH CTAs of the pair; allocates the same columns in both TMEMs
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_holder, uint32_t ncols) {
  if (CG == 1)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_holder), "r"(ncols)
                 : "memory");
  else
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_holder), "r"(ncols)
                 : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_relinquish() {
  if (CG == 1)
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  else
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] . B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulate
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// pair MMA: D (rows 0..127 in this CTA's TMEM, 128..255 in the peer's) (+)= A (128 rows from each CTA's shared
// memory) . B (BLOCK_N / 2 rows from each); descriptors are shared-memory OFFSETS valid in both CTAs
__device__ __forceinline__ void mma_f16_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued tcgen05 ops of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// pair flavour: arrives on the mbarrier at this offset in every CTA of `mask` once the pair MMAs have completed
__device__ __forceinline__ void mma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

// 32 lanes x 16 consecutive fp32 columns: thread t receives lane (base_lane + t), columns c..c+15
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------
// descriptors (cute/arch/mma_sm100_desc.hpp bit layout)
// ------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_128B, version 1 (Blackwell)
//   bits [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor for kind::f16: D fp32, A/B bf16, dense, no negate
//   [4,6) c_format=1(F32) | [7,10) a_format=1(BF16) | [10,13) b_format=1 | 15 a_major | 16 b_major
//   [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------
// kernel parameters
// ------------------------------------------------------------------------------------------
enum {
  MODE_GEMM = 0,     // C = A[M][K] . B[N][K]^T, both K-contiguous
  MODE_CONV = 1,
  MODE_WGRAD = 2,
  MODE_GEMM_MN = 3,  // C = A^T . B for A[K][M], B[K][N] (M / N contiguous): 2-D flavour of the WGRAD operand fetch
  MODE_DGRAD4 = 4    // stride-2 transposed conv, all four output-parity classes in one CTA: k-steps walk the NINE
                     // distinct input shifts; each shifted A tile feeds the 1, 2 or 4 classes that use it
};
enum {
  EPI_F32 = 0,   // fp32 store
  EPI_BF16 = 1   // (+bias) (LeakyReLU) (x LeakyReLU mask of `mask_src`) -> bf16 store
};

constexpr int MAX_TAPS = 16;
constexpr int MAX_Z = 4;

struct Params {
  // ---- logical tile grid (walked persistently): tile -> (bx, by, bz) ----
  int grid_x, grid_y, grid_z;
  // ---- problem extents ----
  int M, N;             // GEMM: logical rows / cols of the output (store masks)
  int k_steps;          // total number of 64-wide k-steps
  int steps_per_split;  // k-steps per blockIdx.z slice (GEMM / WGRAD split-K)
  // ---- CONV / WGRAD pixel tiling (th*tw = 128 for CONV, 64 for WGRAD) ----
  int tw, th, tiles_w, tiles_h;  // box and number of boxes per image
  int c_chunks;                  // CONV: 64-channel chunks per tap (k_steps = taps*c_chunks)
  int taps;                      // CONV: taps per z; WGRAD: total taps (blockIdx.y)
  int tap_map[MAX_Z * MAX_TAPS]; // which A (CONV) / B (WGRAD) tensor map a tap reads
  int tap_dh[MAX_Z * MAX_TAPS];  // box origin offset (rows) of the tap
  int tap_dw[MAX_Z * MAX_TAPS];  // box origin offset (cols) of the tap
  int b_rows_per_z;              // CONV: row offset of B per blockIdx.z (dgrad parity classes)
  // DGRAD4: for input shift s = (dh+1)*3 + (dw+1): the classes z that read it and the tap index t of that class
  int d4_n[9], d4_z[9][4], d4_t[9][4];
  int a_boxes;                   // WGRAD: 64-channel boxes of A actually loaded (1 or 2)
  // ---- epilogue ----
  int epi;
  void* out;                     // fp32 or bf16
  long long ld_out;              // elements between consecutive rows (pixels)
  long long z_stride_out;        // GEMM / WGRAD: elements between split-K partials
  long long tap_stride_out;      // WGRAD: elements between per-tap blocks
  int oh_ext[MAX_Z], ow_ext[MAX_Z];  // CONV: valid output extent (tile space) per z
  int out_h, out_w;                  // CONV: full output grid
  int sy, sx;                        // CONV: output row/col stride (dgrad: 2)
  int oy[MAX_Z], ox[MAX_Z];          // CONV: output row/col offset per z (dgrad parity)
  const float* bias;                 // nullable, fp32 [N]
  float slope;                       // LeakyReLU slope applied to the result (1 = none)
  const __nv_bfloat16* mask_src;     // nullable: multiply by (mask_src[off] > 0 ? 1 : mask_slope)
  float mask_slope;
  int tma_store;                     // CONV / EPI_BF16: the tile leaves through TMA stores over map_o[z] (64-channel slabs)
};

// A CTA tile is MT x 128 rows: MT accumulators of BLOCK_N fp32 columns that share every B stage (per byte of shared
// memory filled, MT = 2 does twice the MMA work on B).  Two sets of accumulators (double buffering against the
// epilogue) when 2 * MT * BLOCK_N <= 512 columns, otherwise one.  Power-of-two allocation.
template <int BLOCK_N, int MT>
struct TmemCols {
  static constexpr int per_tile = MT * BLOCK_N;
  static constexpr int nacc = 2 * per_tile <= 512 ? 2 : 1;
  static constexpr int need = nacc * per_tile;
  static constexpr int value = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
};

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // bf16 elements per k-step = one 128-byte swizzle row
constexpr int UMMA_K = 16;
// Two launch shapes (template parameter EW = epilogue warps):
//   EW = 8: one CTA per SM, 320 threads, two epilogue warps per TMEM lane quarter taking alternate 32-column chunks
//           (128 / 64 contiguous bytes per row per store) -- tiles whose epilogue would outlast their MMAs
//   EW = 4: two CTAs per SM, 192 threads each, 16-column chunks -- long-K, small-N tiles (weight gradients), where two
//           independent TMA/MMA pipelines per SM hide the per-k-step latencies better than one
template <int EW>
struct EpiCfg {
  static constexpr int chunk = EW == 8 ? 32 : 16;   // accumulator columns staged per pass
  static constexpr int pitch = chunk + 4;           // floats per staged row (16-byte aligned, conflict-free float4 rows)
  static constexpr int threads = 64 + 32 * EW;      // TMA producer warp + MMA warp + epilogue warps
};

// AT / BT = A / B tiles per stage: K-major modes may stack AT = 2 row tiles on one B tile (MT = 2); the pixel-reduction
// WGRAD mode may feed BT = 2 or 4 taps' B tiles from one A tile.
constexpr int TMA_SLAB_BYTES = BLOCK_M * 128;   // one 64-channel bf16 slab of a 128-row tile, SWIZZLE_128B
constexpr int TMA_SLABS = 3;                    // staging buffers in rotation: one barrier per slab (see the epilogue)
template <int BLOCK_N, int STAGES, int AT = 1, int EW = 8, int CL = 1, int BT = 1, bool TMA_EPI = false>
struct SmemLayout {
  static constexpr int A_BYTES = AT * BLOCK_M * BLOCK_K * 2;  // 16 KB per 128-row sub-tile
  static constexpr int B_BYTES = (BLOCK_N / CL) * BLOCK_K * 2;  // one B tile; a pair keeps half of it in each CTA
  static constexpr int B_TILE = ((B_BYTES + 1023) / 1024) * 1024;
  static constexpr int STAGE_BYTES = A_BYTES + BT * B_TILE;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  // barriers take <= 21 x 8 B; with the TMA-store epilogue the staging area must start on a 1024-byte boundary and holds
  // TMA_SLABS slabs, which the per-warp transpose buffers of the other epilogue overlay
  static constexpr int STG_OFFSET = BAR_OFFSET + (TMA_EPI ? 1024 : 256);
  static constexpr int STG_PLAIN = EW * 32 * EpiCfg<EW>::pitch * 4;       // per-warp [32][pitch] fp32
  static constexpr int STG_BYTES = TMA_EPI && TMA_SLABS * TMA_SLAB_BYTES > STG_PLAIN ? TMA_SLABS * TMA_SLAB_BYTES : STG_PLAIN;
  static constexpr int ROW_OFFSET = STG_OFFSET + STG_BYTES;
  static constexpr int TOTAL = ROW_OFFSET + EW * 32 * 8 + 1024;           // + row tables + alignment slack
};

// the stage layout of umma_kernel<MODE, BLOCK_N, STAGES, CL, MT, EW> (shared with the launcher)
// the TMA-store epilogue exists for single-accumulator convolution tiles made of whole 64-channel slabs
template <int MODE, int BLOCK_N, int MT, int EW>
__host__ __device__ constexpr bool has_tma_epi() { return MODE == MODE_CONV && EW == 8 && MT == 1 && BLOCK_N % 64 == 0; }
template <int MODE, int BLOCK_N, int STAGES, int CL, int MT, int EW>
using KernelSmem = SmemLayout<BLOCK_N, STAGES, (((MODE == MODE_WGRAD && MT > 1) || MODE == MODE_DGRAD4) ? 1 : MT), EW, CL,
                              (((MODE == MODE_WGRAD && MT > 1) || MODE == MODE_DGRAD4) ? MT : 1),
                              has_tma_epi<MODE, BLOCK_N, MT, EW>()>;

struct TileCoord {
  int z, n0, m0, img, oh0, ow0, wg_tap, ks_begin, nsteps;
  bool valid;
};

// CL = 1: tile -> (bx, by, bz).  CL = 2 (CTA pair): `tile` indexes PAIRS of x-neighbours, this CTA takes
// bx = 2 * pair + rank; the second CTA of a pair hanging over the end of an odd grid_x still feeds the pair MMA,
// from out-of-range coordinates -- TMA zero-fills, `valid` masks its stores.
template <int MODE, int BLOCK_N, int CL, int MT>
__device__ __forceinline__ TileCoord decode_tile(const Params& P, int tile, int rank) {
  TileCoord t;
  constexpr int TAPS = MODE == MODE_WGRAD ? MT : 1;  // WGRAD: MT counts the taps a CTA computes from one A tile
  const int gx = CL == 1 ? P.grid_x : (P.grid_x + 1) / 2;
  const int bx = CL == 1 ? tile % gx : 2 * (tile % gx) + rank;
  const int by = (tile / gx) % P.grid_y;
  t.z = tile / (gx * P.grid_y);
  t.valid = bx < P.grid_x;
  t.m0 = t.img = t.oh0 = t.ow0 = t.wg_tap = 0;
  t.ks_begin = 0;
  int ks_end = P.k_steps;
  if (MODE == MODE_GEMM || MODE == MODE_GEMM_MN) {
    t.m0 = bx * (BLOCK_M * MT);
    t.n0 = by * BLOCK_N;
    t.ks_begin = t.z * P.steps_per_split;
    ks_end = min(P.k_steps, t.ks_begin + P.steps_per_split);
  } else if (MODE == MODE_CONV || MODE == MODE_DGRAD4) {
    t.ow0 = (bx % P.tiles_w) * P.tw;
    t.oh0 = ((bx / P.tiles_w) % P.tiles_h) * P.th;
    t.img = bx / (P.tiles_w * P.tiles_h);
    t.n0 = by * BLOCK_N;
  } else {
    // bx = m_tile + m_tiles * n_tile  (m_tiles = ceil(M/128)), by = tap, bz = split
    const int m_tiles = (P.M + BLOCK_M - 1) / BLOCK_M;
    t.m0 = (bx % m_tiles) * BLOCK_M;
    t.n0 = (bx / m_tiles) * BLOCK_N;
    t.wg_tap = by * TAPS;  // first tap of the group
    t.ks_begin = t.z * P.steps_per_split;
    ks_end = min(P.k_steps, t.ks_begin + P.steps_per_split);
  }
  t.nsteps = max(ks_end - t.ks_begin, 0);
  return t;
}

template <int MODE, int BLOCK_N, int STAGES, int CL, int MT, int EW>
__global__ void __launch_bounds__(EpiCfg<EW>::threads, (EW == 4 ? 2 : 1))
umma_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
            const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a3,
            const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_o0,
            const __grid_constant__ CUtensorMap map_o1, const __grid_constant__ CUtensorMap map_o2,
            const __grid_constant__ CUtensorMap map_o3, const __grid_constant__ Params P) {
  constexpr bool MULTI_TAP = MODE == MODE_WGRAD && MT > 1;
  constexpr bool MULTI_B = MULTI_TAP || MODE == MODE_DGRAD4;  // MT accumulators fed by MT B tiles from one A tile
  using L = KernelSmem<MODE, BLOCK_N, STAGES, CL, MT, EW>;
  static_assert(!MULTI_B || CL == 1, "multi-accumulator tiles run on single CTAs");
  static_assert(MODE != MODE_DGRAD4 || MT == 4, "DGRAD4: one accumulator per output-parity class");
  constexpr int EPI_CHUNK = EpiCfg<EW>::chunk, EPI_PITCH = EpiCfg<EW>::pitch;
  constexpr int CPL = EPI_CHUNK / 4;  // columns per lane on the way out (8 or 4)
  static_assert(EW == 8 || 2 * TmemCols<BLOCK_N, MT>::value <= 512, "two CTAs per SM need <= 256 TMEM columns each");
  constexpr int TMEM_COLS = TmemCols<BLOCK_N, MT>::value;
  constexpr int NACC = TmemCols<BLOCK_N, MT>::nacc;
  static_assert(MT == 1 || MODE != MODE_GEMM_MN, "MN-major tiles are 128 rows tall (WGRAD: MT = taps per CTA)");
  static_assert(CL == 1 || (MT == 1 && EW == 8 && BLOCK_N % 16 == 0), "pair mode: one CTA per SM, 256 x BLOCK_N MMAs");
  static_assert(CL == 1 || (MODE != MODE_WGRAD && MODE != MODE_GEMM_MN) || BLOCK_N % 128 == 0,
                "pair mode, MN-major: each CTA holds BLOCK_N / 2 channels of B as 64-channel boxes");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + L::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_holder = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_holder_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_holder - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = CL == 1 ? 0 : (int)cluster_ctarank();
  const int first_tile = CL == 1 ? (int)blockIdx.x : (int)blockIdx.x / CL;   // cluster id
  const int tile_stride = CL == 1 ? (int)gridDim.x : (int)gridDim.x / CL;
  const int total_tiles = (CL == 1 ? P.grid_x : (P.grid_x + 1) / 2) * P.grid_y * P.grid_z;
  constexpr uint16_t CL_MASK = (1u << CL) - 1;
  const bool leader = rank == 0;

  // ---- one-time setup ----
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a0);
    prefetch_tmap(&map_b);
    if (MODE != MODE_GEMM) {
      prefetch_tmap(&map_a1);
      prefetch_tmap(&map_a2);
      prefetch_tmap(&map_a3);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);   // pair: only the leader's is used; it counts the bytes of both CTAs
      mbar_init(empty_bar(s), 1);  // pair: the leader's commit arrives on both CTAs' copies
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), CL * EW);  // one arrival per epilogue warp (of both CTAs, on the leader's copy)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc<CL>(tmem_holder, TMEM_COLS);
    tmem_relinquish<CL>();
  }
  tc_fence_before();
  if (CL == 1) __syncthreads(); else cluster_sync_all();  // the peer's barriers must exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder_ptr;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    // The whole warp walks the loop, converged; ONE elected lane issues the copies (elect_one: single-thread instructions
    // with uniform-register operands cost ~14 extra instructions each when they sit in divergent control flow).
    {
      const CUtensorMap* amaps[4] = {&map_a0, &map_a1, &map_a2, &map_a3};
      uint32_t it = 0;  // k-steps issued so far: the ring keeps streaming across tiles
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
        const TileCoord t = decode_tile<MODE, BLOCK_N, CL, MT>(P, tile, rank);
        for (int i = 0; i < t.nsteps; ++i, ++it) {
          const int ks = t.ks_begin + i;
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t sa = smem_base + s * L::STAGE_BYTES;
          const uint32_t sb = sa + L::A_BYTES;
          if (elect_one()) {
          if (CL == 2) {
            // pair: both CTAs load their own A rows and their half of the B rows into their own shared memory; all
            // bytes are counted on the LEADER's barrier, which the leader arms for the two CTAs together (a peer box
            // landing before the leader has armed the phase only drives the byte count negative for a moment)
            const uint32_t lbar = mapa_rank(full_bar(s), 0);
            if (leader) mbar_arrive_expect_tx(full_bar(s), 2 * (L::A_BYTES + L::B_BYTES));
            const int nb = t.n0 + rank * (BLOCK_N / 2);
            if (MODE == MODE_GEMM_MN) {
              // (the leader armed 2 x (A + B) bytes: both CTAs load P.a_boxes boxes of A -- callers keep M > 64)
              for (int a = 0; a < 2; ++a)
                tma_load_2d_pair(&map_a0, sa + a * 8192, lbar, t.m0 + a * 64, ks * BLOCK_K);
#pragma unroll
              for (int b = 0; b < BLOCK_N / 128; ++b)
                tma_load_2d_pair(&map_b, sb + b * 8192, lbar, nb + b * 64, ks * BLOCK_K);
            } else if (MODE == MODE_WGRAD) {
              const int bx = ks % P.tiles_w, by = (ks / P.tiles_w) % P.tiles_h, im = ks / (P.tiles_w * P.tiles_h);
              for (int a = 0; a < 2; ++a)
                tma_load_4d_pair(&map_a0, sa + a * 8192, lbar, t.m0 + a * 64, bx * P.tw, by * P.th, im);
              const int mi = P.tap_map[t.wg_tap];
              const CUtensorMap* bm = mi == 0 ? &map_b : mi == 1 ? &map_a1 : mi == 2 ? &map_a2 : &map_a3;
#pragma unroll
              for (int b = 0; b < BLOCK_N / 128; ++b)
                tma_load_4d_pair(bm, sb + b * 8192, lbar, nb + b * 64, bx * P.tw + P.tap_dw[t.wg_tap],
                                 by * P.th + P.tap_dh[t.wg_tap], im);
            } else if (MODE == MODE_GEMM) {
              tma_load_2d_pair(&map_a0, sa, lbar, ks * BLOCK_K, t.m0);
              tma_load_2d_pair(&map_b, sb, lbar, ks * BLOCK_K, nb);
            } else {
              const int tap = ks / P.c_chunks, cc = ks - tap * P.c_chunks;
              const int e = t.z * P.taps + tap;
              tma_load_4d_pair(amaps[P.tap_map[e]], sa, lbar, cc * BLOCK_K, t.ow0 + P.tap_dw[e], t.oh0 + P.tap_dh[e],
                               t.img);
              tma_load_2d_pair(&map_b, sb, lbar, ks * BLOCK_K, t.z * P.b_rows_per_z + nb);
            }
          } else if (MODE == MODE_GEMM) {
            mbar_arrive_expect_tx(full_bar(s), L::A_BYTES + L::B_BYTES);
            tma_load_2d(&map_a0, sa, full_bar(s), ks * BLOCK_K, t.m0);
            tma_load_2d(&map_b, sb, full_bar(s), ks * BLOCK_K, t.n0);
          } else if (MODE == MODE_CONV) {
            const int tap = ks / P.c_chunks, cc = ks - tap * P.c_chunks;
            const int e = t.z * P.taps + tap;
            mbar_arrive_expect_tx(full_bar(s), L::A_BYTES + L::B_BYTES);
            tma_load_4d(amaps[P.tap_map[e]], sa, full_bar(s), cc * BLOCK_K, t.ow0 + P.tap_dw[e],
                        t.oh0 + P.tap_dh[e], t.img);
            tma_load_2d(&map_b, sb, full_bar(s), ks * BLOCK_K, t.z * P.b_rows_per_z + t.n0);
          } else if (MODE == MODE_DGRAD4) {
            const int sh = ks / P.c_chunks, cc = ks - sh * P.c_chunks;
            const int nb = P.d4_n[sh];
            mbar_arrive_expect_tx(full_bar(s), L::A_BYTES + nb * L::B_BYTES);
            tma_load_4d(&map_a0, sa, full_bar(s), cc * BLOCK_K, t.ow0 + (sh % 3) - 1, t.oh0 + (sh / 3) - 1, t.img);
            for (int j = 0; j < nb; ++j)
              tma_load_2d(&map_b, sb + j * L::B_TILE, full_bar(s), (P.d4_t[sh][j] * P.c_chunks + cc) * BLOCK_K,
                          P.d4_z[sh][j] * P.b_rows_per_z + t.n0);
          } else if (MODE == MODE_GEMM_MN) {
            // k-step = 64 rows of K: A = columns [m0, m0+128) of A[K][M], B = columns [n0, n0+BLOCK_N) of B[K][N]
            const uint32_t box_bytes = 64 * 64 * 2;
            mbar_arrive_expect_tx(full_bar(s), (P.a_boxes + BLOCK_N / 64) * box_bytes);
            for (int a = 0; a < P.a_boxes; ++a)
              tma_load_2d(&map_a0, sa + a * box_bytes, full_bar(s), t.m0 + a * 64, ks * BLOCK_K);
#pragma unroll
            for (int b = 0; b < BLOCK_N / 64; ++b)
              tma_load_2d(&map_b, sb + b * box_bytes, full_bar(s), t.n0 + b * 64, ks * BLOCK_K);
          } else {
            // k-step = one box of 64 pixels: A = dY channels [m0, m0+128), B = X@tap channels [n0, n0+BLOCK_N)
            const int bx = ks % P.tiles_w, by = (ks / P.tiles_w) % P.tiles_h, im = ks / (P.tiles_w * P.tiles_h);
            const uint32_t box_bytes = 64 * 64 * 2;
            constexpr int TAPS = MULTI_TAP ? MT : 1;  // taps fed from this one A tile
            mbar_arrive_expect_tx(full_bar(s), (P.a_boxes + TAPS * (BLOCK_N / 64)) * box_bytes);
            for (int a = 0; a < P.a_boxes; ++a)
              tma_load_4d(&map_a0, sa + a * box_bytes, full_bar(s), t.m0 + a * 64, bx * P.tw, by * P.th, im);
#pragma unroll
            for (int j = 0; j < TAPS; ++j) {
              const int tap = t.wg_tap + j;
              const int mi = P.tap_map[tap];
              const CUtensorMap* bm = mi == 0 ? &map_b : mi == 1 ? &map_a1 : mi == 2 ? &map_a2 : &map_a3;
#pragma unroll
              for (int b = 0; b < BLOCK_N / 64; ++b)
                tma_load_4d(bm, sb + j * L::B_TILE + b * box_bytes, full_bar(s), t.n0 + b * 64,
                            bx * P.tw + P.tap_dw[tap], by * P.th + P.tap_dh[tap], im);
            }
          }
          }  // elect_one
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (leader) {  // pair: one (elected) thread of the leader CTA's warp drives the tensor cores of both SMs
      constexpr bool MN_MAJOR = MODE == MODE_WGRAD || MODE == MODE_GEMM_MN;
      constexpr uint32_t idesc = MN_MAJOR ? make_idesc(CL * BLOCK_M, BLOCK_N, 1, 1) : make_idesc(CL * BLOCK_M, BLOCK_N, 0, 0);
      uint32_t it = 0, tile_iter = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++tile_iter) {
        const TileCoord t = decode_tile<MODE, BLOCK_N, CL, MT>(P, tile, rank);
        const uint32_t acc = NACC == 2 ? (tile_iter & 1) : 0;
        const uint32_t acc_ph = NACC == 2 ? ((tile_iter >> 1) & 1) : (tile_iter & 1);
        mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1);  // the epilogue has drained this accumulator set
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * (MT * BLOCK_N);
        uint32_t started = 0;  // DGRAD4: classes whose accumulator already holds a partial sum
        for (int i = 0; i < t.nsteps; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * L::STAGE_BYTES;
          const uint32_t sb = sa + L::A_BYTES;
          const bool issuer = elect_one();   // (always the same lane: tcgen05.commit covers the MMAs of ITS thread)
          if constexpr (MODE == MODE_DGRAD4) {
            const int sh = (t.ks_begin + i) / P.c_chunks;
            const int nb = P.d4_n[sh];
            if (issuer) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024);
              for (int j = 0; j < nb; ++j) {
                const int z = P.d4_z[sh][j];
                const uint64_t dbj = make_smem_desc(sb + j * L::B_TILE + k * 32, 16, 1024);
                mma_f16_ss(tmem_d + z * BLOCK_N, da, dbj, idesc, (((started >> z) & 1u) || k > 0) ? 1u : 0u);
              }
            }
            mma_commit(empty_bar(s));
            }
            for (int j = 0; j < nb; ++j) started |= 1u << P.d4_z[sh][j];
          } else if (issuer) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              uint64_t da, db;
              if (MN_MAJOR) {
                // MN-major, SW128: 64-channel chunks LBO = 64 px * 128 B apart, 8-pixel groups SBO = 1 KB apart;
                // advancing K by 16 pixels = 2 KB
                da = make_smem_desc(sa + k * 2048, 8192, 1024);
                db = make_smem_desc(sb + k * 2048, 8192, 1024);
              } else {
                // K-major, SW128: 8-row groups SBO = 1 KB apart; advancing K by 16 elements = 32 B inside the row
                da = make_smem_desc(sa + k * 32, 16, 1024);
                db = make_smem_desc(sb + k * 32, 16, 1024);
              }
              if (CL == 2) mma_f16_ss_pair(tmem_d, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
              else mma_f16_ss(tmem_d, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
              if (MULTI_TAP) {  // the other taps of the group: same A (dY), their own B (X at the tap) and accumulator
#pragma unroll
                for (int j = 1; j < MT; ++j) {
                  const uint64_t dbj = make_smem_desc(sb + j * L::B_TILE + k * 2048, 8192, 1024);
                  mma_f16_ss(tmem_d + j * BLOCK_N, da, dbj, idesc, (i > 0 || k > 0) ? 1u : 0u);
                }
              } else if (MT == 2) {  // rows 128..255 of the tile: next 16 KB of the A stage, same B
                const uint64_t da1 = make_smem_desc(sa + BLOCK_M * BLOCK_K * 2 + k * 32, 16, 1024);
                mma_f16_ss(tmem_d + BLOCK_N, da1, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
              }
            }
            // frees the stage (in both CTAs of a pair) once these MMAs have read it
            if (CL == 1) mma_commit(empty_bar(s)); else mma_commit_pair(empty_bar(s), CL_MASK);
          }
          __syncwarp();
        }
        // accumulator of this tile complete (pair: both halves, each CTA's epilogue waits on its own copy)
        if (elect_one()) {
          if (CL == 1) mma_commit(tmem_full_bar(acc)); else mma_commit_pair(tmem_full_bar(acc), CL_MASK);
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== epilogue ===============================
    // TMEM -> registers (thread = accumulator row) -> a per-warp shared-memory transpose buffer -> global.
    // Going through shared memory turns the natural "one thread owns one row" ownership into stores where the four
    // lanes of a quad write 128 (fp32) / 64 (bf16) contiguous bytes of one row and a warp instruction covers eight
    // complete row segments, instead of 32 scattered 16-byte pieces.  Bias, LeakyReLU and the LeakyReLU-mask multiply
    // are applied on the way out (mask reads are coalesced the same way).  Eight warps: two per TMEM lane quarter,
    // taking alternate 32-column chunks, so that short-K tiles (whose epilogue outlasts their MMAs) drain faster.
    const int ew = warp - 2;            // 0..EW-1
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int half = ew >> 2;           // which of the (EW / 4) warps of that quarter
    float* stg = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + L::STG_OFFSET) +
                 ew * (32 * EPI_PITCH);
    long long* row_tab = reinterpret_cast<long long*>(smem_raw + (smem_base - smem_u32(smem_raw)) + L::ROW_OFFSET) +
                         ew * 32;  // element offset of each of this warp's 32 rows, -1 = masked row
    const int rr = lane >> 2, cq = lane & 3;  // write-out role: row (of 8) and group of CPL columns
    uint32_t tile_iter = 0;
    uint32_t slab_iter = 0;                   // TMA-store epilogue: slabs staged so far (buffer = slab_iter % TMA_SLABS)
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++tile_iter) {
      const TileCoord t = decode_tile<MODE, BLOCK_N, CL, MT>(P, tile, rank);
      const uint32_t acc = NACC == 2 ? (tile_iter & 1) : 0;
      const uint32_t acc_ph = NACC == 2 ? ((tile_iter >> 1) & 1) : (tile_iter & 1);
      if constexpr (has_tma_epi<MODE, BLOCK_N, MT, EW>()) {
        if (P.tma_store) {
          // ---- TMA-store epilogue ----
          // The tile's 128 rows are the th x tw pixel box in row-major order, i.e. exactly the row order of a
          // {64 ch, tw, th, 1} TMA box over the NHWC output (or its stride-2 parity view for the data gradients): thread
          // = row writes its 64 channels of a slab as one 128-byte SWIZZLE_128B row (16-byte chunk c of row r at
          // chunk c ^ (r & 7): conflict-free), the two warps of a lane quarter taking 32 channels each; one barrier,
          // then ONE elected thread stores the slab with cp.async.bulk.tensor -- full 128-byte lines, rows and columns
          // beyond the image clipped by the TMA unit (no row tables, no per-lane bounds).  Three staging buffers rotate:
          // the issuing thread waits until at most one store still reads shared memory right after issuing, so the
          // buffer reused two slabs later is known to be free by everyone who passes the next barrier.
          const CUtensorMap* omap = t.z == 0 ? &map_o0 : t.z == 1 ? &map_o1 : t.z == 2 ? &map_o2 : &map_o3;
          const int r = q * 32 + lane;
          const int dy = r / P.tw, dx = r - dy * P.tw;
          const int oh = t.oh0 + dy, ow = t.ow0 + dx;
          const bool row_ok = t.valid && oh < P.oh_ext[t.z] && ow < P.ow_ext[t.z];
          const long long row_off = (((long long)t.img * P.out_h + (oh * P.sy + P.oy[t.z])) * P.out_w +
                                     (ow * P.sx + P.ox[t.z])) * P.ld_out;   // LeakyReLU-mask source of this row
          uint8_t* stage_base = smem_raw + (smem_base - smem_u32(smem_raw)) + L::STG_OFFSET;
          mbar_wait(tmem_full_bar(acc), acc_ph);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * (MT * BLOCK_N);
#pragma unroll 1
          for (int sl = 0; sl < BLOCK_N / 64; ++sl, ++slab_iter) {
            const int c0 = sl * 64 + half * 32;       // first of this warp's 32 accumulator columns
            const int n = t.n0 + c0;
            uint4 mk[4];
            if (P.mask_src) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                mk[c] = row_ok ? __ldg(reinterpret_cast<const uint4*>(P.mask_src + row_off + n) + c) : make_uint4(0u, 0u, 0u, 0u);
            }
            float v[32];
            if (t.nsteps > 0) {
              tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            uint8_t* buf = stage_base + (slab_iter % TMA_SLABS) * TMA_SLAB_BYTES;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint32_t mw[4] = {mk[c].x, mk[c].y, mk[c].z, mk[c].w};
              uint32_t pk[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float x0 = v[c * 8 + 2 * i], x1 = v[c * 8 + 2 * i + 1];
                if (P.bias) { x0 += __ldg(P.bias + n + c * 8 + 2 * i); x1 += __ldg(P.bias + n + c * 8 + 2 * i + 1); }
                if (P.slope != 1.f) { x0 = x0 > 0.f ? x0 : x0 * P.slope; x1 = x1 > 0.f ? x1 : x1 * P.slope; }
                if (P.mask_src) {   // bf16 > 0  <=>  sign bit clear and not zero
                  const uint32_t lo = mw[i] & 0xffffu, hi = mw[i] >> 16;
                  x0 *= (lo != 0u && lo < 0x8000u) ? 1.f : P.mask_slope;
                  x1 *= (hi != 0u && hi < 0x8000u) ? 1.f : P.mask_slope;
                }
                __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
                pk[i] = *reinterpret_cast<uint32_t*>(&h);
              }
              const int ch = half * 4 + c;   // 16-byte chunk of the 128-byte row
              *reinterpret_cast<uint4*>(buf + r * 128 + ((ch ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            fence_async_smem();
            named_bar_sync(1, EW * 32);      // all 128 rows x 64 channels of the slab are staged
            if (ew == 0 && elect_one()) {
              if (t.valid) tma_store_4d(omap, smem_base + L::STG_OFFSET + (slab_iter % TMA_SLABS) * TMA_SLAB_BYTES,
                                        t.n0 + sl * 64, t.ow0, t.oh0, t.img);
              bulk_commit();
              bulk_wait_read<1>();           // the store issued one slab ago has finished reading its buffer
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CL == 1) mbar_arrive(tmem_empty_bar(acc)); else mbar_arrive_cluster(mapa_rank(tmem_empty_bar(acc), 0));
          }
          continue;
        }
      }
      mbar_wait(tmem_full_bar(acc), acc_ph);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < MT; ++sub) {
        const uint32_t tmem_d = tmem_base + acc * (MT * BLOCK_N) + sub * BLOCK_N;
        {  // row table: thread = row
          const int r = (MULTI_B ? 0 : sub * BLOCK_M) + q * 32 + lane;
          bool row_ok;
          long long row_off;
          if (MODE == MODE_GEMM || MODE == MODE_GEMM_MN) {
            row_ok = t.valid && (t.m0 + r) < P.M;
            row_off = (long long)t.z * P.z_stride_out + (long long)(t.m0 + r) * P.ld_out;
          } else if (MODE == MODE_CONV || MODE == MODE_DGRAD4) {
            const int zc = MODE == MODE_DGRAD4 ? sub : t.z;  // output-parity class of this accumulator
            const int dy = r / P.tw, dx = r - dy * P.tw;
            const int oh = t.oh0 + dy, ow = t.ow0 + dx;
            row_ok = t.valid && oh < P.oh_ext[zc] && ow < P.ow_ext[zc];
            const int fh = oh * P.sy + P.oy[zc], fw = ow * P.sx + P.ox[zc];
            row_off = (((long long)t.img * P.out_h + fh) * P.out_w + fw) * P.ld_out;
          } else {
            row_ok = (t.m0 + r) < P.M;
            row_off = (long long)t.z * P.z_stride_out + (long long)(t.wg_tap + sub) * P.tap_stride_out +
                      (long long)(t.m0 + r) * P.ld_out;
          }
          __syncwarp();
          row_tab[lane] = row_ok ? row_off : -1;
        }
#pragma unroll 1
        for (int c = half * EPI_CHUNK; c < BLOCK_N; c += (EW / 4) * EPI_CHUNK) {
          float v[EPI_CHUNK];
          __syncwarp();  // previous chunk's reads of stg are done; also reconverges for the .aligned tcgen05.ld
          if (t.nsteps > 0) {
            const uint32_t ta = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
            if (EPI_CHUNK == 32) tmem_ld32(ta, v); else tmem_ld16(ta, v);  // may run past BLOCK_N: masked below
          } else {
#pragma unroll
            for (int i = 0; i < EPI_CHUNK; ++i) v[i] = 0.f;
          }
          float4* srow = reinterpret_cast<float4*>(stg + lane * EPI_PITCH);
#pragma unroll
          for (int i = 0; i < EPI_CHUNK / 4; ++i)
            srow[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          __syncwarp();
          const int cl = c + cq * CPL;      // first of this lane's CPL columns inside the tile
          const int n = t.n0 + cl;
          if (cl >= BLOCK_N || n >= P.N) continue;
          const int n_ok = min(min(BLOCK_N - cl, P.N - n), CPL);  // valid columns of this lane
          const bool full = n_ok == CPL;
          float bv[CPL];
#pragma unroll
          for (int i = 0; i < CPL; ++i) bv[i] = 0.f;
          if (P.epi == EPI_BF16 && P.bias) {
#pragma unroll
            for (int i = 0; i < CPL; ++i) bv[i] = i < n_ok ? __ldg(P.bias + n + i) : 0.f;
          }
          // LeakyReLU masks of this lane's four rows, fetched up front: issued back to back, their (HBM) latencies
          // overlap instead of adding up between the stores of the row loop below
          uint32_t mpre[4][CPL / 2];
          if (P.epi == EPI_BF16 && P.mask_src) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const long long off = row_tab[j * 8 + rr];
#pragma unroll
              for (int i = 0; i < CPL / 2; ++i) mpre[j][i] = 0u;
              if (off < 0) continue;
              const __nv_bfloat16* ms = P.mask_src + off + n;
              const bool vec = full && (((reinterpret_cast<uintptr_t>(P.out) + 2 * (off + n)) & (2 * CPL - 1)) == 0);
              if (!vec) continue;
              if (CPL == 8) {
                const uint4 m = __ldg(reinterpret_cast<const uint4*>(ms));
                mpre[j][0] = m.x; mpre[j][1] = m.y; mpre[j][CPL / 2 - 2] = m.z; mpre[j][CPL / 2 - 1] = m.w;
              } else {
                const uint2 m = __ldg(reinterpret_cast<const uint2*>(ms));
                mpre[j][0] = m.x; mpre[j][1] = m.y;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rl = j * 8 + rr;
            const long long off = row_tab[rl];
            if (off < 0) continue;
            float o[CPL];
#pragma unroll
            for (int g = 0; g < CPL / 4; ++g) {
              const float4 x = *reinterpret_cast<const float4*>(stg + rl * EPI_PITCH + cq * CPL + 4 * g);
              o[4 * g] = x.x; o[4 * g + 1] = x.y; o[4 * g + 2] = x.z; o[4 * g + 3] = x.w;
            }
            if (P.epi == EPI_F32) {
              float* dst = reinterpret_cast<float*>(P.out) + off + n;
              if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                for (int g = 0; g < CPL / 4; ++g)
                  reinterpret_cast<float4*>(dst)[g] = make_float4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
              } else {
                for (int i = 0; i < n_ok; ++i) dst[i] = o[i];
              }
            } else {
#pragma unroll
              for (int i = 0; i < CPL; ++i) {
                o[i] += bv[i];
                if (P.slope != 1.f) o[i] = o[i] > 0.f ? o[i] : o[i] * P.slope;
              }
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(P.out) + off + n;
              const bool vec = full && ((reinterpret_cast<uintptr_t>(dst) & (2 * CPL - 1)) == 0);
              uint32_t mw[CPL / 2];
              if (P.mask_src) {
                const __nv_bfloat16* ms = P.mask_src + off + n;
                if (vec) {
#pragma unroll
                  for (int i = 0; i < CPL / 2; ++i) mw[i] = mpre[j][i];
#pragma unroll
                  for (int i = 0; i < CPL / 2; ++i) {  // bf16 > 0  <=>  sign bit clear and not zero
                    const uint32_t lo = mw[i] & 0xffffu, hi = mw[i] >> 16;
                    o[2 * i] *= (lo != 0u && lo < 0x8000u) ? 1.f : P.mask_slope;
                    o[2 * i + 1] *= (hi != 0u && hi < 0x8000u) ? 1.f : P.mask_slope;
                  }
                } else {
                  for (int i = 0; i < n_ok; ++i) o[i] *= (__bfloat162float(ms[i]) > 0.f ? 1.f : P.mask_slope);
                }
              }
              if (vec) {
                uint32_t pk[CPL / 2];
#pragma unroll
                for (int i = 0; i < CPL / 2; ++i) {
                  __nv_bfloat162 h = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
                  pk[i] = *reinterpret_cast<uint32_t*>(&h);
                }
                if (CPL == 8)
                  *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[CPL / 2 - 2], pk[CPL / 2 - 1]);
                else
                  *reinterpret_cast<uint2*>(dst) = make_uint2(pk[0], pk[1]);
              } else {
                for (int i = 0; i < n_ok; ++i) dst[i] = __float2bfloat16(o[i]);
              }
            }
          }
        }
      }  // sub
      // this thread's TMEM reads of the tile are complete (tcgen05.wait::ld in tmem_ld32): release the buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CL == 1) mbar_arrive(tmem_empty_bar(acc)); else mbar_arrive_cluster(mapa_rank(tmem_empty_bar(acc), 0));
      }
    }
  }

  // ---- teardown (a CTA may not exit while its peer can still read its shared memory or signal its barriers) ----
  if (has_tma_epi<MODE, BLOCK_N, MT, EW>() && warp == 2) bulk_wait_read<0>();   // (the issuing lane's groups; no-op elsewhere)
  tc_fence_before();
  if (CL == 1) __syncthreads(); else cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<CL>(tmem_base, TMEM_COLS);
  }
}

}  // namespace umma
}  // namespace asn
