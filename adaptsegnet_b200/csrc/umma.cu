// Host side of the tcgen05 implicit-GEMM kernel: tensor-map encoding (driver entry point fetched
// through the runtime, so libcuda is not a link-time dependency), template dispatch, and the raw
// GEMM entry point asn_gemm_bf16_tn.
#include "umma_host.cuh"

#include <stdlib.h>

namespace asn {
namespace umma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int encode(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                  const uint32_t* box) {
  EncodeTiledFn fn = get_encode();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return ASN_ECUDA;
  }
  if (reinterpret_cast<uintptr_t>(base) & 15) {
    set_error("tensor map base %p is not 16-byte aligned", base);
    return ASN_EINVAL;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (box[i] == 0 || box[i] > 256) {
      set_error("tensor map box[%d]=%u out of range", i, box[i]);
      return ASN_EINVAL;
    }
  }
  for (int i = 0; i < rank - 1; ++i) {
    gstr[i] = strides[i];
    if (strides[i] & 15) {
      set_error("tensor map stride[%d]=%llu is not a multiple of 16 bytes", i, (unsigned long long)strides[i]);
      return ASN_EINVAL;
    }
  }
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu,%llu box %u,%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return ASN_ECUDA;
  }
  return ASN_OK;
}

int encode_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
              uint32_t box_rows) {
  uint64_t dims[2] = {inner, rows};
  uint64_t strides[1] = {row_stride_bytes};
  uint32_t box[2] = {64, box_rows};
  return encode(m, base, 2, dims, strides, box);
}

int encode_4d(CUtensorMap* m, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
              uint32_t box1, uint32_t box2) {
  uint32_t box[4] = {64, box1, box2, 1};
  return encode(m, base, 4, dims, strides_bytes, box);
}

// Launch shape per (mode, BLOCK_N): EW = 8 -> one persistent CTA per SM, ring of <= 184 KB; EW = 4 -> two CTAs per SM,
// ring of <= 100 KB each (umma.cuh, EpiCfg).  Measured on B200: the weight-gradient tiles (long K over pixels, N <= 128)
// and the 32-column dgrad of conv1 run faster as two pipelines per SM, everything else with the eight-warp epilogue.
template <int MODE, int BLOCK_N, int MT, int CL = 1>
struct Shape {
  static constexpr int EW =
      (MT == 1 && CL == 1 && ((MODE == MODE_WGRAD && BLOCK_N <= 128) || (MODE == MODE_CONV && BLOCK_N <= 32))) ? 4 : 8;
  static_assert(MODE != MODE_GEMM_MN || BLOCK_N % 64 == 0, "MN-major B tiles are made of 64-column boxes");
  static constexpr int ctas_per_sm = EW == 4 ? 2 : 1;
  static constexpr int b_tile = (((BLOCK_N / CL) * 128 + 1023) / 1024) * 1024;
  // K-major modes: MT row tiles of A per B tile; WGRAD: MT taps' B tiles per A tile
  static constexpr int stage = (MODE == MODE_WGRAD || MODE == MODE_DGRAD4) ? 16384 + MT * b_tile : MT * 16384 + b_tile;
  // ring budget: 184 KB, or 172 KB when the 48 KB of TMA-store staging replace the 36 KB of transpose buffers (same number
  // of stages for every BLOCK_N in use)
  static constexpr int fit = ((EW == 4 ? 100 : (has_tma_epi<MODE, BLOCK_N, MT, EW>() ? 172 : 184)) * 1024) / stage;
  static constexpr int stages = fit > 8 ? 8 : (fit < 2 ? 2 : fit);
};

// cluster size the launcher uses for (mode, block_n): 2 = CTA pairs (tcgen05 cta_group::2) over x-neighbouring
// tiles -- callers must encode the B tensor map with box rows block_n / 2 (each CTA of the pair fetches half of the B
// tile).  ASN_PAIR=0 falls back to single-CTA tiles everywhere (A/B measurements; parity tests pass in both settings).
int cluster_size(int mode, int block_n) {
  static const bool off = getenv("ASN_PAIR") != nullptr && getenv("ASN_PAIR")[0] == '0';
  if (off) return 1;
  // measured on B200 (profiles/README.md): pairs pay for the GEMMs (ASPP forward / dgrad: +5-10 %) and the MN-major
  // weight gradients (+25-50 %), not for the implicit-GEMM convolutions, whose 4-D activation boxes bound them
  static const bool conv_pairs = getenv("ASN_PAIR_CONV") != nullptr && getenv("ASN_PAIR_CONV")[0] == '1';
  if (mode == MODE_CONV) return conv_pairs && block_n >= 64 ? 2 : 1;
  if (mode == MODE_GEMM) return block_n >= 128 ? 2 : 1;
  // MN-major modes: each CTA of the pair holds block_n / 2 channels of B as 64-channel boxes; callers additionally
  // need an even number of 128-row M tiles (pairs are x-neighbours) and 128 rows of A per CTA
  return block_n % 128 == 0 ? 2 : 1;
}

template <int MODE, int BLOCK_N, int CL, int MT>
static int launch_t(const CUtensorMap maps[5], const CUtensorMap* omaps, const Params& P, dim3 grid, cudaStream_t st,
                    const char* name, double flops, double bytes) {
  using Sh = Shape<MODE, BLOCK_N, MT, CL>;
  constexpr int STAGES = Sh::stages, EW = Sh::EW, NUM_THREADS = EpiCfg<EW>::threads;
  using L = KernelSmem<MODE, BLOCK_N, STAGES, CL, MT, EW>;
  auto kern = umma_kernel<MODE, BLOCK_N, STAGES, CL, MT, EW>;
  static PerDevice configured, clusters;  // clusters: CL = 2, CTA pairs the device can keep resident at once (one per TPC)
  int max_clusters = clusters.get();
  if (!configured.get()) {
    ASN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    if (CL > 1) {
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(sm_count() / CL * CL);
      q.blockDim = dim3(NUM_THREADS);
      q.dynamicSmemBytes = L::TOTAL;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = CL;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      q.attrs = qa;
      q.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &q) != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        max_clusters = sm_count() / CL;
      }
      clusters.set(max_clusters);
    }
    configured.set(1);
  }
  Params Pp = P;
  // output tensor maps of the TMA-store epilogue (CONV / EPI_BF16 tiles made of 64-channel slabs); anything else keeps the
  // register -> shared memory -> st.global epilogue and never touches them
  const CUtensorMap* om = omaps ? omaps : maps;
  if (!omaps || !has_tma_epi<MODE, BLOCK_N, MT, EW>()) Pp.tma_store = 0;
  Pp.grid_x = (int)grid.x;
  Pp.grid_y = (int)grid.y;
  Pp.grid_z = (int)grid.z;
  const long long tiles = (long long)((grid.x + CL - 1) / CL) * grid.y * grid.z;  // (pairs of) tiles
  const long long slots = CL == 1 ? (long long)sm_count() * Sh::ctas_per_sm : (long long)max_clusters;
  const int ctas = (int)(tiles < slots ? tiles : slots) * CL;
  prof::Scope ps(name, flops, bytes, st);
  if (CL == 1) {
    kern<<<ctas, NUM_THREADS, L::TOTAL, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], om[0], om[1], om[2], om[3], Pp);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ASN_CUDA(cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], om[0], om[1], om[2], om[3], Pp));
  }
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

bool block_n_supported(int mode, int block_n) {
  switch (mode) {
    case MODE_GEMM: return block_n == 128 || block_n == 176 || block_n == 256;
    case MODE_CONV: return block_n == 32 || block_n == 64 || block_n == 128 || block_n == 256;
    case MODE_WGRAD: return block_n == 64 || block_n == 128 || block_n == 256;
    case MODE_GEMM_MN: return block_n == 128 || block_n == 192 || block_n == 256;
    case MODE_DGRAD4: return block_n == 32 || block_n == 64;
  }
  return false;
}

// 256-row CTA tiles (MT = 2) pay off when there is still more than a wave of them; `m_tiles_128` = number of
// 128-row tiles along M, `other` = tiles along the remaining grid dimensions.
int rows_per_cta(int mode, long long m_tiles_128, long long other) {
  // measured on B200: no gain at the shapes of this path (the epilogue, not the operand feed, bounds them) -> opt-in
  static const bool on = getenv("ASN_MT2") != nullptr && getenv("ASN_MT2")[0] == '1';
  if (!on || mode == MODE_WGRAD) return BLOCK_M;
  const long long tiles2 = ((m_tiles_128 + 1) / 2) * other;
  return tiles2 * 4 >= 5LL * sm_count() ? 2 * BLOCK_M : BLOCK_M;
}

int launch(int mode, int block_n, const CUtensorMap maps[5], const Params& P, dim3 grid, cudaStream_t st,
           const char* prof_name, double prof_flops, double prof_bytes, int rows, const CUtensorMap* omaps) {
  int cl = rows == BLOCK_M ? cluster_size(mode, block_n) : 1;
  // MN-major pairs are two x-neighbouring 128-row M tiles of the same N tile: needs an even number of M tiles
  if ((mode == MODE_WGRAD || mode == MODE_GEMM_MN) && (cdiv(P.M, BLOCK_M) % 2 != 0)) cl = 1;
  const int mt = rows / BLOCK_M;
#define ASN_CASE(MODE, BN, CLS, MTS) \
  if (mode == MODE && block_n == BN && cl == CLS && mt == MTS)  \
    return launch_t<MODE, BN, CLS, MTS>(maps, omaps, P, grid, st, prof_name, prof_flops, prof_bytes);
  ASN_CASE(MODE_GEMM, 128, 1, 1)
  ASN_CASE(MODE_GEMM, 176, 1, 1)
  ASN_CASE(MODE_GEMM, 256, 1, 1)
  ASN_CASE(MODE_GEMM, 128, 2, 1)
  ASN_CASE(MODE_GEMM, 176, 2, 1)
  ASN_CASE(MODE_GEMM, 256, 2, 1)
  ASN_CASE(MODE_GEMM, 128, 1, 2)
  ASN_CASE(MODE_GEMM, 176, 1, 2)
  ASN_CASE(MODE_GEMM, 256, 1, 2)
  ASN_CASE(MODE_CONV, 32, 1, 1)
  ASN_CASE(MODE_CONV, 64, 1, 1)
  ASN_CASE(MODE_CONV, 128, 1, 1)
  ASN_CASE(MODE_CONV, 256, 1, 1)
  ASN_CASE(MODE_CONV, 64, 2, 1)
  ASN_CASE(MODE_CONV, 128, 2, 1)
  ASN_CASE(MODE_CONV, 256, 2, 1)
  ASN_CASE(MODE_CONV, 32, 1, 2)
  ASN_CASE(MODE_CONV, 64, 1, 2)
  ASN_CASE(MODE_CONV, 128, 1, 2)
  ASN_CASE(MODE_CONV, 256, 1, 2)
  ASN_CASE(MODE_WGRAD, 64, 1, 1)
  ASN_CASE(MODE_WGRAD, 128, 1, 1)
  ASN_CASE(MODE_WGRAD, 256, 1, 1)
  ASN_CASE(MODE_DGRAD4, 32, 1, 4)
  ASN_CASE(MODE_DGRAD4, 64, 1, 4)
  ASN_CASE(MODE_WGRAD, 64, 1, 2)
  ASN_CASE(MODE_WGRAD, 64, 1, 4)
  ASN_CASE(MODE_WGRAD, 128, 1, 2)
  ASN_CASE(MODE_WGRAD, 128, 2, 1)
  ASN_CASE(MODE_WGRAD, 256, 2, 1)
  ASN_CASE(MODE_GEMM_MN, 192, 1, 1)
  ASN_CASE(MODE_GEMM_MN, 128, 1, 1)
  ASN_CASE(MODE_GEMM_MN, 256, 1, 1)
  ASN_CASE(MODE_GEMM_MN, 128, 2, 1)
  ASN_CASE(MODE_GEMM_MN, 256, 2, 1)
#undef ASN_CASE
  set_error("umma::launch: no kernel for mode %d block_n %d rows %d", mode, block_n, rows);
  return ASN_EUNSUPPORTED;
}

// plain GEMM: C[M,N] = A[M,K] . B[N,K]^T  (+ split-K partials)
int gemm_tn(const void* A, const void* B, float* C, int M, int N, int K, int lda, int ldb, long long ldc,
            int split_k, long long split_stride, int block_n, cudaStream_t st, const char* prof_name,
            double prof_flops) {
  ASN_CHECK_ARG(A && B && C, "gemm_tn: null pointer");
  ASN_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_tn: bad shape %d %d %d", M, N, K);
  ASN_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && lda >= K && ldb >= K, "gemm_tn: lda/ldb must be >= K and multiples of 8");
  ASN_CHECK_ARG(block_n_supported(MODE_GEMM, block_n), "gemm_tn: unsupported block_n %d", block_n);
  CUtensorMap maps[5];
  const int k_steps_all = cdiv(K, BLOCK_K);
  int split_req = split_k < 1 ? 1 : (split_k > k_steps_all ? k_steps_all : split_k);
  const int rows = rows_per_cta(MODE_GEMM, cdiv(M, BLOCK_M), (long long)cdiv(N, block_n) * split_req);
  int rc = encode_2d(&maps[0], A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, (uint32_t)rows);
  if (rc) return rc;
  rc = encode_2d(&maps[4], B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2,
                 (uint32_t)(block_n / (rows == BLOCK_M ? cluster_size(MODE_GEMM, block_n) : 1)));
  if (rc) return rc;
  maps[1] = maps[2] = maps[3] = maps[0];
  Params P;
  memset(&P, 0, sizeof(P));
  P.M = M;
  P.N = N;
  P.k_steps = cdiv(K, BLOCK_K);
  if (split_k < 1) split_k = 1;
  if (split_k > P.k_steps) split_k = P.k_steps;
  P.steps_per_split = cdiv(P.k_steps, split_k);
  split_k = cdiv(P.k_steps, P.steps_per_split);
  P.epi = EPI_F32;
  P.out = C;
  P.ld_out = ldc;
  P.z_stride_out = split_stride;
  P.slope = 1.f;
  dim3 grid(cdiv(M, rows), cdiv(N, block_n), split_k);
  return launch(MODE_GEMM, block_n, maps, P, grid, st, prof_name, prof_flops >= 0 ? prof_flops : 2.0 * M * N * K,
                2.0 * M * K + 2.0 * N * K + 4.0 * M * N * split_k, rows);
}

// C[M,N] (fp32, ldc; split_k partials at split_stride) = A^T . B for A[K][M] (lda), B[K][N] (ldb), bf16, M / N contiguous
int gemm_nt_mn(const void* A, const void* B, float* C, int M, int N, int K, int lda, int ldb, long long ldc,
               int split_k, long long split_stride, int block_n, cudaStream_t st, const char* prof_name,
               double prof_flops) {
  ASN_CHECK_ARG(A && B && C, "gemm_nt_mn: null pointer");
  ASN_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_nt_mn: bad shape %d %d %d", M, N, K);
  ASN_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && lda >= M && ldb >= N, "gemm_nt_mn: lda/ldb must be >= M/N and multiples of 8");
  ASN_CHECK_ARG(block_n_supported(MODE_GEMM_MN, block_n), "gemm_nt_mn: unsupported block_n %d", block_n);
  CUtensorMap maps[5];
  int rc = encode_2d(&maps[0], A, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64);
  if (rc) return rc;
  rc = encode_2d(&maps[4], B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64);
  if (rc) return rc;
  maps[1] = maps[2] = maps[3] = maps[0];
  Params P;
  memset(&P, 0, sizeof(P));
  P.M = M;
  P.N = N;
  P.k_steps = cdiv(K, BLOCK_K);
  if (split_k < 1) split_k = 1;
  if (split_k > P.k_steps) split_k = P.k_steps;
  P.steps_per_split = cdiv(P.k_steps, split_k);
  split_k = cdiv(P.k_steps, P.steps_per_split);
  P.a_boxes = M <= 64 ? 1 : 2;
  P.epi = EPI_F32;
  P.out = C;
  P.ld_out = ldc;
  P.z_stride_out = split_stride;
  P.slope = 1.f;
  dim3 grid(cdiv(M, BLOCK_M), cdiv(N, block_n), split_k);
  return launch(MODE_GEMM_MN, block_n, maps, P, grid, st, prof_name, prof_flops >= 0 ? prof_flops : 2.0 * M * N * K,
                2.0 * M * K + 2.0 * N * K + 4.0 * M * N * split_k);
}

int effective_split(int K, int split_k) {
  int k_steps = cdiv(K, BLOCK_K);
  if (split_k < 1) split_k = 1;
  if (split_k > k_steps) split_k = k_steps;
  int sps = cdiv(k_steps, split_k);
  return cdiv(k_steps, sps);
}

}  // namespace umma
}  // namespace asn

extern "C" int asn_gemm_bf16_nt_mn(const void* A, const void* B, float* C, int M, int N, int K, int lda, int ldb,
                                   int ldc, int split_k, void* stream) {
  using namespace asn;
  return umma::gemm_nt_mn(A, B, C, M, N, K, lda, ldb, ldc, split_k, (long long)M * ldc, N > 128 ? 192 : 128,
                          static_cast<cudaStream_t>(stream), "gemm_nt_mn", -1.0);
}

extern "C" int asn_gemm_bf16_tn(const void* A, const void* B, float* C, int M, int N, int K, int lda, int ldb,
                                int ldc, int split_k, void* stream) {
  using namespace asn;
  int bn = N > 128 ? 256 : 128;
  if (const char* e = getenv("ASN_GEMM_BN")) bn = atoi(e);  // probing only (tools/gemm_probe.py)
  return umma::gemm_tn(A, B, C, M, N, K, lda, ldb, ldc, split_k, (long long)M * ldc, bn,
                       static_cast<cudaStream_t>(stream));
}
