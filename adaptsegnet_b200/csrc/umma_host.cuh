// host-side interface of the tcgen05 implicit-GEMM kernel (implemented in umma.cu)
#pragma once
#include "umma.cuh"

namespace asn {
namespace umma {

// bf16 row-major matrix [rows][inner] with `row_stride_bytes` between rows (multiple of 16),
// box = 64 x box_rows, SWIZZLE_128B, zero fill out of bounds.
int encode_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
              uint32_t box_rows);
// bf16 tensor with dims {d0 (contiguous), d1, d2, d3}, byte strides for d1..d3, box {64, b1, b2, 1}
int encode_4d(CUtensorMap* m, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
              uint32_t box1, uint32_t box2);

// maps = {a0, a1, a2, a3, b}.  grid as documented on umma_kernel.
// rows = rows of the CTA tile: 128, or 256 (two accumulators sharing every B stage; the A tensor map / pixel box must
// then cover 256 rows).  grid.x counts tiles of that height.
int launch(int mode, int block_n, const CUtensorMap maps[5], const Params& P, dim3 grid, cudaStream_t st,
           const char* prof_name = "umma", double prof_flops = 0, double prof_bytes = 0, int rows = BLOCK_M,
           const CUtensorMap* omaps = nullptr);
// omaps: four {64 ch, tw, th, 1} SWIZZLE_128B maps over the bf16 NHWC output, one per blockIdx.z class (CONV mode with
// P.tma_store = 1): the tile is written with TMA stores (encode with encode_4d, box = the pixel tile)
// 128 or 256: the tile height the launcher recommends for this many 128-row tiles x other grid dimensions
int rows_per_cta(int mode, long long m_tiles_128, long long other);

// BLOCK_N choices compiled for each mode
bool block_n_supported(int mode, int block_n);
// 1 or 2: with 2 (CTA pairs, cta_group::2), encode the B tensor map with box rows block_n / 2 (each CTA loads its half)
int cluster_size(int mode, int block_n);

}  // namespace umma
}  // namespace asn

namespace asn {
namespace umma {
// C[M,N] (fp32, ldc) = A[M,K] . B[N,K]^T, bf16 K-contiguous operands; split_k partials at split_stride
int gemm_tn(const void* A, const void* B, float* C, int M, int N, int K, int lda, int ldb, long long ldc,
            int split_k, long long split_stride, int block_n, cudaStream_t st, const char* prof_name = "gemm_bf16_tn",
            double prof_flops = -1.0);
// C[M,N] (fp32, ldc) = A^T . B for bf16 A[K][M], B[K][N] (M / N contiguous): MN-major operands, no transposed copies
int gemm_nt_mn(const void* A, const void* B, float* C, int M, int N, int K, int lda, int ldb, long long ldc,
               int split_k, long long split_stride, int block_n, cudaStream_t st,
               const char* prof_name = "gemm_bf16_nt_mn", double prof_flops = -1.0);
// number of z-slices gemm_tn actually launches for a requested split
int effective_split(int K, int split_k);
}  // namespace umma
}  // namespace asn
