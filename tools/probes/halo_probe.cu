// Round-2 feasibility probe: can a tcgen05.mma A descriptor point at a window of a
// SWIZZLE_128B shared-memory tile that starts r0 rows (r0 * 128 bytes) below the 1024-byte aligned tile base?
// One CTA: TMA loads A[144][64] bf16 (SW128) and B[64][64]; for every r0 in 0..15 and both settings of the
// descriptor's base-offset field it computes D[128][64] = A[r0 .. r0+127] . B^T and compares with the host result.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I adaptsegnet_b200/csrc tools/probes/halo_probe.cu -o build/halo_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_bf16.h>
#include "umma.cuh"

using namespace asn::umma;

__device__ __forceinline__ uint64_t desc_with_base_offset(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  return make_smem_desc(saddr, lbo, sbo) | ((uint64_t)(base_off & 7u) << 49);
}

__global__ void __launch_bounds__(128)
halo_probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* out,
                  int r0, int use_base_offset) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sa = base, sb = base + 144 * 128, bar = sb + 64 * 128, bar2 = bar + 8, holder = bar + 16;
  volatile uint32_t* holder_ptr = reinterpret_cast<volatile uint32_t*>(raw + (holder - smem_u32(raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc<1>(holder, 64);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *holder_ptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 144 * 128 + 64 * 128);
    tma_load_2d(&map_a, sa, bar, 0, 0);
    tma_load_2d(&map_b, sb, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc(128, 64, 0, 0);
    for (int k = 0; k < 4; ++k) {
      const uint32_t a_start = sa + r0 * 128 + k * 32;
      const uint64_t da = desc_with_base_offset(a_start, 16, 1024, use_base_offset ? (uint32_t)(r0 & 7) : 0u);
      const uint64_t db = make_smem_desc(sb + k * 32, 16, 1024);
      mma_f16_ss(tmem, da, db, idesc, k > 0 ? 1u : 0u);
    }
    mma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  float v[32];
  for (int c = 0; c < 64; c += 32) {
    __syncwarp();
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 64 + c + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode(CUtensorMap* m, void* base, uint64_t rows, uint32_t box_rows) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return 1;
  cuuint64_t gdim[2] = {64, rows};
  cuuint64_t gstr[1] = {128};
  cuuint32_t box[2] = {64, box_rows}, estr[2] = {1, 1};
  return reinterpret_cast<EncodeTiledFn>(p)(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstr, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

int main() {
  std::vector<__nv_bfloat16> A(144 * 64), B(64 * 64);
  std::vector<float> Af(144 * 64), Bf(64 * 64);
  srand(1);
  for (size_t i = 0; i < A.size(); ++i) { A[i] = __float2bfloat16((rand() % 17 - 8) / 8.f); Af[i] = __bfloat162float(A[i]); }
  for (size_t i = 0; i < B.size(); ++i) { B[i] = __float2bfloat16((rand() % 13 - 6) / 4.f); Bf[i] = __bfloat162float(B[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dOut;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dOut, 128 * 64 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ma, mb;
  if (encode(&ma, dA, 144, 144) || encode(&mb, dB, 64, 64)) { printf("tensor map encode failed\n"); return 1; }
  const size_t smem = 144 * 128 + 64 * 128 + 64 + 1024;
  cudaFuncSetAttribute(halo_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> out(128 * 64);
  printf("{\n");
  for (int mode = 0; mode < 2; ++mode) {
    printf(" \"base_offset_field_%s\": {", mode ? "r0_mod_8" : "zero");
    for (int r0 = 0; r0 < 16; ++r0) {
      cudaMemset(dOut, 0, out.size() * 4);
      halo_probe_kernel<<<1, 128, smem>>>(ma, mb, dOut, r0, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("\"cuda_error\": \"%s\"}}\n", cudaGetErrorString(e)); return 2; }
      cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0, maxref = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += (double)Af[(r0 + m) * 64 + k] * Bf[n * 64 + k];
          maxerr = fmax(maxerr, fabs(ref - out[m * 64 + n]));
          maxref = fmax(maxref, fabs(ref));
        }
      printf("%s\"%d\": %.3g", r0 ? ", " : "", r0, maxerr / maxref);
    }
    printf("}%s\n", mode ? "" : ",");
  }
  printf("}\n");
  return 0;
}
