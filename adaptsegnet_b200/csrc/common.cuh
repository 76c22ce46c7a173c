// Shared helpers for libasn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/asn_b200.h"

namespace asn {

void set_error(const char* fmt, ...);
int sm_count();
int device_index();  // current CUDA device clamped to [0, ASN_MAX_DEVICES)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and occupancy answers are PER DEVICE: one-time kernel configuration
// is remembered per device (a process may drive several GPUs, e.g. nn.DataParallel), with atomics because several
// host threads may launch concurrently.  The configuration calls are idempotent, so a lost race only repeats them.
constexpr int ASN_MAX_DEVICES = 64;
struct PerDevice {
  int v[ASN_MAX_DEVICES];  // zero-initialised (static storage)
  int get() const { return __atomic_load_n(&v[device_index()], __ATOMIC_ACQUIRE); }
  void set(int x) { __atomic_store_n(&v[device_index()], x, __ATOMIC_RELEASE); }
};

// Optional per-kernel timing with CUDA events on the launching stream (asn_prof_enable /
// asn_prof_report): bench.py uses it to time the dominant kernel live inside the timed steps.
namespace prof {
void count_launch();
bool enabled();
void begin(const char* name, double flops, double bytes, cudaStream_t st);
void end(cudaStream_t st);
struct Scope {
  cudaStream_t st;
  bool on;
  Scope(const char* name, double flops, double bytes, cudaStream_t s) : st(s), on(enabled()) {
    count_launch();
    if (on) begin(name, flops, bytes, st);
  }
  ~Scope() {
    if (on) end(st);
  }
};
}  // namespace prof

#define ASN_CHECK_ARG(cond, ...)              \
  do {                                        \
    if (!(cond)) {                            \
      asn::set_error(__VA_ARGS__);            \
      return ASN_EINVAL;                      \
    }                                         \
  } while (0)

#define ASN_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      asn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return ASN_ECUDA;                                                                  \
    }                                                                                    \
  } while (0)

#define ASN_LAUNCH_CHECK() ASN_CUDA(cudaGetLastError())

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// grid for a grid-stride kernel: whole waves of `ctas_per_sm` CTAs on every SM
static inline int wave_grid(int64_t work_items, int threads, int ctas_per_sm) {
  int64_t need = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// one work item per thread: CTAs launch and retire continuously, which keeps a steady stream of loads in
// flight (a capped grid-stride grid phase-locks: every CTA loads, then every CTA computes)
static inline int full_grid(int64_t work_items, int threads) {
  int64_t need = (work_items + threads - 1) / threads;
  if (need < 1) need = 1;
  const int64_t cap = 0x7fffffff;
  return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[o] = sum_{n,p} src[n][o][p] by ONE CTA of 256 threads (fp64 accumulation, fixed order -> deterministic)
__device__ __forceinline__ void block_channel_sum(const float* __restrict__ src, float* __restrict__ out, int N, int O,
                                                  int P, int o) {
  __shared__ double cs_part[8];
  double acc = 0.0;
  for (int n = 0; n < N; ++n) {
    const float* s = src + ((int64_t)n * O + o) * P;
    if ((P & 3) == 0 && (reinterpret_cast<uintptr_t>(s) & 15) == 0) {
      // 16-byte loads, eight in flight per thread; four fp32 values folded pairwise into the double accumulator per load
#pragma unroll 8
      for (int i = threadIdx.x; i < P / 4; i += 256) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(s) + i);
        acc += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
      }
    } else {
      for (int i = threadIdx.x; i < P; i += 256) acc += (double)s[i];
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) cs_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < 8 ? cs_part[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[o] = (float)v;
  }
}

// streaming 128-bit accesses (read-once / write-once tensors)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// bilinear source coordinate, align_corners=True, exactly the float ops ATen performs:
// scale = (float)(in-1)/(out-1) (0 when out == 1); r = scale*dst; i0 = (int)r; lam = r - i0.
struct Lerp {
  int i0, i1;
  float l0, l1;
};
__host__ __device__ __forceinline__ float lerp_scale(int n_in, int n_out) {
  return n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.f;
}
__device__ __forceinline__ Lerp lerp_at(int dst, float scale, int n_in) {
  Lerp L;
  float r = __fmul_rn(scale, (float)dst);  // no FMA contraction: lam must equal the oracle's
  L.i0 = min((int)r, n_in - 1);
  L.i1 = L.i0 + (L.i0 < n_in - 1 ? 1 : 0);
  L.l1 = __fsub_rn(r, (float)L.i0);
  L.l0 = __fsub_rn(1.f, L.l1);
  return L;
}

}  // namespace asn
