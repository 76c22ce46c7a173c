// Fused optimizer steps over flat fp32 buffers: one launch per optimizer instead of the
// ~10 multi-tensor launches torch.optim issues, and one pass over the data.
//   asn_sgd_step  : torch.optim.SGD(momentum, weight_decay; dampening 0, no nesterov) as constructed at
//                   train_gta2cityscapes_multi.py:244,347,532 and stepped at :681.  The reference's parameter groups list
//                   most trunk parameters several times (model/deeplab_multi.py:196-218 walks modules AND their
//                   sub-modules), so its optimizer applies the update `repeat` times per step to those
//                   parameters; the kernel does the `repeat` updates in registers, in order, exactly as a sequential
//                   loop over the list would.
//   asn_adam_step : torch.optim.Adam(betas, eps; no weight decay, no amsgrad), train...:351-355,538-540, stepped at :682-683.
// HBM-bound: 20 B (SGD) / 28 B (Adam) per parameter.
#include "../../include/asn_b200.h"
#include "common.cuh"

namespace asn {

constexpr int OPT_THREADS = 256;
constexpr int OPT_MAX_GROUPS = 8;

struct SgdArgs {
  float* p;
  const float* g;
  float* buf;
  long long n4;                 // float4 elements
  const long long* seg_begin;   // [n_seg + 1] element offsets (multiples of 4), ascending
  const int* seg_group;         // [n_seg] learning-rate group of the segment
  const int* seg_repeat;        // [n_seg] how many times the reference's parameter list names it
  int n_seg;
  float lr[OPT_MAX_GROUPS];
  float momentum, weight_decay;
  float grad_scale;             // gradients are multiplied by this first (1 / world size under data parallelism)
  int first_step;
};

__device__ __forceinline__ void sgd_one(float& p, float g, float& buf, float lr, float mu, float wd, int repeat,
                                        bool first) {
  for (int r = 0; r < repeat; ++r) {
    const float d = wd != 0.f ? fmaf(wd, p, g) : g;  // d_p = grad + wd * p
    // first step(): torch collects the (absent) buffers of ALL list entries before updating any, so every mention of
    // the parameter starts a fresh buffer = d_p and the last one is kept (torch/optim/sgd.py, _single_tensor_sgd)
    buf = first ? d : fmaf(mu, buf, d);
    p = fmaf(-lr, buf, p);
  }
}

// Thread = OPT_VEC float4 elements, OPT_THREADS apart (every load instruction of a warp is 512 contiguous bytes); all
// 3 x OPT_VEC loads are issued before the first use.  The segment table (where each parameter begins, its learning-rate
// group and repeat count) is staged in shared memory once per CTA, so the per-element lookup is a handful of shared-memory
// reads instead of a chain of dependent global loads in front of the streaming loads.
constexpr int OPT_VEC = 4;
constexpr int OPT_MAX_SEG_SMEM = 1024;

__global__ void __launch_bounds__(OPT_THREADS) sgd_step_kernel(const SgdArgs a) {
  __shared__ long long s_begin[OPT_MAX_SEG_SMEM];
  __shared__ int s_meta[OPT_MAX_SEG_SMEM];   // group | repeat << 8
  const bool staged = a.n_seg <= OPT_MAX_SEG_SMEM;
  if (staged) {
    for (int i = threadIdx.x; i < a.n_seg; i += OPT_THREADS) {
      s_begin[i] = __ldg(a.seg_begin + i);
      s_meta[i] = __ldg(a.seg_group + i) | (__ldg(a.seg_repeat + i) << 8);
    }
    __syncthreads();
  }
  const long long base = (long long)blockIdx.x * (OPT_THREADS * OPT_VEC) + threadIdx.x;
  float4 p[OPT_VEC], g[OPT_VEC], b[OPT_VEC];
#pragma unroll
  for (int k = 0; k < OPT_VEC; ++k) {
    const long long i = base + (long long)k * OPT_THREADS;
    if (i < a.n4) {
      p[k] = reinterpret_cast<const float4*>(a.p)[i];
      g[k] = ld_stream(reinterpret_cast<const float4*>(a.g) + i);
      b[k] = a.first_step ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<const float4*>(a.buf)[i];
    }
  }
  const bool first = a.first_step != 0;
#pragma unroll
  for (int k = 0; k < OPT_VEC; ++k) {
    const long long i = base + (long long)k * OPT_THREADS;
    if (i >= a.n4) continue;
    const long long e = i * 4;
    int lo = 0, hi = a.n_seg - 1;  // last segment whose begin <= e
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      const long long bm = staged ? s_begin[mid] : __ldg(a.seg_begin + mid);
      if (bm <= e) lo = mid; else hi = mid - 1;
    }
    const int meta = staged ? s_meta[lo] : (__ldg(a.seg_group + lo) | (__ldg(a.seg_repeat + lo) << 8));
    const float lr = a.lr[meta & 0xff];
    const int repeat = meta >> 8;
    if (a.grad_scale != 1.f) {  // a separate rounded multiply: bit-identical to `grad.mul_(scale)` followed by the step
      g[k].x = __fmul_rn(g[k].x, a.grad_scale); g[k].y = __fmul_rn(g[k].y, a.grad_scale);
      g[k].z = __fmul_rn(g[k].z, a.grad_scale); g[k].w = __fmul_rn(g[k].w, a.grad_scale);
    }
    sgd_one(p[k].x, g[k].x, b[k].x, lr, a.momentum, a.weight_decay, repeat, first);
    sgd_one(p[k].y, g[k].y, b[k].y, lr, a.momentum, a.weight_decay, repeat, first);
    sgd_one(p[k].z, g[k].z, b[k].z, lr, a.momentum, a.weight_decay, repeat, first);
    sgd_one(p[k].w, g[k].w, b[k].w, lr, a.momentum, a.weight_decay, repeat, first);
    reinterpret_cast<float4*>(a.p)[i] = p[k];
    reinterpret_cast<float4*>(a.buf)[i] = b[k];
  }
}

__global__ void __launch_bounds__(OPT_THREADS)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 long long n4, float beta1, float beta2, float eps, float step_size, float inv_sqrt_bc2,
                 float grad_scale) {
  const long long base = (long long)blockIdx.x * (OPT_THREADS * OPT_VEC) + threadIdx.x;
  float4 P[OPT_VEC], G[OPT_VEC], M[OPT_VEC], V[OPT_VEC];
#pragma unroll
  for (int k = 0; k < OPT_VEC; ++k) {   // all loads first: 4 x OPT_VEC independent 16-byte loads in flight per thread
    const long long i = base + (long long)k * OPT_THREADS;
    if (i < n4) {
      P[k] = reinterpret_cast<const float4*>(p)[i];
      G[k] = ld_stream(reinterpret_cast<const float4*>(g) + i);
      M[k] = reinterpret_cast<const float4*>(m)[i];
      V[k] = reinterpret_cast<const float4*>(v)[i];
    }
  }
  auto one = [&](float& pp, float gg, float& mm, float& vv) {
    if (grad_scale != 1.f) gg = __fmul_rn(gg, grad_scale);
    mm = fmaf(beta1, mm, (1.f - beta1) * gg);          // exp_avg.lerp_(grad, 1 - beta1)
    vv = fmaf(beta2, vv, (1.f - beta2) * gg * gg);     // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
    pp = fmaf(-step_size, mm / denom, pp);
  };
#pragma unroll
  for (int k = 0; k < OPT_VEC; ++k) {
    const long long i = base + (long long)k * OPT_THREADS;
    if (i >= n4) continue;
    one(P[k].x, G[k].x, M[k].x, V[k].x);
    one(P[k].y, G[k].y, M[k].y, V[k].y);
    one(P[k].z, G[k].z, M[k].z, V[k].z);
    one(P[k].w, G[k].w, M[k].w, V[k].w);
    reinterpret_cast<float4*>(p)[i] = P[k];
    reinterpret_cast<float4*>(m)[i] = M[k];
    reinterpret_cast<float4*>(v)[i] = V[k];
  }
}

}  // namespace asn

using namespace asn;

extern "C" int asn_sgd_step(float* params, const float* grads, float* momentum_buf, int64_t n,
                            const int64_t* seg_begin, const int* seg_group, const int* seg_repeat, int n_seg,
                            const float* group_lr_host, int n_groups, float momentum, float weight_decay,
                            int first_step, float grad_scale, void* stream) {
  ASN_CHECK_ARG(params && grads && momentum_buf && seg_begin && seg_group && seg_repeat && group_lr_host,
                "asn_sgd_step: null pointer");
  ASN_CHECK_ARG(n > 0 && n % 4 == 0 && n_seg > 0, "asn_sgd_step: n must be a positive multiple of 4 (got %lld)", (long long)n);
  ASN_CHECK_ARG(n_groups > 0 && n_groups <= OPT_MAX_GROUPS, "asn_sgd_step: 1..%d learning-rate groups", OPT_MAX_GROUPS);
  ASN_CHECK_ARG(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) |
                  reinterpret_cast<uintptr_t>(momentum_buf)) & 15) == 0, "asn_sgd_step: buffers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SgdArgs a;
  memset(&a, 0, sizeof(a));
  a.p = params; a.g = grads; a.buf = momentum_buf;
  a.n4 = n / 4;
  a.seg_begin = reinterpret_cast<const long long*>(seg_begin);
  a.seg_group = seg_group; a.seg_repeat = seg_repeat; a.n_seg = n_seg;
  for (int i = 0; i < n_groups; ++i) a.lr[i] = group_lr_host[i];
  a.momentum = momentum; a.weight_decay = weight_decay; a.first_step = first_step;
  a.grad_scale = grad_scale;
  prof::Scope ps("sgd_step", 0, 5.0 * 4.0 * (double)n, st);
  sgd_step_kernel<<<(unsigned)cdiv(a.n4, (long long)OPT_THREADS * OPT_VEC), OPT_THREADS, 0, st>>>(a);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

extern "C" int asn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, int64_t step, float grad_scale, void* stream) {
  ASN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq, "asn_adam_step: null pointer");
  ASN_CHECK_ARG(n > 0 && n % 4 == 0 && step >= 1, "asn_adam_step: n must be a positive multiple of 4, step >= 1");
  ASN_CHECK_ARG(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) |
                  reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0,
                "asn_adam_step: buffers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // torch/optim/adam.py (_single_tensor_adam): step_size = lr / (1 - beta1^t); denom = sqrt(v) / sqrt(1 - beta2^t) + eps
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  prof::Scope ps("adam_step", 0, 7.0 * 4.0 * (double)n, st);
  adam_step_kernel<<<(unsigned)cdiv(n / 4, (int64_t)OPT_THREADS * OPT_VEC), OPT_THREADS, 0, st>>>(
      params, grads, exp_avg, exp_avg_sq, n / 4, beta1, beta2, eps, step_size, inv_sqrt_bc2, grad_scale);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}
