// K7: 19x19 confusion matrix (compute_iou.py:15-17) -- HBM-bound integer kernel.
//
// Layout: label (u8 | i32 | i64) and pred (u8) are flat arrays of n_px pixels.
// Each thread streams 16 consecutive pixels per iteration with 128-bit loads
// (1 load of pred, 1/4/8 loads of labels).  Groups of four identical (a,b) pairs
// -- recognised on the raw words -- extend a run kept in registers (labels are
// blocky, so most shared-memory atomics vanish); mixed groups add pixel by pixel
// into a warp-private shared-memory histogram.  One 64-bit global
// atomic per non-zero bin per CTA at the end.  Algorithmic bytes: 2 B/px (u8
// labels) or 9 B/px (i64 labels, as the reference holds them).
// Optional 256-entry LUT applied to the labels on the way in: compute_iou.py:24-28 (label_mapping, one full pass over
// the label image per mapping entry in the reference) fused into the counting.
#include "common.cuh"

namespace asn {

constexpr int HIST_THREADS = 256;
constexpr int HIST_WARPS = HIST_THREADS / 32;

// A chunk of 16 consecutive labels as loaded (raw words), decoded lazily: class index or -1 for anything outside [0, n).
// uniform4(g): the four labels of group g (pixels 4g .. 4g+3) are bit-identical -- decided on the raw words.
template <typename T>
struct RawLabels;

template <>
struct RawLabels<uint8_t> {
  uint32_t w[4];
  __device__ __forceinline__ void load(const uint8_t* p) {
    const uint4 v = ld_stream(reinterpret_cast<const uint4*>(p));
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  }
  __device__ __forceinline__ bool uniform4(int g) const { return w[g] == (w[g] & 0xffu) * 0x01010101u; }
  __device__ __forceinline__ int decode(int i, int n, const uint8_t* lut) const {
    int x = (w[i >> 2] >> ((i & 3) * 8)) & 0xff;
    if (lut) x = lut[x];
    return x < n ? x : -1;
  }
  static __device__ __forceinline__ int load1(const uint8_t* p, int n, const uint8_t* lut) {
    int x = *p;
    if (lut) x = lut[x];
    return x < n ? x : -1;
  }
};
template <>
struct RawLabels<int32_t> {
  uint32_t w[16];
  // labels outside [0, 256) are not in the LUT: label_mapping leaves them unchanged (and they are invalid anyway)
  static __device__ __forceinline__ int map1(uint32_t x, int n, const uint8_t* lut) {
    if (lut && x < 256u) x = lut[x];
    return x < (uint32_t)n ? (int)x : -1;
  }
  __device__ __forceinline__ void load(const int32_t* p) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + j);   // (L1-allocating: see the int64 loader)
      w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
    }
  }
  __device__ __forceinline__ bool uniform4(int g) const {
    return w[4 * g] == w[4 * g + 1] && w[4 * g + 2] == w[4 * g + 3] && w[4 * g] == w[4 * g + 2];
  }
  __device__ __forceinline__ int decode(int i, int n, const uint8_t* lut) const { return map1(w[i], n, lut); }
  static __device__ __forceinline__ int load1(const int32_t* p, int n, const uint8_t* lut) {
    return map1((uint32_t)*p, n, lut);
  }
};
template <>
struct RawLabels<int64_t> {
  uint32_t lo[16], hi[16];
  static __device__ __forceinline__ int map1(uint32_t l, uint32_t h, int n, const uint8_t* lut) {
    if (h != 0u) return -1;  // negative or >= 2^32: never valid, never in the LUT
    if (lut && l < 256u) l = lut[l];
    return l < (uint32_t)n ? (int)l : -1;
  }
  __device__ __forceinline__ void load(const int64_t* p) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // little endian: (x,y) = first int64, (z,w) = second.  A thread's 128 bytes are one cache line and a warp instruction
      // touches 32 different lines, 16 bytes each: the loads allocate in L1 (plain __ldg, not the streaming form), so the
      // line comes from L2 once and the other seven loads hit L1 instead of requesting the same sectors again
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + j);
      lo[2 * j] = v.x; hi[2 * j] = v.y; lo[2 * j + 1] = v.z; hi[2 * j + 1] = v.w;
    }
  }
  __device__ __forceinline__ bool uniform4(int g) const {
    return lo[4 * g] == lo[4 * g + 1] && lo[4 * g + 2] == lo[4 * g + 3] && lo[4 * g] == lo[4 * g + 2] &&
           (hi[4 * g] | hi[4 * g + 1] | hi[4 * g + 2] | hi[4 * g + 3]) == 0u;   // (a non-zero high word: slow path, all invalid)
  }
  __device__ __forceinline__ int decode(int i, int n, const uint8_t* lut) const { return map1(lo[i], hi[i], n, lut); }
  static __device__ __forceinline__ int load1(const int64_t* p, int n, const uint8_t* lut) {
    const unsigned long long x = (unsigned long long)*p;
    return map1((uint32_t)x, (uint32_t)(x >> 32), n, lut);
  }
};

// run-length aggregation in registers: equal flat indices that follow each other cost one compare and one add
struct RunAgg {
  int cur;
  uint32_t cnt;
  uint32_t ovf;
  __device__ __forceinline__ void push(int idx, uint32_t k, uint32_t* h, int nbins) {
    if (idx == cur) {
      cnt += k;
    } else {
      flush(h, nbins);
      cur = idx;
      cnt = k;
    }
  }
  __device__ __forceinline__ void flush(uint32_t* h, int nbins) {
    if (cur >= 0 && cnt) {
      if (cur < nbins)
        atomicAdd(&h[cur], cnt);
      else
        ovf += cnt;
    }
    cnt = 0;
  }
};

// 16 pixels in four groups of four.  A group whose four (label, prediction) pairs are identical -- the common case in
// segmentation maps, decided with two compares on the raw words -- decodes ONE label and extends the current run by 4;
// any other group goes pixel by pixel straight to the shared-memory histogram (decode, index, one ATOMS each: no run
// bookkeeping, no data-dependent branch per pixel).
template <typename LabelT>
__device__ __forceinline__ void hist_chunk(const RawLabels<LabelT>& L, const uint4& pv, int n_cls, const uint8_t* lut,
                                           RunAgg& agg, uint32_t* myh, int nbins) {
  const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint32_t w = pw[g];
    const uint32_t b0 = w & 0xffu;
    if (w == b0 * 0x01010101u && L.uniform4(g)) {
      const int a = L.decode(4 * g, n_cls, lut);
      agg.push(a >= 0 ? a * n_cls + (int)b0 : -1, 4u, myh, nbins);
    } else {
      agg.flush(myh, nbins);
      agg.cur = -1;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int a = L.decode(4 * g + i, n_cls, lut);
        const int idx = a * n_cls + (int)((w >> (8 * i)) & 0xffu);
        if (a >= 0) {
          if (idx < nbins)
            atomicAdd(&myh[idx], 1u);
          else
            ++agg.ovf;
        }
      }
    }
  }
}

template <typename LabelT>
__global__ void __launch_bounds__(HIST_THREADS, 4)
fast_hist_kernel(const LabelT* __restrict__ label, const uint8_t* __restrict__ pred, int64_t n_px,
                 int n_cls, int n_sub, unsigned long long* __restrict__ hist,
                 unsigned long long* __restrict__ overflow, int vec_ok, const uint8_t* __restrict__ lut_g) {
  extern __shared__ uint32_t sh[];
  __shared__ uint8_t lut_s[256];
  const int nbins = n_cls * n_cls;
  for (int i = threadIdx.x; i < nbins * n_sub; i += HIST_THREADS) sh[i] = 0;
  if (lut_g) lut_s[threadIdx.x] = lut_g[threadIdx.x];  // HIST_THREADS == 256
  const uint8_t* lut = lut_g ? lut_s : nullptr;
  __syncthreads();
  uint32_t* myh = sh + ((threadIdx.x >> 5) % n_sub) * nbins;

  RunAgg agg{-1, 0, 0};
  const int64_t tid = (int64_t)blockIdx.x * HIST_THREADS + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * HIST_THREADS;
  const int64_t n_vec = vec_ok ? (n_px >> 4) : 0;  // chunks of 16 pixels
  // u8 labels move only 32 bytes per thread and chunk: the next chunk is requested before the current one is counted, so
  // every thread always has a pair of 16-byte loads in flight (the wider label types carry 80 / 144 bytes per chunk already)
  constexpr bool PREFETCH = sizeof(LabelT) == 1;
  if (PREFETCH) {
    RawLabels<LabelT> L;
    uint4 pv = make_uint4(0u, 0u, 0u, 0u);
    int64_t c = tid;
    if (c < n_vec) {
      L.load(label + c * 16);
      pv = ld_stream(reinterpret_cast<const uint4*>(pred + c * 16));
    }
    while (c < n_vec) {
      const int64_t cn = c + nthreads;
      RawLabels<LabelT> Ln = L;
      uint4 pn = pv;
      if (cn < n_vec) {
        Ln.load(label + cn * 16);
        pn = ld_stream(reinterpret_cast<const uint4*>(pred + cn * 16));
      }
      hist_chunk<LabelT>(L, pv, n_cls, lut, agg, myh, nbins);
      L = Ln;
      pv = pn;
      c = cn;
    }
  } else {
    for (int64_t c = tid; c < n_vec; c += nthreads) {
      RawLabels<LabelT> L;
      L.load(label + c * 16);
      const uint4 pv = ld_stream(reinterpret_cast<const uint4*>(pred + c * 16));
      hist_chunk<LabelT>(L, pv, n_cls, lut, agg, myh, nbins);
    }
  }
  for (int64_t i = n_vec * 16 + tid; i < n_px; i += nthreads) {
    const int a = RawLabels<LabelT>::load1(label + i, n_cls, lut);
    agg.push(a >= 0 ? a * n_cls + (int)pred[i] : -1, 1u, myh, nbins);
  }
  agg.flush(myh, nbins);
  if (agg.ovf) atomicAdd(overflow, (unsigned long long)agg.ovf);
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += HIST_THREADS) {
    unsigned long long s = 0;
    for (int k = 0; k < n_sub; ++k) s += sh[k * nbins + i];
    if (s) atomicAdd(&hist[i], s);
  }
}

template <typename LabelT>
static int launch_hist(const void* label, const uint8_t* pred, int64_t n_px, int n_cls,
                       int64_t* hist, int64_t* overflow, const uint8_t* lut, cudaStream_t st) {
  static_assert(HIST_THREADS == 256, "the kernel stages the 256-entry LUT with one thread per entry");
  const int nbins = n_cls * n_cls;
  int n_sub = (48 * 1024) / (nbins * 4);
  if (n_sub > HIST_WARPS) n_sub = HIST_WARPS;
  ASN_CHECK_ARG(n_sub >= 1, "asn_fast_hist: n_cls=%d too large for the shared-memory histogram", n_cls);
  int vec_ok = ((reinterpret_cast<uintptr_t>(label) | reinterpret_cast<uintptr_t>(pred)) & 15) == 0;
  // 16 px per thread-iteration; at least 4 iterations per thread before adding CTAs; a persistent grid of exactly the
  // 4 CTAs per SM the launch bounds guarantee (one wave: no second, partly filled round of the grid-stride loop)
  int grid = wave_grid((n_px + 63) / 64, HIST_THREADS, 4);
  // 32-bit shared counters: keep every CTA below 2^31 pixels
  while ((n_px + grid - 1) / grid > (int64_t)1 << 31) grid *= 2;
  prof::Scope ps("fast_hist", 0, (double)n_px * (sizeof(LabelT) + 1), st);
  fast_hist_kernel<LabelT><<<grid, HIST_THREADS, (size_t)n_sub * nbins * 4, st>>>(
      static_cast<const LabelT*>(label), pred, n_px, n_cls, n_sub,
      reinterpret_cast<unsigned long long*>(hist), reinterpret_cast<unsigned long long*>(overflow),
      vec_ok, lut);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

// compute_iou.py:20-21,61-64 on the device: iu[c] = hist[c][c] / (row_c + col_c - hist[c][c]) in float64 (0/0 -> nan,
// like numpy), miou = nanmean(iu) (nan when every class is nan).  The int64 counts are exact in float64 below 2^53, which
// is how the reference holds them (np.zeros((n, n)) accumulates float64).  One warp: lane strides the classes.
__global__ void per_class_iu_kernel(const long long* __restrict__ hist, int n, double* __restrict__ iu,
                                    double* __restrict__ miou) {
  const int lane = threadIdx.x;
  double sum = 0.0;
  long long cnt = 0;
  for (int c = lane; c < n; c += 32) {
    double row = 0.0, col = 0.0;
    for (int j = 0; j < n; ++j) {   // numpy sums the float64 matrix entry by entry in this order
      row += (double)hist[(long long)c * n + j];
      col += (double)hist[(long long)j * n + c];
    }
    const double d = (double)hist[(long long)c * n + c];
    const double v = d / (row + col - d);
    iu[c] = v;
    if (v == v) { sum += v; ++cnt; }
  }
  // per-lane partial sums are combined in lane order so the result does not depend on the shuffle tree
  __shared__ double s_sum[32];
  __shared__ long long s_cnt[32];
  s_sum[lane] = sum;
  s_cnt[lane] = cnt;
  __syncwarp();
  if (lane == 0) {
    double t = 0.0;
    long long k = 0;
    for (int i = 0; i < 32; ++i) { t += s_sum[i]; k += s_cnt[i]; }
    *miou = k ? t / (double)k : nan("");
  }
}

}  // namespace asn

extern "C" int asn_per_class_iu(const int64_t* hist, int n_cls, double* iu, double* miou, void* stream) {
  using namespace asn;
  ASN_CHECK_ARG(hist && iu && miou, "asn_per_class_iu: null pointer");
  ASN_CHECK_ARG(n_cls >= 1 && n_cls <= 4096, "asn_per_class_iu: bad n_cls");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  prof::Scope ps("per_class_iu", 0, 8.0 * n_cls * n_cls, st);
  per_class_iu_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const long long*>(hist), n_cls, iu, miou);
  ASN_LAUNCH_CHECK();
  return ASN_OK;
}

static int fast_hist_impl(const void* label, int label_dtype, const uint8_t* lut, const uint8_t* pred, int64_t n_px,
                          int n_cls, int64_t* hist, int64_t* overflow, void* stream) {
  using namespace asn;
  ASN_CHECK_ARG(n_px >= 0 && n_cls >= 1 && n_cls <= 255, "asn_fast_hist: bad n_px/n_cls");
  ASN_CHECK_ARG(hist && overflow, "asn_fast_hist: null output");
  if (n_px == 0) return ASN_OK;
  ASN_CHECK_ARG(label && pred, "asn_fast_hist: null input");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (label_dtype) {
    case ASN_LABEL_U8: return launch_hist<uint8_t>(label, pred, n_px, n_cls, hist, overflow, lut, st);
    case ASN_LABEL_I32: return launch_hist<int32_t>(label, pred, n_px, n_cls, hist, overflow, lut, st);
    case ASN_LABEL_I64: return launch_hist<int64_t>(label, pred, n_px, n_cls, hist, overflow, lut, st);
  }
  set_error("asn_fast_hist: unknown label_dtype %d", label_dtype);
  return ASN_EINVAL;
}

extern "C" int asn_fast_hist(const void* label, int label_dtype, const uint8_t* pred, int64_t n_px,
                             int n_cls, int64_t* hist, int64_t* overflow, void* stream) {
  return fast_hist_impl(label, label_dtype, nullptr, pred, n_px, n_cls, hist, overflow, stream);
}

extern "C" int asn_fast_hist_lut(const void* label, int label_dtype, const uint8_t* lut256, const uint8_t* pred,
                                 int64_t n_px, int n_cls, int64_t* hist, int64_t* overflow, void* stream) {
  ASN_CHECK_ARG(lut256, "asn_fast_hist_lut: null LUT");
  return fast_hist_impl(label, label_dtype, lut256, pred, n_px, n_cls, hist, overflow, stream);
}
