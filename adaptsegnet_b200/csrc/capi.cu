// error / device plumbing of the C ABI.
#include "common.cuh"

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <string.h>
#include <vector>

namespace asn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int device_index() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev < ASN_MAX_DEVICES ? dev : ASN_MAX_DEVICES - 1;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

namespace prof {

struct Record {
  const char* name;
  double flops, bytes;
  cudaEvent_t e0, e1;
};
static bool g_on = false;
static std::vector<Record> g_recs;
static std::vector<cudaEvent_t> g_pool;
static std::mutex g_mu;
constexpr size_t MAX_RECORDS = 200000;

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool enabled() { return g_on; }

static cudaEvent_t get_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

void begin(const char* name, double flops, double bytes, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_recs.size() >= MAX_RECORDS) return;
  Record r{name, flops, bytes, get_event(), nullptr};
  cudaEventRecord(r.e0, st);
  g_recs.push_back(r);
}

void end(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_recs.empty() || g_recs.back().e1) return;
  g_recs.back().e1 = get_event();
  cudaEventRecord(g_recs.back().e1, st);
}

static void clear() {
  for (auto& r : g_recs) {
    if (r.e0) g_pool.push_back(r.e0);
    if (r.e1) g_pool.push_back(r.e1);
  }
  g_recs.clear();
}

}  // namespace prof
}  // namespace asn

extern "C" int64_t asn_launch_count(void) { return asn::prof::g_launches.load(); }

extern "C" int asn_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(asn::prof::g_mu);
  asn::prof::clear();
  asn::prof::g_on = on != 0;
  return ASN_OK;
}

// JSON: {"name": {"launches": n, "ms": total, "flops": total, "bytes": total}, ...}; returns the length needed
extern "C" int64_t asn_prof_report(char* buf, int64_t cap) {
  using namespace asn::prof;
  std::lock_guard<std::mutex> lk(g_mu);
  struct Agg { long long n; double ms, flops, bytes; };
  std::map<std::string, Agg> agg;
  for (auto& r : g_recs) {
    if (!r.e1) continue;
    if (cudaEventSynchronize(r.e1) != cudaSuccess) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
    Agg& a = agg[r.name];
    a.n += 1; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
  }
  std::string out = "{";
  bool first = true;
  char tmp[512];
  for (auto& kv : agg) {
    snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}",
             first ? "" : ", ", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops, kv.second.bytes);
    out += tmp;
    first = false;
  }
  out += "}";
  if (buf && cap > 0) {
    size_t n = out.size() < (size_t)cap - 1 ? out.size() : (size_t)cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (int64_t)out.size() + 1;
}

#ifndef ASN_BUILD_ID
#define ASN_BUILD_ID "unknown"
#endif
extern "C" const char* asn_build_id(void) { return ASN_BUILD_ID; }
// names of the recorded scopes in launch order, '\n'-separated (profiling tools: lines up ncu's launch list with the
// library's own kernel names); returns the length needed
extern "C" int64_t asn_prof_sequence(char* buf, int64_t cap) {
  using namespace asn::prof;
  std::lock_guard<std::mutex> lk(g_mu);
  std::string out;
  for (auto& r : g_recs) {
    out += r.name;
    out += '\n';
  }
  if (buf && cap > 0) {
    size_t n = out.size() < (size_t)cap - 1 ? out.size() : (size_t)cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (int64_t)out.size() + 1;
}

extern "C" int asn_abi_version(void) { return ASN_ABI_VERSION; }
extern "C" const char* asn_last_error(void) { return asn::g_err; }
extern "C" int asn_sm_count(int* out_host) {
  if (!out_host) return ASN_EINVAL;
  *out_host = asn::sm_count();
  return ASN_OK;
}
