// error / device plumbing of the C ABI.
#include "common.cuh"

namespace asn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
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

}  // namespace asn

extern "C" int asn_abi_version(void) { return ASN_ABI_VERSION; }
extern "C" const char* asn_last_error(void) { return asn::g_err; }
extern "C" int asn_sm_count(int* out_host) {
  if (!out_host) return ASN_EINVAL;
  *out_host = asn::sm_count();
  return ASN_OK;
}
