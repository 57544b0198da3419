#include "common.h"

#include <cstring>

namespace mtp {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static bool g_pdl = true;
bool pdl_enabled() { return g_pdl; }
void set_pdl(bool on) { g_pdl = on; }

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

}  // namespace mtp

extern "C" const char* mtp_last_error(void) { return mtp::g_err; }
extern "C" int mtp_version(void) { return 100; }
extern "C" int mtp_set_pdl(int enabled) {
  mtp::set_pdl(enabled != 0);
  return MTP_OK;
}
extern "C" int mtp_num_sms(void) { return mtp::num_sms(); }
