// Error plumbing and library-wide helpers of libtedm_b200.so (see include/tedm_b200.h).
#include "common.cuh"
#include "../../include/tedm_b200.h"

#include <cstring>

static thread_local char g_err[512] = "";

int tedm_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int tedm_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

extern "C" int tedm_version(void) { return TEDM_ABI_VERSION; }
extern "C" const char* tedm_last_error(void) { return g_err; }
