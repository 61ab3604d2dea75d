// Error string, device query and ABI version of libloft_b200.so.
#include "common.cuh"
#include "loft_b200.h"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void loft_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int loft_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

extern "C" {
const char* loft_last_error(void) { return g_err; }
int loft_abi_version(void) { return 2; }
}
