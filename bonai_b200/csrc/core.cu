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

static int g_sm_reserve = 0;

// SMs the persistent grids may fill: all of them, minus what loft_reserve_sms() set aside for a
// concurrently running collective (its CTAs cannot share an SM with a 197 KB-smem GEMM CTA; a
// statically striped persistent grid that does not fit runs a ragged second wave instead).
int loft_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  const int avail = n - g_sm_reserve;
  return avail > 2 ? (avail & ~1) : 2;
}

extern "C" {
const char* loft_last_error(void) { return g_err; }
int loft_abi_version(void) { return 3; }
int loft_reserve_sms(int n) {
  const int prev = g_sm_reserve;
  g_sm_reserve = n > 0 ? n : 0;
  return prev;
}
}
