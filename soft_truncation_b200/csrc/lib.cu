// Library-level entry points: version, last-error text.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void st_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool st_pdl_on(cudaStream_t stream) {
  static const int enabled = [] { const char* v = getenv("ST_PDL"); return v ? atoi(v) : 1; }();
  if (!enabled) return false;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) { cudaGetLastError(); return false; }
  return st == cudaStreamCaptureStatusNone;
}

extern "C" __attribute__((visibility("default"))) int st_version(void) { return 100; }
extern "C" __attribute__((visibility("default"))) const char* st_last_error(void) { return g_err; }
