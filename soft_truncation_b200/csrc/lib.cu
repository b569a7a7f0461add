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
  // under stream capture the attribute becomes a programmatic dependency edge of the graph (CUDA >= 12.3):
  // ST_PDL_CAPTURE=0 keeps captured launches fully serialised
  static const int in_capture = [] { const char* v = getenv("ST_PDL_CAPTURE"); return v ? atoi(v) : 1; }();
  if (in_capture) return true;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) { cudaGetLastError(); return false; }
  return st == cudaStreamCaptureStatusNone;
}

// Device-resident addend of every in-kernel dropout seed (see st_set_dropout_seed_offset in st_b200.h).
static const uint64_t* g_seed_offset = nullptr;
const uint64_t* st_seed_offset() { return g_seed_offset; }
extern "C" __attribute__((visibility("default"))) int st_set_dropout_seed_offset(const void* dev_u64) {
  g_seed_offset = reinterpret_cast<const uint64_t*>(dev_u64);
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_version(void) { return 200; }
extern "C" __attribute__((visibility("default"))) const char* st_last_error(void) { return g_err; }
