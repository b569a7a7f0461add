// Library-level entry points: version, last-error text.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void st_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" __attribute__((visibility("default"))) int st_version(void) { return 100; }
extern "C" __attribute__((visibility("default"))) const char* st_last_error(void) { return g_err; }
