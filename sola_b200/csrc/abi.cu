// C-ABI plumbing shared by every entry point: version, per-thread error string, launch check.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace sola {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;   // kernel launches issued through this library (all threads)

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return SOLA_ERR_CUDA;
  }
  return SOLA_OK;
}

}  // namespace sola

extern "C" {

// major*10000 + minor*100 + patch
int sola_version(void) { return 201; }

const char* sola_last_error_string(void) { return sola::g_err; }

// Compiled architecture tag, so the loader can refuse a library built for something else.
const char* sola_build_arch(void) { return "sm_100a"; }

// sha256 of csrc/ + the nvcc flags this library was compiled from (sola_b200/_build.py passes it with -D): the loader compares
// it with the digest of the sources it sees and refuses a stale library.
#ifndef SOLA_SOURCE_DIGEST
#define SOLA_SOURCE_DIGEST "unknown"
#endif
const char* sola_build_digest(void) { return SOLA_SOURCE_DIGEST; }

// Number of kernel launches issued by this library in this process (all threads); used by bench.py's gpu_launches.
unsigned long long sola_launch_count(void) { return __atomic_load_n(&sola::g_launches, __ATOMIC_RELAXED); }

}
