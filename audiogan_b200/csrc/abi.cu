// Library-level entry points: version, error string, device info.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

namespace ag {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static thread_local char g_path[256] = "";
static thread_local char g_decline[192] = "";
void set_decline(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_decline, sizeof(g_decline), fmt, ap);
  va_end(ap);
}
void clear_decline() { g_decline[0] = 0; }
// records which kernel family took the call; a pending decline reason (a fast path that refused the shape) is appended and,
// with AUDIOGAN_VERBOSE=1, printed once per distinct message
void set_path(const char* family) {
  if (g_decline[0]) snprintf(g_path, sizeof(g_path), "%s (%s)", family, g_decline);
  else snprintf(g_path, sizeof(g_path), "%s", family);
  if (g_decline[0]) {
    static int verbose = -1;
    if (verbose < 0) { const char* e = getenv("AUDIOGAN_VERBOSE"); verbose = (e && e[0] == '1') ? 1 : 0; }
    if (verbose) {
      static char seen[16][256];
      static int nseen = 0;
      bool dup = false;
      for (int i = 0; i < nseen; ++i) dup = dup || strcmp(seen[i], g_path) == 0;
      if (!dup && nseen < 16) {
        strncpy(seen[nseen++], g_path, 255);
        fprintf(stderr, "[audiogan_b200] recurrent kernel fell off the fast path: %s\n", g_path);
      }
    }
  }
  g_decline[0] = 0;
}
const char* last_path() { return g_path; }
static int g_sm = 0, g_smem = 0, g_cc = 0;
static int query() {
  if (g_sm) return 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return -1;
  g_smem = (int)p.sharedMemPerBlockOptin;
  g_cc = p.major * 10 + p.minor;
  g_sm = p.multiProcessorCount;
  return 0;
}
int sm_count() { query(); return g_sm ? g_sm : 148; }
int smem_optin() { query(); return g_smem ? g_smem : 232448; }
}  // namespace ag

extern "C" {
int ag_version(void) { return 100; }
const char* ag_last_error_string(void) { return ag::g_err; }
const char* ag_lstm_last_path(void) { return ag::last_path(); }
int ag_sync_check(void* stream) {
  AG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  AG_CUDA(cudaGetLastError());
  return AG_OK;
}
int ag_device_info(int* sm, int* smem, int* cc) {
  if (ag::query() != 0) { ag::set_error("no CUDA device"); return AG_ECUDA; }
  if (sm) *sm = ag::g_sm;
  if (smem) *smem = ag::g_smem;
  if (cc) *cc = ag::g_cc;
  return AG_OK;
}
}
