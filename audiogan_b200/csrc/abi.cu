// Library-level entry points: version, error string, device info.
#include "common.cuh"
#include <stdarg.h>

namespace ag {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
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
