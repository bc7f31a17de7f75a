// Library-level entry points and error plumbing.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace incagg {

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }
unsigned long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int sm_count() {
  static thread_local int cached[16] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int32_t* device_error_word() {
  static int32_t* words[16] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  if (words[dev] == nullptr) {
    int32_t* w = nullptr;
    if (cudaMalloc(&w, sizeof(int32_t)) != cudaSuccess) return nullptr;
    cudaMemset(w, 0, sizeof(int32_t));
    words[dev] = w;
  }
  return words[dev];
}

static int g_tune[INCAGG_TUNE_COUNT];
static bool g_tune_set[INCAGG_TUNE_COUNT];
int tune_get(int key, int dflt) {
  if (key < 0 || key >= INCAGG_TUNE_COUNT || !g_tune_set[key]) return dflt;
  return g_tune[key];
}

bool pdl_enabled() {
  static const bool on = [] { const char* v = getenv("INCAGG_PDL"); return v != nullptr && v[0] == '1'; }();
  return on;
}

}  // namespace incagg

extern "C" int incagg_version(void) { return 104; }

extern "C" int incagg_tune_set(int key, int value) {
  IA_CHECK_ARG(key >= 0 && key < INCAGG_TUNE_COUNT, "unknown tuning key %d", key);
  incagg::g_tune[key] = value;
  incagg::g_tune_set[key] = true;
  return INCAGG_OK;
}

extern "C" int incagg_device_errors(int32_t* out, int reset) {
  IA_CHECK_ARG(out != nullptr, "out is NULL");
  int32_t* w = incagg::device_error_word();
  IA_CHECK_ARG(w != nullptr, "no device error word");
  IA_CUDA(cudaMemcpy(out, w, sizeof(int32_t), cudaMemcpyDeviceToHost));  // synchronises with the device
  if (reset && *out != 0) IA_CUDA(cudaMemset(w, 0, sizeof(int32_t)));
  return INCAGG_OK;
}

extern "C" int64_t incagg_launch_count(void) { return (int64_t)incagg::launches(); }

extern "C" const char* incagg_last_error(void) { return incagg::err_buf(); }

extern "C" int incagg_device_info(int* sm_count_out, int* cc_major, int* cc_minor) {
  int dev = 0;
  IA_CUDA(cudaGetDevice(&dev));
  int n = 0, maj = 0, min_ = 0;
  IA_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  IA_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  IA_CUDA(cudaDeviceGetAttribute(&min_, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count_out) *sm_count_out = n;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min_;
  return INCAGG_OK;
}

// Let kernels of the current device load / store memory of `peer_device` over NVLink (the history
// shards of other ranks, mapped into this process through CUDA IPC).  Idempotent.
extern "C" int incagg_enable_peer_access(int peer_device) {
  int dev = 0;
  IA_CUDA(cudaGetDevice(&dev));
  if (dev == peer_device) return INCAGG_OK;
  int can = 0;
  IA_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  IA_CHECK_ARG(can != 0, "device %d cannot access peer device %d", dev, peer_device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return INCAGG_OK;
  }
  IA_CUDA(e);
  return INCAGG_OK;
}
