// Shared helpers for the incagg_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/incagg_b200.h"

namespace incagg {

// ---- error plumbing (thread-local text, C-ABI status codes) ---------------
char* err_buf();
int set_err(int code, const char* fmt, ...);

#define IA_CHECK_ARG(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) return incagg::set_err(INCAGG_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define IA_CUDA(call)                                                               \
  do {                                                                              \
    cudaError_t _e = (call);                                                        \
    if (_e != cudaSuccess)                                                          \
      return incagg::set_err(INCAGG_ERR_CUDA, "%s failed: %s (%s:%d)", #call,        \
                             cudaGetErrorString(_e), __FILE__, __LINE__);           \
  } while (0)

// Every kernel launch is followed by IA_LAUNCH_CHECK(), which also counts it (incagg_launch_count).
#define IA_LAUNCH_CHECK()                                                           \
  do {                                                                              \
    incagg::count_launch();                                                         \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess)                                                          \
      return incagg::set_err(INCAGG_ERR_CUDA, "kernel launch failed: %s (%s:%d)",    \
                             cudaGetErrorString(_e), __FILE__, __LINE__);           \
  } while (0)

inline cudaStream_t as_stream(incagg_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // cached SM count of the current device (148 on B200)
// True exactly once per (flag array, current device): function attributes such as the dynamic
// shared-memory limit are per device, so a "done" flag must be too.
inline bool first_use_on_device(bool (&done)[16]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return true;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}
int tune_get(int key, int dflt);  // experiment knobs (incagg_tune_set); `dflt` when unset
// Device-side error word of the current device (one int32, zero = no error): kernels OR a bit into it
// when they meet an index outside its table (INCAGG_DEVERR_*), the host reads it with
// incagg_device_errors().  Allocated on first use; nullptr if that fails.
int32_t* device_error_word();
void count_launch();  // process-wide count of kernels launched by this library
unsigned long long launches();

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
__host__ __device__ inline bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }

// ---- device helpers --------------------------------------------------------
// Streaming (read-once) loads that do not pollute L1: index/value arrays.
__device__ __forceinline__ int ldg_stream(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Inclusive warp scan.
template <typename T>
__device__ __forceinline__ T warp_scan_incl(T v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// Block-wide exclusive scan for blockDim.x == BLOCK (multiple of 32, <= 1024).
// Returns the exclusive prefix of `v`; *total receives the block sum.
template <typename T, int BLOCK>
__device__ __forceinline__ T block_scan_excl(T v, T* total) {
  __shared__ T warp_tot[BLOCK / 32];
  __shared__ T block_tot;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T incl = warp_scan_incl(v);
  if (lane == 31) warp_tot[w] = incl;
  __syncthreads();
  if (w == 0) {
    T t = (lane < BLOCK / 32) ? warp_tot[lane] : T(0);
    T ti = warp_scan_incl(t);
    if (lane < BLOCK / 32) warp_tot[lane] = ti - t;  // exclusive warp offsets
    if (lane == 31) block_tot = ti;
  }
  __syncthreads();
  T out = incl - v + warp_tot[w];
  *total = block_tot;
  __syncthreads();  // shared arrays may be reused by the next call
  return out;
}

// ---- device-wide exclusive scan over int64 counts (3-phase, deterministic) --
// Used by relabel / transpose.  `scratch` needs scan_scratch_elems(n) int64 slots.
constexpr int SCAN_BLOCK = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;
inline int64_t scan_num_tiles(int64_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------
// Consecutive kernels of the library on one stream (or consecutive kernel nodes of a captured graph)
// leave the GPU idle for a launch latency between them.  With INCAGG_PDL=1 every launch carries the
// programmatic-stream-serialization attribute and every kernel starts with
//     griddepcontrol.launch_dependents   (the next kernel's CTAs may be scheduled as SMs free up)
//     griddepcontrol.wait                (block until the previous grid has completed and flushed)
// so the next kernel's launch and prologue overlap this kernel's tail; all global memory accesses come
// after the wait, which keeps the stream's semantics.  Without the attribute both instructions are
// no-ops.  Kernels of other libraries in between (ATen) are ordinary launches and fully serialise.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_trigger();
  pdl_wait();
}
bool pdl_enabled();   // common.cu: INCAGG_PDL=1

template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  if (!pdl_enabled()) {
    kernel<<<grid, block, smem, st>>>(static_cast<KArgs>(args)...);
    return;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace incagg
