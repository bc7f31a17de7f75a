// Device-wide exclusive scan over int64 counts: three small launches, fixed combination
// order (deterministic).  Used by relabel and the CSR transpose; n is a row count (<= a few
// million), so this is latency- not bandwidth-relevant.
#pragma once
#include "common.cuh"

namespace incagg {

static __global__ void __launch_bounds__(SCAN_BLOCK)
scan_tile_sums_kernel(const int64_t* __restrict__ in, int64_t n, int64_t* __restrict__ tile_sums) {
  pdl_prologue();
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int64_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int64_t i = base + (int64_t)k * SCAN_BLOCK + threadIdx.x;
    if (i < n) s += in[i];
  }
  int64_t total;
  (void)block_scan_excl<int64_t, SCAN_BLOCK>(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// One block: exclusive scan of the tile sums in place; grand total to *total_out.
static __global__ void __launch_bounds__(SCAN_BLOCK)
scan_tiles_kernel(int64_t* __restrict__ tile_sums, int64_t num_tiles, int64_t* __restrict__ total_out) {
  pdl_prologue();
  int64_t carry = 0;
  for (int64_t b = 0; b < num_tiles; b += SCAN_BLOCK) {
    const int64_t i = b + threadIdx.x;
    const int64_t v = (i < num_tiles) ? tile_sums[i] : 0;
    int64_t total;
    const int64_t ex = block_scan_excl<int64_t, SCAN_BLOCK>(v, &total);
    if (i < num_tiles) tile_sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

// out[i] = exclusive prefix of in[i] (in place allowed).  Thread t owns SCAN_ITEMS consecutive
// elements of its tile so the in-thread order is the array order.
static __global__ void __launch_bounds__(SCAN_BLOCK)
scan_apply_kernel(const int64_t* in, int64_t n, const int64_t* __restrict__ tile_offsets,
                  int64_t* out) {
  pdl_prologue();
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int64_t v[SCAN_ITEMS];
  int64_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int64_t i = base + k;
    v[k] = (i < n) ? in[i] : 0;
    s += v[k];
  }
  int64_t total;
  int64_t ex = block_scan_excl<int64_t, SCAN_BLOCK>(s, &total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int64_t i = base + k;
    if (i < n) out[i] = ex;
    ex += v[k];
  }
}

// Two launches for up to SCAN_BLOCK tiles (n <= 2 M): tile sums, then every block adds up the sums
// of the tiles before it by itself (<= 512 values) and scans its own tile.
static __global__ void __launch_bounds__(SCAN_BLOCK)
scan_apply_self_kernel(const int64_t* in, int64_t n, const int64_t* __restrict__ tile_sums, int64_t num_tiles,
                       int64_t* out, int64_t* __restrict__ total_out) {
  pdl_prologue();
  __shared__ int64_t s_off;
  // offset of this tile = sum of tile_sums[0 .. blockIdx.x); the last block also publishes the total
  int64_t part = 0, all = 0;
  for (int64_t i = threadIdx.x; i < num_tiles; i += SCAN_BLOCK) {
    const int64_t v = tile_sums[i];
    all += v;
    if (i < (int64_t)blockIdx.x) part += v;
  }
  int64_t tot_part, tot_all;
  (void)block_scan_excl<int64_t, SCAN_BLOCK>(part, &tot_part);
  (void)block_scan_excl<int64_t, SCAN_BLOCK>(all, &tot_all);
  if (threadIdx.x == 0) {
    s_off = tot_part;
    if (blockIdx.x == 0) *total_out = tot_all;
  }
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int64_t v[SCAN_ITEMS];
  int64_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int64_t i = base + k;
    v[k] = (i < n) ? in[i] : 0;
    s += v[k];
  }
  int64_t total;
  int64_t ex = block_scan_excl<int64_t, SCAN_BLOCK>(s, &total) + s_off;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int64_t i = base + k;
    if (i < n) out[i] = ex;
    ex += v[k];
  }
}

// scratch: at least scan_num_tiles(n) int64 slots.  total_out: device int64.
static inline int exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, int64_t* scratch,
                                     int64_t* total_out, cudaStream_t st) {
  if (n == 0) {
    IA_CUDA(cudaMemsetAsync(total_out, 0, sizeof(int64_t), st));
    return INCAGG_OK;
  }
  if (scan_num_tiles(n) <= SCAN_BLOCK) {
    const int64_t tiles = scan_num_tiles(n);
    launch(scan_tile_sums_kernel, dim3((unsigned)tiles), dim3(SCAN_BLOCK), (size_t)(0), st, in, n, scratch);
    IA_LAUNCH_CHECK();
    launch(scan_apply_self_kernel, dim3((unsigned)tiles), dim3(SCAN_BLOCK), (size_t)(0), st, in, n, scratch, tiles, out, total_out);
    IA_LAUNCH_CHECK();
    return INCAGG_OK;
  }
  const int64_t tiles = scan_num_tiles(n);
  launch(scan_tile_sums_kernel, dim3((unsigned)tiles), dim3(SCAN_BLOCK), (size_t)(0), st, in, n, scratch);
  IA_LAUNCH_CHECK();
  launch(scan_tiles_kernel, dim3(1), dim3(SCAN_BLOCK), (size_t)(0), st, scratch, tiles, total_out);
  IA_LAUNCH_CHECK();
  launch(scan_apply_kernel, dim3((unsigned)tiles), dim3(SCAN_BLOCK), (size_t)(0), st, in, n, scratch, out);
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}

}  // namespace incagg
