// History gather / scatter / slice copies (sm_100a).
//
// Replaces History.pull/push (history.py:33-65) and the copy loops of read_async / write_async
// (csrc/cuda/async_cuda.cu:61-111,139-163).  Pure data movement: every byte is read once and
// written once, so the design is vector width + requests in flight:
//   * rows are moved as 16/8/4-byte vectors (widest the alignment allows), consecutive lanes
//     take consecutive vectors of a row -> full 128 B sectors per warp request;
//   * each thread moves UNROLL vectors per trip, loads first and stores after, so UNROLL
//     independent 128-bit requests per lane are in flight (what hides HBM, and PCIe for
//     pinned-host sources, latency);
//   * the gather source may be pinned host memory: the kernel then reads it through UVA, which
//     replaces the reference's CPU index_select into a pinned bounce buffer + H2D copy;
//   * contiguous slices between pinned host memory and the device go through cudaMemcpyAsync
//     (DMA engines, overlappable with kernels); device<->device slices use one kernel launch
//     for up to 64 slices instead of one memcpy per slice.
#include <stdlib.h>

#include "common.cuh"

namespace incagg {

template <int VB> struct Bytes;
template <> struct Bytes<16> { using type = uint4; };
template <> struct Bytes<8> { using type = uint2; };
template <> struct Bytes<4> { using type = uint32_t; };
template <> struct Bytes<2> { using type = uint16_t; };
template <> struct Bytes<1> { using type = uint8_t; };  // label / mask columns (bool, int8)

constexpr int ROWS_THREADS = 256;
constexpr int ROWS_UNROLL = 4;

// mode 0: dst[i] = src[idx[i]]   mode 1: dst[idx[i]] = src[i]
template <int VB, int MODE>
__global__ void __launch_bounds__(ROWS_THREADS)
index_rows_kernel(const char* __restrict__ src, int64_t src_ld, const int64_t* __restrict__ idx,
                  int64_t n, char* __restrict__ dst, int64_t dst_ld, int nvec, int64_t limit_rows,
                  int32_t* __restrict__ err) {
  pdl_prologue();
  using V = typename Bytes<VB>::type;
  const int64_t total = n * nvec;
  const int64_t stride = (int64_t)gridDim.x * ROWS_THREADS;
  int64_t t = (int64_t)blockIdx.x * ROWS_THREADS + threadIdx.x;
  for (; t < total; t += stride * ROWS_UNROLL) {
    V v[ROWS_UNROLL];
    int64_t doff[ROWS_UNROLL];
#pragma unroll
    for (int u = 0; u < ROWS_UNROLL; ++u) {
      const int64_t g = t + u * stride;
      doff[u] = -1;
      if (g < total) {
        const int64_t r = g / nvec;
        const int c = (int)(g - r * nvec);
        const int64_t j = idx[r];
        if (j >= 0 && j < limit_rows) {
          const int64_t srow = (MODE == 0) ? j : r;
          const int64_t drow = (MODE == 0) ? r : j;
          v[u] = *reinterpret_cast<const V*>(src + srow * src_ld + (int64_t)c * VB);
          doff[u] = drow * dst_ld + (int64_t)c * VB;
        } else {  // the reference raises here: flag it; a gather returns a zero row, a scatter skips it
          if (c == 0 && err) atomicOr(err, INCAGG_DEVERR_ROW_INDEX);
          if (MODE == 0) {
            v[u] = V{};
            doff[u] = r * dst_ld + (int64_t)c * VB;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < ROWS_UNROLL; ++u)
      if (doff[u] >= 0) *reinterpret_cast<V*>(dst + doff[u]) = v[u];
  }
}

// Row gather out of a table that is sharded by contiguous row ranges over up to 16 memories (the
// history shards of all ranks: the local one and the peers' HBM mapped over NVLink through CUDA IPC).
// dst[i] = shard[s][idx[i] - bounds[s]] with bounds[s] <= idx[i] < bounds[s+1]: the remote rows are
// fetched by plain loads over NVLink inside the same kernel that packs them.
constexpr int MAX_SHARDS = 16;
struct ShardTable {
  const char* base[MAX_SHARDS];
  int64_t bounds[MAX_SHARDS + 1];
  int n;
};

template <int VB>
__global__ void __launch_bounds__(ROWS_THREADS)
sharded_gather_kernel(const ShardTable tab, int64_t src_ld, const int64_t* __restrict__ idx, int64_t n,
                      char* __restrict__ dst, int64_t dst_ld, int nvec, int32_t* __restrict__ err) {
  pdl_prologue();
  using V = typename Bytes<VB>::type;
  const int64_t total = n * nvec;
  const int64_t stride = (int64_t)gridDim.x * ROWS_THREADS;
  int64_t t = (int64_t)blockIdx.x * ROWS_THREADS + threadIdx.x;
  for (; t < total; t += stride * ROWS_UNROLL) {
    V v[ROWS_UNROLL];
    int64_t doff[ROWS_UNROLL];
#pragma unroll
    for (int u = 0; u < ROWS_UNROLL; ++u) {
      const int64_t g = t + u * stride;
      doff[u] = -1;
      if (g < total) {
        const int64_t r = g / nvec;
        const int c = (int)(g - r * nvec);
        const int64_t j = idx[r];
        if (j >= tab.bounds[0] && j < tab.bounds[tab.n]) {
          int s = 0;
          while (j >= tab.bounds[s + 1]) ++s;
          v[u] = *reinterpret_cast<const V*>(tab.base[s] + (j - tab.bounds[s]) * src_ld + (int64_t)c * VB);
          doff[u] = r * dst_ld + (int64_t)c * VB;
        } else {  // id owned by no shard: zero row + error flag
          if (c == 0 && err) atomicOr(err, INCAGG_DEVERR_ROW_INDEX);
          v[u] = V{};
          doff[u] = r * dst_ld + (int64_t)c * VB;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < ROWS_UNROLL; ++u)
      if (doff[u] >= 0) *reinterpret_cast<V*>(dst + doff[u]) = v[u];
  }
}

constexpr int MAX_SLICES = 64;
struct SliceTable {
  int64_t off[MAX_SLICES];          // row offset in the strided (history) tensor
  int64_t prefix[MAX_SLICES + 1];   // row offset in the packed tensor
  int k;
};

// direction 0: packed dst <- strided src slices; direction 1: strided dst slices <- packed src
template <int VB>
__global__ void __launch_bounds__(ROWS_THREADS)
slice_rows_kernel(const char* __restrict__ src, int64_t src_ld, char* __restrict__ dst,
                  int64_t dst_ld, int nvec, int direction, const SliceTable tab) {
  pdl_prologue();
  using V = typename Bytes<VB>::type;
  const int64_t total = tab.prefix[tab.k] * nvec;
  const int64_t stride = (int64_t)gridDim.x * ROWS_THREADS;
  int64_t t = (int64_t)blockIdx.x * ROWS_THREADS + threadIdx.x;
  for (; t < total; t += stride * ROWS_UNROLL) {
    V v[ROWS_UNROLL];
    int64_t doff[ROWS_UNROLL];
#pragma unroll
    for (int u = 0; u < ROWS_UNROLL; ++u) {
      const int64_t g = t + u * stride;
      doff[u] = -1;
      if (g < total) {
        const int64_t r = g / nvec;  // packed row
        const int c = (int)(g - r * nvec);
        int lo = 0, hi = tab.k;      // slice i with prefix[i] <= r < prefix[i+1]
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (tab.prefix[mid] <= r) lo = mid; else hi = mid;
        }
        const int64_t strided = tab.off[lo] + (r - tab.prefix[lo]);
        const int64_t srow = direction == 0 ? strided : r;
        const int64_t drow = direction == 0 ? r : strided;
        v[u] = *reinterpret_cast<const V*>(src + srow * src_ld + (int64_t)c * VB);
        doff[u] = drow * dst_ld + (int64_t)c * VB;
      }
    }
#pragma unroll
    for (int u = 0; u < ROWS_UNROLL; ++u)
      if (doff[u] >= 0) *reinterpret_cast<V*>(dst + doff[u]) = v[u];
  }
}

// ---- TMA bulk slice copies ------------------------------------------------------------------------
// A partition's rows are one contiguous byte range on both sides when the tables are dense
// (ld == row bytes): the push / pull of a slice is then a plain byte-range copy, which the TMA engine
// does without occupying load/store units: cp.async.bulk global -> shared (completion counted on an
// mbarrier) and cp.async.bulk shared -> global (bulk groups), a ring of BULK_STAGES chunks per CTA,
// one elected thread issues everything.  Needs 16-byte aligned addresses and sizes.
constexpr int BULK_CHUNK = 16384;
constexpr int BULK_STAGES = 4;

struct BulkTable {
  int64_t src_off[MAX_SLICES];      // byte offset of each slice in src
  int64_t dst_off[MAX_SLICES];      // byte offset of each slice in dst
  int64_t chunk_prefix[MAX_SLICES + 1];  // chunks before slice i
  int64_t bytes[MAX_SLICES];
  int k;
};

__device__ __forceinline__ uint32_t rows_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__global__ void __launch_bounds__(32)
slice_bulk_kernel(const char* __restrict__ src, char* __restrict__ dst, const BulkTable tab) {
  pdl_prologue();
  extern __shared__ __align__(128) char bulk_smem[];
  __shared__ uint64_t bar[BULK_STAGES];
  if (threadIdx.x != 0) return;  // one thread drives the TMA engine
  for (int s = 0; s < BULK_STAGES; ++s)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(rows_smem_u32(&bar[s])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const int64_t n_chunks = tab.chunk_prefix[tab.k];
  const int64_t first = blockIdx.x, step = gridDim.x;
  // chunk index -> (src pointer, dst pointer, bytes)
  auto locate = [&](int64_t c, const char*& sp, char*& dp, uint32_t& nb) {
    int lo = 0, hi = tab.k;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (tab.chunk_prefix[mid] <= c) lo = mid; else hi = mid;
    }
    const int64_t within = (c - tab.chunk_prefix[lo]) * BULK_CHUNK;
    const int64_t left = tab.bytes[lo] - within;
    nb = (uint32_t)(left < BULK_CHUNK ? left : BULK_CHUNK);
    sp = src + tab.src_off[lo] + within;
    dp = dst + tab.dst_off[lo] + within;
  };
  auto issue_load = [&](int64_t c, int stage) {
    const char* sp; char* dp; uint32_t nb;
    locate(c, sp, dp, nb);
    const uint32_t b = rows_smem_u32(&bar[stage]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(nb) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(rows_smem_u32(bulk_smem + (size_t)stage * BULK_CHUNK)), "l"(sp), "r"(nb), "r"(b)
                 : "memory");
  };
  // prologue: fill the ring
  int64_t next = first;
  for (int s = 0; s < BULK_STAGES && next < n_chunks; ++s, next += step) issue_load(next, s);
  int it = 0;
  for (int64_t c = first; c < n_chunks; c += step, ++it) {
    const int stage = it % BULK_STAGES;
    const uint32_t parity = (uint32_t)((it / BULK_STAGES) & 1);
    const uint32_t b = rows_smem_u32(&bar[stage]);
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_LOAD:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_LOAD;\n\tbra WAIT_LOAD;\n\tDONE_LOAD:\n\t}" ::"r"(b), "r"(parity) : "memory");
    const char* sp; char* dp; uint32_t nb;
    locate(c, sp, dp, nb);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dp), "r"(rows_smem_u32(bulk_smem + (size_t)stage * BULK_CHUNK)), "r"(nb) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (next < n_chunks) {
      // the stage is reloaded only after its store has finished READING shared memory
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      issue_load(next, stage);
      next += step;
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores complete before exit
}

static bool is_host_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;  // unregistered pageable memory
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeUnregistered;
}

// A gather whose source is pinned host memory (zero-copy over PCIe) is bound by the link, not by
// the SMs: ~110 KB in flight saturate it.  A full-size grid would park 2 K threads on every SM for
// the whole transfer and keep the training kernels of the concurrent step off the GPU, so such
// launches use a few CTAs only (INCAGG_HOST_GATHER_CTAS, default 48: 0.8 MB in flight).
static int host_gather_ctas() {
  static const int v = getenv("INCAGG_HOST_GATHER_CTAS") ? atoi(getenv("INCAGG_HOST_GATHER_CTAS")) : 48;
  return v < 1 ? 1 : v;
}

static int pick_vb(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t row_bytes) {
  auto ok = [&](int vb) {
    return row_bytes % vb == 0 && lda % vb == 0 && ldb % vb == 0 &&
           reinterpret_cast<uintptr_t>(a) % vb == 0 && reinterpret_cast<uintptr_t>(b) % vb == 0;
  };
  return ok(16) ? 16 : (ok(8) ? 8 : (ok(4) ? 4 : (ok(2) ? 2 : 1)));
}

static int grid_for(int64_t total_vec) {
  const int64_t per_block = (int64_t)ROWS_THREADS * ROWS_UNROLL;
  int64_t want = (total_vec + per_block - 1) / per_block;
  const int64_t cap = (int64_t)sm_count() * 16;  // 16 CTAs of 256 threads = 2 full waves / SM
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

template <int MODE>
static int index_rows(const void* src, int64_t src_ld, const int64_t* idx, int64_t n, void* dst,
                      int64_t dst_ld, int64_t row_bytes, int64_t limit_rows, cudaStream_t st) {
  IA_CHECK_ARG(n >= 0 && row_bytes >= 0, "negative size");
  if (n == 0 || row_bytes == 0) return INCAGG_OK;
  IA_CHECK_ARG(src && dst && idx, "NULL argument");
  IA_CHECK_ARG(src_ld >= row_bytes && dst_ld >= row_bytes, "leading dimension smaller than a row");
  const int vb = pick_vb(src, src_ld, dst, dst_ld, row_bytes);
  const int64_t nvec64 = row_bytes / vb;
  IA_CHECK_ARG(nvec64 <= 0x7fffffff, "row too wide");
  const int nvec = (int)nvec64;
  int grid = grid_for(n * nvec64);
  if (is_host_ptr(MODE == 0 ? src : (const void*)dst) && grid > host_gather_ctas()) grid = host_gather_ctas();
  const char* s = static_cast<const char*>(src);
  char* d = static_cast<char*>(dst);
  int32_t* err = device_error_word();
  if (vb == 16)
    launch(index_rows_kernel<16, MODE>, dim3(grid), dim3(ROWS_THREADS), (size_t)(0), st, s, src_ld, idx, n, d, dst_ld, nvec, limit_rows, err);
  else if (vb == 8)
    launch(index_rows_kernel<8, MODE>, dim3(grid), dim3(ROWS_THREADS), (size_t)(0), st, s, src_ld, idx, n, d, dst_ld, nvec, limit_rows, err);
  else if (vb == 4)
    launch(index_rows_kernel<4, MODE>, dim3(grid), dim3(ROWS_THREADS), (size_t)(0), st, s, src_ld, idx, n, d, dst_ld, nvec, limit_rows, err);
  else if (vb == 2)
    launch(index_rows_kernel<2, MODE>, dim3(grid), dim3(ROWS_THREADS), (size_t)(0), st, s, src_ld, idx, n, d, dst_ld, nvec, limit_rows, err);
  else
    launch(index_rows_kernel<1, MODE>, dim3(grid), dim3(ROWS_THREADS), (size_t)(0), st, s, src_ld, idx, n, d, dst_ld, nvec, limit_rows, err);
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}

}  // namespace incagg

using namespace incagg;

extern "C" int incagg_gather_rows(const void* src, int64_t src_ld_bytes, int64_t src_rows,
                                  const int64_t* idx, int64_t n, void* dst, int64_t dst_ld_bytes,
                                  int64_t row_bytes, incagg_stream_t stream) {
  return index_rows<0>(src, src_ld_bytes, idx, n, dst, dst_ld_bytes, row_bytes, src_rows,
                       as_stream(stream));
}

extern "C" int incagg_gather_rows_sharded(const void* const* shard_ptrs, const int64_t* bounds, int num_shards,
                                          int64_t src_ld_bytes, const int64_t* idx, int64_t n, void* dst,
                                          int64_t dst_ld_bytes, int64_t row_bytes, incagg_stream_t stream) {
  IA_CHECK_ARG(num_shards >= 1 && num_shards <= MAX_SHARDS, "num_shards must be in [1, %d]", MAX_SHARDS);
  IA_CHECK_ARG(n >= 0 && row_bytes >= 0, "negative size");
  if (n == 0 || row_bytes == 0) return INCAGG_OK;
  IA_CHECK_ARG(shard_ptrs && bounds && idx && dst, "NULL argument");
  IA_CHECK_ARG(src_ld_bytes >= row_bytes && dst_ld_bytes >= row_bytes, "leading dimension smaller than a row");
  ShardTable tab;
  tab.n = num_shards;
  int vb = 16;
  for (int s = 0; s < num_shards; ++s) {
    IA_CHECK_ARG(bounds[s] <= bounds[s + 1], "bounds must be non-decreasing");
    IA_CHECK_ARG(shard_ptrs[s] != nullptr || bounds[s] == bounds[s + 1], "shard %d is NULL", s);
    tab.base[s] = static_cast<const char*>(shard_ptrs[s]);
    tab.bounds[s] = bounds[s];
    if (shard_ptrs[s]) {
      const int v = pick_vb(shard_ptrs[s], src_ld_bytes, dst, dst_ld_bytes, row_bytes);
      if (v < vb) vb = v;
    }
  }
  tab.bounds[num_shards] = bounds[num_shards];
  const int64_t nvec64 = row_bytes / vb;
  IA_CHECK_ARG(nvec64 <= 0x7fffffff, "row too wide");
  const int nvec = (int)nvec64;
  const int grid = grid_for(n * nvec64);
  cudaStream_t st = as_stream(stream);
  char* d = static_cast<char*>(dst);
#define IA_SG(VB_) launch(sharded_gather_kernel<VB_>, dim3(grid), dim3(ROWS_THREADS), (size_t)(0), st, tab, src_ld_bytes, idx, n, d, dst_ld_bytes, nvec, device_error_word())
  if (vb == 16) IA_SG(16); else if (vb == 8) IA_SG(8); else if (vb == 4) IA_SG(4); else if (vb == 2) IA_SG(2); else IA_SG(1);
#undef IA_SG
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}

extern "C" int incagg_scatter_rows(const void* src, int64_t src_ld_bytes, const int64_t* idx,
                                   int64_t n, void* dst, int64_t dst_ld_bytes, int64_t dst_rows,
                                   int64_t row_bytes, incagg_stream_t stream) {
  return index_rows<1>(src, src_ld_bytes, idx, n, dst, dst_ld_bytes, row_bytes, dst_rows,
                       as_stream(stream));
}

extern "C" int incagg_copy_slices(const void* src, int64_t src_ld_bytes, int64_t src_rows,
                                  void* dst, int64_t dst_ld_bytes, int64_t dst_rows,
                                  const int64_t* offset, const int64_t* count, int64_t k,
                                  int64_t row_bytes, int direction, incagg_stream_t stream) {
  IA_CHECK_ARG(k >= 0 && row_bytes >= 0, "negative size");
  IA_CHECK_ARG(direction == 0 || direction == 1, "direction must be 0 (pull) or 1 (push)");
  if (k == 0 || row_bytes == 0) return INCAGG_OK;
  IA_CHECK_ARG(src && dst && offset && count, "NULL argument");
  IA_CHECK_ARG(src_ld_bytes >= row_bytes && dst_ld_bytes >= row_bytes,
               "leading dimension smaller than a row");
  // Bounds, as the reference asserts per slice ("Invalid index", async_cuda.cu:78-79,148-149).
  int64_t packed = 0;
  for (int64_t i = 0; i < k; ++i) {
    const int64_t o = offset[i], c = count[i];
    IA_CHECK_ARG(o >= 0 && c >= 0, "Invalid index (negative offset/count)");
    const int64_t strided_rows = direction == 0 ? src_rows : dst_rows;
    const int64_t packed_rows = direction == 0 ? dst_rows : src_rows;
    IA_CHECK_ARG(o + c <= strided_rows, "Invalid index (slice %lld: %lld+%lld > %lld rows)",
                 (long long)i, (long long)o, (long long)c, (long long)strided_rows);
    IA_CHECK_ARG(packed + c <= packed_rows, "Invalid index (packed side too small)");
    packed += c;
  }
  if (packed == 0) return INCAGG_OK;
  cudaStream_t st = as_stream(stream);
  const char* s = static_cast<const char*>(src);
  char* d = static_cast<char*>(dst);
  if (is_host_ptr(src) || is_host_ptr(dst)) {
    // Host <-> device: one DMA copy per slice on `stream` (2-D when a side is strided wider
    // than the row).  cudaMemcpyDefault lets the driver infer the direction from UVA.
    int64_t p = 0;
    for (int64_t i = 0; i < k; ++i) {
      const int64_t o = offset[i], c = count[i];
      if (c == 0) continue;
      const char* sp = direction == 0 ? s + o * src_ld_bytes : s + p * src_ld_bytes;
      char* dp = direction == 0 ? d + p * dst_ld_bytes : d + o * dst_ld_bytes;
      if (src_ld_bytes == row_bytes && dst_ld_bytes == row_bytes) {
        IA_CUDA(cudaMemcpyAsync(dp, sp, (size_t)(c * row_bytes), cudaMemcpyDefault, st));
      } else {
        IA_CUDA(cudaMemcpy2DAsync(dp, (size_t)dst_ld_bytes, sp, (size_t)src_ld_bytes,
                                  (size_t)row_bytes, (size_t)c, cudaMemcpyDefault, st));
      }
      p += c;
    }
    return INCAGG_OK;
  }
  // Device <-> device, dense rows on both sides, 16-byte aligned: TMA bulk copies.
  static const bool use_bulk = []() { const char* e = getenv("INCAGG_SLICE_BULK"); return !(e && e[0] == '0'); }();
  if (use_bulk && src_ld_bytes == row_bytes && dst_ld_bytes == row_bytes && row_bytes % 16 == 0 &&
      (reinterpret_cast<uintptr_t>(src) % 16) == 0 && (reinterpret_cast<uintptr_t>(dst) % 16) == 0 &&
      packed * row_bytes >= 4 * BULK_CHUNK) {
    static bool attr_set[16] = {false};
    if (first_use_on_device(attr_set)) {
      IA_CUDA(cudaFuncSetAttribute(slice_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   BULK_STAGES * BULK_CHUNK));
    }
    int64_t pk = 0;
    for (int64_t i0 = 0; i0 < k; i0 += MAX_SLICES) {
      BulkTable tab;
      const int kk = (int)((k - i0) < MAX_SLICES ? (k - i0) : MAX_SLICES);
      tab.k = kk;
      tab.chunk_prefix[0] = 0;
      for (int i = 0; i < kk; ++i) {
        const int64_t o = offset[i0 + i] * row_bytes, c = count[i0 + i] * row_bytes, pb = pk * row_bytes;
        tab.src_off[i] = direction == 0 ? o : pb;
        tab.dst_off[i] = direction == 0 ? pb : o;
        tab.bytes[i] = c;
        tab.chunk_prefix[i + 1] = tab.chunk_prefix[i] + (c + BULK_CHUNK - 1) / BULK_CHUNK;
        pk += count[i0 + i];
      }
      const int64_t n_chunks = tab.chunk_prefix[kk];
      if (n_chunks == 0) continue;
      int64_t grid = (int64_t)sm_count() * 3;  // 64 KB of shared memory per CTA -> 3 CTAs per SM
      if (grid > n_chunks) grid = n_chunks;
      launch(slice_bulk_kernel, dim3((unsigned)grid), dim3(32), (size_t)(BULK_STAGES * BULK_CHUNK), st, s, d, tab);
      IA_LAUNCH_CHECK();
    }
    return INCAGG_OK;
  }
  // Device <-> device: up to MAX_SLICES slices per launch.
  const int vb = pick_vb(src, src_ld_bytes, dst, dst_ld_bytes, row_bytes);
  const int nvec = (int)(row_bytes / vb);
  int64_t p = 0;
  for (int64_t i0 = 0; i0 < k; i0 += MAX_SLICES) {
    SliceTable tab;
    const int kk = (int)((k - i0) < MAX_SLICES ? (k - i0) : MAX_SLICES);
    tab.k = kk;
    tab.prefix[0] = 0;
    for (int i = 0; i < kk; ++i) {
      tab.off[i] = offset[i0 + i];
      tab.prefix[i + 1] = tab.prefix[i] + count[i0 + i];
    }
    const int64_t rows_here = tab.prefix[kk];
    if (rows_here > 0) {
      // the packed side starts at packed row p for this group of slices
      const char* sp = direction == 0 ? s : s + p * src_ld_bytes;
      char* dp = direction == 0 ? d + p * dst_ld_bytes : d;
      const int grid = grid_for(rows_here * nvec);
      if (vb == 16)
        launch(slice_rows_kernel<16>, dim3(grid), dim3(ROWS_THREADS), (size_t)(0), st, sp, src_ld_bytes, dp, dst_ld_bytes, nvec, direction, tab);
      else if (vb == 8)
        launch(slice_rows_kernel<8>, dim3(grid), dim3(ROWS_THREADS), (size_t)(0), st, sp, src_ld_bytes, dp, dst_ld_bytes, nvec, direction, tab);
      else if (vb == 4)
        launch(slice_rows_kernel<4>, dim3(grid), dim3(ROWS_THREADS), (size_t)(0), st, sp, src_ld_bytes, dp, dst_ld_bytes, nvec, direction, tab);
      else if (vb == 2)
        launch(slice_rows_kernel<2>, dim3(grid), dim3(ROWS_THREADS), (size_t)(0), st, sp, src_ld_bytes, dp, dst_ld_bytes, nvec, direction, tab);
      else
        launch(slice_rows_kernel<1>, dim3(grid), dim3(ROWS_THREADS), (size_t)(0), st, sp, src_ld_bytes, dp, dst_ld_bytes, nvec, direction, tab);
      IA_LAUNCH_CHECK();
    }
    p += rows_here;
  }
  return INCAGG_OK;
}
