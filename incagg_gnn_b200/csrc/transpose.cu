// CSR transpose (the CSC view the backward SpMM grad_X = A^T grad_out walks), sm_100a.
//
// torch_sparse builds this view with an argsort (csr2csc) for every new SparseTensor; here it is
// a counting sort: column histogram -> exclusive scan -> fill, then every transposed row is put
// into original-edge order so that the backward reduction order (and therefore its fp32 result)
// is the same on every run.
#include "common.cuh"
#include "scan.cuh"

namespace incagg {

constexpr int TR_THREADS = 256;

__global__ void col_hist_kernel(const int32_t* __restrict__ col, int64_t nnz, int64_t cols,
                                unsigned long long* __restrict__ hist) {
  pdl_prologue();
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int32_t c = col[e];
    if (c >= 0 && c < cols) atomicAdd(hist + c, 1ull);
  }
}

__global__ void write_trowptr_kernel(const int64_t* __restrict__ excl, const int64_t* __restrict__ total,
                                     int64_t cols, int32_t* __restrict__ t_rowptr,
                                     int32_t* __restrict__ cursor) {
  pdl_prologue();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cols) {
    t_rowptr[i] = (int32_t)excl[i];
    cursor[i] = (int32_t)excl[i];
  }
  if (i == cols) t_rowptr[cols] = (int32_t)(*total);
}

// One warp per source row: slot = cursor[c]++ ; store the edge position there.
__global__ void __launch_bounds__(TR_THREADS)
fill_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t rows,
            int64_t cols, int32_t* cursor, int32_t* __restrict__ t_perm) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (TR_THREADS / 32) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int s = rowptr[r], e = rowptr[r + 1];
  for (int j = s + lane; j < e; j += 32) {
    const int32_t c = col[j];
    if (c >= 0 && c < cols) {
      const int slot = atomicAdd(cursor + c, 1);
      t_perm[slot] = j;
    }
  }
}

// Put every transposed row into increasing edge-position order.  One warp per transposed row:
// rows up to 32 entries sort in registers (bitonic over shuffles); longer rows use a rank sort
// through shared memory in chunks (O(L^2 / 32) per row, rows are short on average and the cost
// is paid once per batch structure, which is cached).
__device__ __forceinline__ int warp_bitonic_sort(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool up = ((lane & k) == 0);
      const bool lower = ((lane & j) == 0);
      const int mn = min(v, o), mx = max(v, o);
      v = (lower == up) ? mn : mx;
    }
  }
  return v;
}

constexpr int MED_SORT = 1024;  // segments of 33 .. 1024 entries: small CTAs, 4 KB static smem

__global__ void __launch_bounds__(TR_THREADS)
sort_segments_kernel(const int32_t* __restrict__ t_rowptr, int64_t cols, int32_t* t_perm,
                     int32_t* __restrict__ long_rows, int32_t* long_count, int long_capacity) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * (TR_THREADS / 32) + (threadIdx.x >> 5);
  if (c >= cols) return;
  const int s = t_rowptr[c], e = t_rowptr[c + 1];
  const int L = e - s;
  if (L <= 1) return;
  if (L <= 32) {
    int v = (lane < L) ? t_perm[s + lane] : INT32_MAX;
    v = warp_bitonic_sort(v);
    if (lane < L) t_perm[s + lane] = v;
    return;
  }
  // two lists sharing one array: medium segments grow from the front, big ones from the back
  if (lane == 0) {
    if (L <= MED_SORT) {
      const int slot = atomicAdd(long_count, 1);
      long_rows[slot] = (int32_t)c;
    } else {
      const int slot = atomicAdd(long_count + 1, 1);
      long_rows[long_capacity - 1 - slot] = (int32_t)c;
    }
  }
}

// Long transposed rows: one CTA per row.  Uniform-direction bitonic network (every
// compare-exchange puts the smaller key at the lower index), so the virtual padding up to the
// next power of two behaves as +inf and never has to be stored.  Rows up to BIG_SORT entries are
// sorted in shared memory, longer ones in place in global memory (L2-resident).
constexpr int BIG_SORT = 32768;  // 128 KB of dynamic shared memory

template <typename Get, typename Put>
__device__ __forceinline__ void bitonic_uniform(int L, int P, Get get, Put put) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int partner = (j == (k >> 1)) ? (i ^ (k - 1)) : (i ^ j);
        if (partner > i && partner < L) {  // i < partner < L: both real keys
          const int a = get(i), b = get(partner);
          if (a > b) { put(i, b); put(partner, a); }
        }
      }
      __syncthreads();
    }
  }
}

// Medium segments (33 .. 1024 entries, the bulk of the deferred ones): 128-thread CTAs, 4 KB smem,
// many CTAs per SM.
constexpr int MED_THREADS = 128;
__global__ void __launch_bounds__(MED_THREADS)
sort_medium_segments_kernel(const int32_t* __restrict__ t_rowptr, int32_t* t_perm,
                            const int32_t* __restrict__ long_rows, const int32_t* __restrict__ long_count) {
  pdl_prologue();
  __shared__ int32_t sm[MED_SORT];
  const int n_med = long_count[0];
  for (int li = blockIdx.x; li < n_med; li += gridDim.x) {
    const int c = long_rows[li];
    const int s = t_rowptr[c], e = t_rowptr[c + 1];
    const int L = e - s;
    int P = 64;
    while (P < L) P <<= 1;
    int32_t* g = t_perm + s;
    for (int i = threadIdx.x; i < L; i += MED_THREADS) sm[i] = g[i];
    __syncthreads();
    bitonic_uniform(L, P, [&](int i) { return sm[i]; }, [&](int i, int v) { sm[i] = v; });
    for (int i = threadIdx.x; i < L; i += MED_THREADS) g[i] = sm[i];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(TR_THREADS)
sort_long_segments_kernel(const int32_t* __restrict__ t_rowptr, int32_t* t_perm,
                          const int32_t* __restrict__ long_rows, const int32_t* __restrict__ long_count,
                          int long_capacity) {
  pdl_prologue();
  extern __shared__ int32_t sm[];
  const int n_long = long_count[1];
  for (int li = blockIdx.x; li < n_long; li += gridDim.x) {
    const int c = long_rows[long_capacity - 1 - li];
    const int s = t_rowptr[c], e = t_rowptr[c + 1];
    const int L = e - s;
    int P = 64;
    while (P < L) P <<= 1;
    int32_t* g = t_perm + s;
    if (L <= BIG_SORT) {
      for (int i = threadIdx.x; i < L; i += TR_THREADS) sm[i] = g[i];
      __syncthreads();
      bitonic_uniform(L, P, [&](int i) { return sm[i]; }, [&](int i, int v) { sm[i] = v; });
      for (int i = threadIdx.x; i < L; i += TR_THREADS) g[i] = sm[i];
      __syncthreads();
    } else {
      bitonic_uniform(L, P, [&](int i) { return g[i]; }, [&](int i, int v) { g[i] = v; });
    }
  }
}

// t_col[q] = source row of edge t_perm[q]; t_val[q] = val[t_perm[q]].  The source row is found
// by binary search in rowptr (edge positions are increasing inside a transposed row, so the
// searches of neighbouring lanes touch neighbouring rowptr entries).
__global__ void gather_transposed_kernel(const int32_t* __restrict__ rowptr, int64_t rows,
                                         const float* __restrict__ val,
                                         const int32_t* __restrict__ t_perm, int64_t nnz,
                                         int32_t* __restrict__ t_col, float* __restrict__ t_val) {
  pdl_prologue();
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nnz;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int32_t ed = t_perm[q];
    int64_t lo = 0, hi = rows;  // largest r with rowptr[r] <= ed
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (rowptr[mid] <= ed) lo = mid; else hi = mid;
    }
    t_col[q] = (int32_t)lo;
    if (t_val) t_val[q] = val[ed];
  }
}

struct TransposeWs {
  int64_t* hist;      // [cols+1] histogram -> exclusive prefix
  int64_t* scan;      // scan scratch
  int64_t* total;     // [1]
  int32_t* cursor;    // [cols]
  int32_t* perm;      // [nnz] (used when the caller does not want t_perm)
  int32_t* long_cnt;  // [1]
  int32_t* long_rows; // [cols]
};

static size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

static TransposeWs carve(void* ws, int64_t cols, int64_t nnz) {
  TransposeWs w;
  char* p = static_cast<char*>(ws);
  w.hist = reinterpret_cast<int64_t*>(p);
  p += align16(sizeof(int64_t) * (size_t)(cols + 1));
  w.scan = reinterpret_cast<int64_t*>(p);
  p += align16(sizeof(int64_t) * (size_t)(scan_num_tiles(cols + 1) + 8));
  w.total = reinterpret_cast<int64_t*>(p);
  p += 16;
  w.cursor = reinterpret_cast<int32_t*>(p);
  p += align16(sizeof(int32_t) * (size_t)cols);
  w.perm = reinterpret_cast<int32_t*>(p);
  p += align16(sizeof(int32_t) * (size_t)nnz);
  w.long_cnt = reinterpret_cast<int32_t*>(p);
  p += 16;
  w.long_rows = reinterpret_cast<int32_t*>(p);
  return w;
}

}  // namespace incagg

using namespace incagg;

extern "C" size_t incagg_csr_transpose_workspace_bytes(int64_t rows, int64_t cols, int64_t nnz) {
  (void)rows;
  if (cols < 0 || nnz < 0) return 0;
  return align16(sizeof(int64_t) * (size_t)(cols + 1)) +
         align16(sizeof(int64_t) * (size_t)(scan_num_tiles(cols + 1) + 8)) + 16 +
         align16(sizeof(int32_t) * (size_t)cols) + align16(sizeof(int32_t) * (size_t)nnz) + 16 +
         align16(sizeof(int32_t) * (size_t)cols) + 64;
}

extern "C" int incagg_csr_transpose(const int32_t* rowptr, const int32_t* col, const float* val,
                                    int64_t rows, int64_t cols, int64_t nnz, int32_t* t_rowptr,
                                    int32_t* t_col, float* t_val, int32_t* t_perm, void* workspace,
                                    size_t workspace_bytes, incagg_stream_t stream) {
  IA_CHECK_ARG(rows >= 0 && cols >= 0 && nnz >= 0, "negative size");
  IA_CHECK_ARG(nnz < 0x7fffffff && cols < 0x7fffffff && rows < 0x7fffffff, "sizes exceed int32");
  IA_CHECK_ARG(t_rowptr != nullptr, "t_rowptr is NULL");
  IA_CHECK_ARG(workspace != nullptr &&
                   workspace_bytes >= incagg_csr_transpose_workspace_bytes(rows, cols, nnz),
               "workspace too small");
  IA_CHECK_ARG((val == nullptr) == (t_val == nullptr) || nnz == 0, "val / t_val must both be set or NULL");
  cudaStream_t st = as_stream(stream);
  TransposeWs w = carve(workspace, cols, nnz);
  IA_CUDA(cudaMemsetAsync(w.hist, 0, sizeof(int64_t) * (size_t)(cols + 1), st));
  if (nnz > 0) {
    IA_CHECK_ARG(rowptr && col && t_col, "NULL argument");
    const int64_t want = (nnz + 255) / 256;
    const int blocks = (int)(want < (int64_t)sm_count() * 16 ? want : (int64_t)sm_count() * 16);
    launch(col_hist_kernel, dim3(blocks), dim3(256), (size_t)(0), st, col, nnz, cols,
                                            reinterpret_cast<unsigned long long*>(w.hist));
    IA_LAUNCH_CHECK();
  }
  int rc = exclusive_scan_i64(w.hist, w.hist, cols, w.scan, w.total, st);
  if (rc != INCAGG_OK) return rc;
  launch(write_trowptr_kernel, dim3((unsigned)((cols + 1 + 255) / 256)), dim3(256), (size_t)(0), st, w.hist, w.total, cols,
                                                                           t_rowptr, w.cursor);
  IA_LAUNCH_CHECK();
  if (nnz == 0) return INCAGG_OK;
  int32_t* perm = t_perm ? t_perm : w.perm;
  const int wpb = TR_THREADS / 32;
  launch(fill_kernel, dim3((unsigned)((rows + wpb - 1) / wpb)), dim3(TR_THREADS), (size_t)(0), st, rowptr, col, rows, cols,
                                                                        w.cursor, perm);
  IA_LAUNCH_CHECK();
  IA_CUDA(cudaMemsetAsync(w.long_cnt, 0, 2 * sizeof(int32_t), st));
  if (cols > 0) {
    launch(sort_segments_kernel, dim3((unsigned)((cols + wpb - 1) / wpb)), dim3(TR_THREADS), (size_t)(0), st, 
        t_rowptr, cols, perm, w.long_rows, w.long_cnt, (int)cols);
    IA_LAUNCH_CHECK();
    static thread_local bool smem_set = false;
    if (!smem_set) {
      IA_CUDA(cudaFuncSetAttribute(sort_long_segments_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(BIG_SORT * sizeof(int32_t))));
      smem_set = true;
    }
    launch(sort_medium_segments_kernel, dim3(sm_count() * 12), dim3(MED_THREADS), (size_t)(0), st, t_rowptr, perm, w.long_rows, w.long_cnt);
    IA_LAUNCH_CHECK();
    launch(sort_long_segments_kernel, dim3(sm_count()), dim3(TR_THREADS), (size_t)(BIG_SORT * sizeof(int32_t)), st, 
        t_rowptr, perm, w.long_rows, w.long_cnt, (int)cols);
    IA_LAUNCH_CHECK();
  }
  {
    const int64_t want = (nnz + 255) / 256;
    const int blocks = (int)(want < (int64_t)sm_count() * 16 ? want : (int64_t)sm_count() * 16);
    launch(gather_transposed_kernel, dim3(blocks), dim3(256), (size_t)(0), st, rowptr, rows, val, perm, nnz, t_col, t_val);
    IA_LAUNCH_CHECK();
  }
  return INCAGG_OK;
}
