// GPU relabel_one_hop / relabel_one_hop_within_batch (sm_100a), bit-exact with the reference's
// sequential CPU implementation (csrc/cpu/relabel_cpu.cpp:3-108 and :111-214).
//
// The reference walks the batch rows in order and every row's columns in CSR order, giving each
// previously unseen column the next local id (std::unordered_map, one thread).  The parallel
// formulation here reproduces that order exactly:
//   1. mark      map[idx[i]] = max over duplicates of i      (reference: later duplicate overwrites)
//   2. degrees   -> exclusive scan -> out_rowptr             (edge position p of every copied edge)
//   3. map_edges in-batch columns get map[w]; halo columns do atomicMin(first_pos[w], p)
//   4. an edge is "first" iff first_pos[w] == p; per-row counts of first edges are scanned and,
//      inside a row, ranked with warp ballots in CSR order -> halo id = B + rank, n_id[B+rank] = w
//   5. halo columns are rewritten with their id; the touched table entries are reset.
// Instead of a hash map the lookup structure is a direct-address table over global node ids
// (2 x int32 per node: 20 MB for ogbn-products, 0.01 % of a B200's HBM), allocated once and left
// clean by every call, so a lookup is one 4-byte load and there are no collisions or probes.
// All traffic is HBM/L2-bound integer work; one warp owns one batch row so that index loads are
// coalesced along the row.
#include <limits.h>

#include "common.cuh"
#include "scan.cuh"

namespace incagg {

constexpr int RL_THREADS = 256;
constexpr int RL_WARPS = RL_THREADS / 32;

struct RelabelWs {
  int32_t* map;        // [N]  local id of a global node, -1 if absent
  int32_t* first_pos;  // [N]  smallest edge position that references a halo node
  int64_t* rowtmp;     // [N+1] per-row degrees -> exclusive prefix (edge position of each row)
  int64_t* rowtmp2;    // [N+1] per-row counts of first-seen halo edges -> exclusive prefix
  int64_t* scan;       // [scan tiles] scratch of exclusive_scan_i64
  int64_t* total;      // [2]
  int64_t* ids;        // [N]  the batch ids with ids outside [0, N) replaced by -1 (see mark_kernel)
};

static size_t ws_scan_slots(int64_t n) { return (size_t)scan_num_tiles(n + 1) + 8; }

static RelabelWs carve_ws(void* ws, int64_t n) {
  RelabelWs w;
  char* p = static_cast<char*>(ws);
  w.map = reinterpret_cast<int32_t*>(p);
  p += sizeof(int32_t) * (size_t)n;
  w.first_pos = reinterpret_cast<int32_t*>(p);
  p += sizeof(int32_t) * (size_t)n;
  p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));
  w.rowtmp = reinterpret_cast<int64_t*>(p);
  p += sizeof(int64_t) * (size_t)(n + 1);
  w.rowtmp2 = reinterpret_cast<int64_t*>(p);
  p += sizeof(int64_t) * (size_t)(n + 1);
  w.scan = reinterpret_cast<int64_t*>(p);
  p += sizeof(int64_t) * ws_scan_slots(n);
  w.total = reinterpret_cast<int64_t*>(p);
  p += sizeof(int64_t) * 8;
  w.ids = reinterpret_cast<int64_t*>(p);
  return w;
}

__global__ void ws_init_kernel(int32_t* map, int32_t* first_pos, int64_t n) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    map[i] = -1;
    first_pos[i] = INT_MAX;
  }
}

// map[idx[i]] = i (largest i wins) and deg[i] = degree of idx[i].
__global__ void mark_kernel(const int64_t* __restrict__ rowptr, const int64_t* __restrict__ idx,
                            int64_t B, int32_t* map, int64_t* __restrict__ deg, int64_t num_nodes,
                            int32_t* __restrict__ err, int64_t* __restrict__ ids) {
  pdl_prologue();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const int64_t v = idx[i];
  if (v < 0 || v >= num_nodes) {
    // the reference would read out of bounds here: set the error bit and treat the id as a node
    // without edges (the later kernels read `ids`, where it is -1)
    if (err) atomicOr(err, INCAGG_DEVERR_NODE_ID);
    if (deg) deg[i] = 0;
    ids[i] = -1;
    return;
  }
  ids[i] = v;
  atomicMax(map + v, (int32_t)i);
  if (deg) deg[i] = rowptr[v + 1] - rowptr[v];
}

__global__ void degree_sum_kernel(const int64_t* __restrict__ rowptr, const int64_t* __restrict__ idx,
                                  int64_t B, unsigned long long* out, int64_t num_nodes) {
  pdl_prologue();
  int64_t s = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = idx[i];
    if (v >= 0 && v < num_nodes) s += rowptr[v + 1] - rowptr[v];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0 && s != 0) atomicAdd(out, (unsigned long long)s);
}

// out_rowptr[i] = excl[i] for i < B, out_rowptr[B] = total
template <typename OT>
__global__ void write_rowptr_kernel(const int64_t* __restrict__ excl, const int64_t* __restrict__ total,
                                    int64_t B, OT* __restrict__ out_rowptr) {
  pdl_prologue();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) out_rowptr[i] = (OT)excl[i];
  if (i == B) out_rowptr[B] = (OT)(*total);
}

// One warp per batch row.  In-batch columns are final here; halo columns get -1 and race for
// the smallest edge position.  Edge values are copied.
template <typename CT, typename OT>
__global__ void __launch_bounds__(RL_THREADS)
map_edges_kernel(const int64_t* __restrict__ rowptr, const CT* __restrict__ col,
                 const float* __restrict__ val, const int64_t* __restrict__ idx, int64_t B,
                 const int64_t* __restrict__ row_start, const int32_t* __restrict__ map,
                 int32_t* first_pos, OT* __restrict__ out_col, float* __restrict__ out_val) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * RL_WARPS + (threadIdx.x >> 5);
  if (i >= B) return;
  const int64_t v = idx[i];
  if (v < 0) return;  // id outside the graph (flagged by mark_kernel)
  const int64_t s = rowptr[v], e = rowptr[v + 1];
  const int64_t p0 = row_start[i];
  for (int64_t j = s + lane; j < e; j += 32) {
    const int64_t w = (int64_t)col[j];
    const int64_t p = p0 + (j - s);
    const int32_t m = map[w];
    if (m >= 0) {
      out_col[p] = (OT)m;
    } else {
      out_col[p] = (OT)-1;
      atomicMin(first_pos + w, (int32_t)p);
    }
    if (val) out_val[p] = val[j];
  }
}

// Per row: number of edges that are the first reference to a halo node.
template <typename CT, typename OT>
__global__ void __launch_bounds__(RL_THREADS)
count_first_kernel(const int64_t* __restrict__ rowptr, const CT* __restrict__ col,
                   const int64_t* __restrict__ idx, int64_t B, const int64_t* __restrict__ row_start,
                   const int32_t* __restrict__ first_pos, const OT* __restrict__ out_col,
                   int64_t* __restrict__ row_first) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * RL_WARPS + (threadIdx.x >> 5);
  if (i >= B) return;
  const int64_t v = idx[i];
  if (v < 0) {  // id outside the graph (flagged by mark_kernel): no edges
    if (lane == 0) row_first[i] = 0;
    return;
  }
  const int64_t s = rowptr[v], e = rowptr[v + 1];
  const int64_t p0 = row_start[i];
  int cnt = 0;
  for (int64_t j = s + lane; j < e; j += 32) {
    const int64_t p = p0 + (j - s);
    if (out_col[p] == (OT)-1 && first_pos[(int64_t)col[j]] == (int32_t)p) ++cnt;
  }
  cnt = warp_sum(cnt);
  if (lane == 0) row_first[i] = cnt;
}

// Per row, in CSR order: first edges receive halo id B + (prefix of earlier rows) + (rank in row).
template <typename CT, typename OT>
__global__ void __launch_bounds__(RL_THREADS)
assign_halo_kernel(const int64_t* __restrict__ rowptr, const CT* __restrict__ col,
                   const int64_t* __restrict__ idx, int64_t B, const int64_t* __restrict__ row_start,
                   const int64_t* __restrict__ row_first_excl, const int32_t* __restrict__ first_pos,
                   const OT* __restrict__ out_col, int32_t* map, int64_t* __restrict__ n_id_out) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * RL_WARPS + (threadIdx.x >> 5);
  if (i >= B) return;
  const int64_t v = idx[i];
  if (v < 0) return;  // id outside the graph (flagged by mark_kernel)
  const int64_t s = rowptr[v], e = rowptr[v + 1];
  const int64_t p0 = row_start[i];
  int64_t rank0 = row_first_excl[i];
  for (int64_t jb = s; jb < e; jb += 32) {
    const int64_t j = jb + lane;
    bool first = false;
    int64_t w = 0;
    if (j < e) {
      const int64_t p = p0 + (j - s);
      w = (int64_t)col[j];
      first = (out_col[p] == (OT)-1) && (first_pos[w] == (int32_t)p);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, first);
    if (first) {
      const int64_t hid = B + rank0 + __popc(bal & ((1u << lane) - 1u));
      n_id_out[hid] = w;
      map[w] = (int32_t)hid;
    }
    rank0 += __popc(bal);
  }
}

// Halo columns (-1) take their final id.
template <typename CT, typename OT>
__global__ void __launch_bounds__(RL_THREADS)
fill_halo_cols_kernel(const int64_t* __restrict__ rowptr, const CT* __restrict__ col,
                      const int64_t* __restrict__ idx, int64_t B,
                      const int64_t* __restrict__ row_start, const int32_t* __restrict__ map,
                      OT* __restrict__ out_col) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * RL_WARPS + (threadIdx.x >> 5);
  if (i >= B) return;
  const int64_t v = idx[i];
  if (v < 0) return;  // id outside the graph (flagged by mark_kernel)
  const int64_t s = rowptr[v], e = rowptr[v + 1];
  const int64_t p0 = row_start[i];
  for (int64_t j = s + lane; j < e; j += 32) {
    const int64_t p = p0 + (j - s);
    if (out_col[p] == (OT)-1) out_col[p] = (OT)map[(int64_t)col[j]];
  }
}

// n_id_out[0:B] = idx; reset table entries of batch nodes and (once H is known on the device)
// of halo nodes.
__global__ void finish_kernel(const int64_t* __restrict__ idx, int64_t B,
                              const int64_t* __restrict__ H_dev, int64_t* __restrict__ n_id_out,
                              int32_t* map, int32_t* first_pos, int64_t num_nodes) {
  pdl_prologue();
  const int64_t H = H_dev ? *H_dev : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B + H;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (i < B) {
      const int64_t v = idx[i];
      if (n_id_out) n_id_out[i] = v;
      if (v >= 0 && v < num_nodes) map[v] = -1;
    } else {
      const int64_t v = n_id_out[i];
      map[v] = -1;
      first_pos[v] = INT_MAX;
    }
  }
}

// within-batch: per-row count of kept edges
template <typename CT>
__global__ void __launch_bounds__(RL_THREADS)
count_kept_kernel(const int64_t* __restrict__ rowptr, const CT* __restrict__ col,
                  const int64_t* __restrict__ idx, int64_t B, const int32_t* __restrict__ map,
                  int64_t* __restrict__ row_kept) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * RL_WARPS + (threadIdx.x >> 5);
  if (i >= B) return;
  const int64_t v = idx[i];
  if (v < 0) {  // id outside the graph (flagged by mark_kernel): no edges
    if (lane == 0) row_kept[i] = 0;
    return;
  }
  const int64_t s = rowptr[v], e = rowptr[v + 1];
  int cnt = 0;
  for (int64_t j = s + lane; j < e; j += 32)
    if (map[(int64_t)col[j]] >= 0) ++cnt;
  cnt = warp_sum(cnt);
  if (lane == 0) row_kept[i] = cnt;
}

// within-batch: ordered compaction of the kept edges of each row
template <typename CT, typename OT>
__global__ void __launch_bounds__(RL_THREADS)
compact_kept_kernel(const int64_t* __restrict__ rowptr, const CT* __restrict__ col,
                    const float* __restrict__ val, const int64_t* __restrict__ idx, int64_t B,
                    const int32_t* __restrict__ map, const int64_t* __restrict__ row_start,
                    OT* __restrict__ out_col, float* __restrict__ out_val) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * RL_WARPS + (threadIdx.x >> 5);
  if (i >= B) return;
  const int64_t v = idx[i];
  if (v < 0) return;  // id outside the graph (flagged by mark_kernel)
  const int64_t s = rowptr[v], e = rowptr[v + 1];
  int64_t p = row_start[i];
  for (int64_t jb = s; jb < e; jb += 32) {
    const int64_t j = jb + lane;
    int32_t m = -1;
    if (j < e) m = map[(int64_t)col[j]];
    const unsigned bal = __ballot_sync(0xffffffffu, m >= 0);
    if (m >= 0) {
      const int64_t q = p + __popc(bal & ((1u << lane) - 1u));
      out_col[q] = (OT)m;
      if (val) out_val[q] = val[j];
    }
    p += __popc(bal);
  }
}

__global__ void store_counts_kernel(const int64_t* a, const int64_t* b, int64_t b_const,
                                    int64_t* counts_out) {
  pdl_prologue();
  counts_out[0] = a ? *a : 0;
  counts_out[1] = b ? *b : b_const;
}

static inline unsigned blocks_for(int64_t n, int per_block) {
  return (unsigned)((n + per_block - 1) / per_block);
}

static int check_relabel_args(const int64_t* rowptr, const void* col, int col_width,
                              const int64_t* idx, int64_t B, int64_t num_nodes, int64_t nnz_b,
                              int out_width, void* workspace) {
  IA_CHECK_ARG(B >= 0 && num_nodes >= 0 && nnz_b >= 0, "negative size");
  IA_CHECK_ARG(col_width == 4 || col_width == 8, "col_width must be 4 or 8");
  IA_CHECK_ARG(out_width == 4 || out_width == 8, "out_width must be 4 or 8");
  IA_CHECK_ARG(workspace != nullptr, "workspace is NULL");
  IA_CHECK_ARG(B <= num_nodes, "batch larger than the graph (B=%lld, N=%lld)", (long long)B,
               (long long)num_nodes);
  IA_CHECK_ARG(nnz_b < (int64_t)INT_MAX, "batch has too many edges for 32-bit edge positions");
  IA_CHECK_ARG(num_nodes < (int64_t)INT_MAX, "graph has too many nodes for 32-bit local ids");
  if (B > 0) IA_CHECK_ARG(rowptr && idx, "NULL argument");
  if (nnz_b > 0) IA_CHECK_ARG(col != nullptr, "col is NULL");
  return INCAGG_OK;
}

template <typename CT, typename OT>
static int relabel_one_hop_impl(const int64_t* rowptr, const CT* col, const float* val,
                                const int64_t* idx, int64_t B, int64_t N, int64_t nnz_b,
                                OT* out_rowptr, OT* out_col, float* out_val, int64_t* n_id_out,
                                int64_t* counts_out, void* workspace, cudaStream_t st) {
  RelabelWs w = carve_ws(workspace, N);
  if (B == 0) {
    IA_CUDA(cudaMemsetAsync(out_rowptr, 0, sizeof(OT), st));
    IA_CUDA(cudaMemsetAsync(counts_out, 0, 2 * sizeof(int64_t), st));
    return INCAGG_OK;
  }
  launch(mark_kernel, dim3(blocks_for(B, 256)), dim3(256), (size_t)(0), st, rowptr, idx, B, w.map, w.rowtmp, N, device_error_word(), w.ids);
  IA_LAUNCH_CHECK();
  int64_t* row_start = w.rowtmp;  // becomes the exclusive prefix of the degrees
  int rc = exclusive_scan_i64(w.rowtmp, row_start, B, w.scan, w.total, st);
  if (rc != INCAGG_OK) return rc;
  launch(write_rowptr_kernel<OT>, dim3(blocks_for(B + 1, 256)), dim3(256), (size_t)(0), st, row_start, w.total, B, out_rowptr);
  IA_LAUNCH_CHECK();
  const unsigned rb = blocks_for(B, RL_WARPS);
  int64_t* row_first = w.rowtmp2;
  launch(map_edges_kernel<CT, OT>, dim3(rb), dim3(RL_THREADS), (size_t)(0), st, rowptr, col, val, w.ids, B, row_start, w.map,
                                                      w.first_pos, out_col, out_val);
  IA_LAUNCH_CHECK();
  launch(count_first_kernel<CT, OT>, dim3(rb), dim3(RL_THREADS), (size_t)(0), st, rowptr, col, w.ids, B, row_start,
                                                        w.first_pos, out_col, row_first);
  IA_LAUNCH_CHECK();
  rc = exclusive_scan_i64(row_first, row_first, B, w.scan, w.total + 1, st);
  if (rc != INCAGG_OK) return rc;
  launch(assign_halo_kernel<CT, OT>, dim3(rb), dim3(RL_THREADS), (size_t)(0), st, rowptr, col, w.ids, B, row_start, row_first,
                                                        w.first_pos, out_col, w.map, n_id_out);
  IA_LAUNCH_CHECK();
  launch(fill_halo_cols_kernel<CT, OT>, dim3(rb), dim3(RL_THREADS), (size_t)(0), st, rowptr, col, w.ids, B, row_start, w.map,
                                                           out_col);
  IA_LAUNCH_CHECK();
  launch(store_counts_kernel, dim3(1), dim3(1), (size_t)(0), st, w.total + 1, w.total, 0, counts_out);
  IA_LAUNCH_CHECK();
  launch(finish_kernel, dim3(sm_count() * 4), dim3(256), (size_t)(0), st, idx, B, w.total + 1, n_id_out, w.map, w.first_pos, N);
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}

template <typename CT, typename OT>
static int relabel_within_impl(const int64_t* rowptr, const CT* col, const float* val,
                               const int64_t* idx, int64_t B, int64_t N, OT* out_rowptr,
                               OT* out_col, float* out_val, int64_t* counts_out, void* workspace,
                               cudaStream_t st) {
  RelabelWs w = carve_ws(workspace, N);
  if (B == 0) {
    IA_CUDA(cudaMemsetAsync(out_rowptr, 0, sizeof(OT), st));
    IA_CUDA(cudaMemsetAsync(counts_out, 0, 2 * sizeof(int64_t), st));
    return INCAGG_OK;
  }
  launch(mark_kernel, dim3(blocks_for(B, 256)), dim3(256), (size_t)(0), st, rowptr, idx, B, w.map, nullptr, N, device_error_word(), w.ids);
  IA_LAUNCH_CHECK();
  const unsigned rb = blocks_for(B, RL_WARPS);
  launch(count_kept_kernel<CT>, dim3(rb), dim3(RL_THREADS), (size_t)(0), st, rowptr, col, w.ids, B, w.map, w.rowtmp);
  IA_LAUNCH_CHECK();
  int rc = exclusive_scan_i64(w.rowtmp, w.rowtmp, B, w.scan, w.total, st);
  if (rc != INCAGG_OK) return rc;
  launch(write_rowptr_kernel<OT>, dim3(blocks_for(B + 1, 256)), dim3(256), (size_t)(0), st, w.rowtmp, w.total, B, out_rowptr);
  IA_LAUNCH_CHECK();
  launch(compact_kept_kernel<CT, OT>, dim3(rb), dim3(RL_THREADS), (size_t)(0), st, rowptr, col, val, w.ids, B, w.map, w.rowtmp,
                                                         out_col, out_val);
  IA_LAUNCH_CHECK();
  launch(store_counts_kernel, dim3(1), dim3(1), (size_t)(0), st, nullptr, w.total, 0, counts_out);
  IA_LAUNCH_CHECK();
  launch(finish_kernel, dim3(sm_count() * 4), dim3(256), (size_t)(0), st, idx, B, nullptr, nullptr, w.map, w.first_pos, N);
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}

}  // namespace incagg

using namespace incagg;

extern "C" size_t incagg_relabel_workspace_bytes(int64_t num_nodes) {
  if (num_nodes < 0) return 0;
  // map + first_pos + two per-row int64 arrays (degree prefix, first-edge counts) + scan scratch
  // + 2 totals; mirrors carve_ws().
  return sizeof(int32_t) * 2 * (size_t)num_nodes + 16 + sizeof(int64_t) * (2 * (size_t)num_nodes + 2) +
         sizeof(int64_t) * (ws_scan_slots(num_nodes) + 8) + sizeof(int64_t) * ((size_t)num_nodes + 8);
}

extern "C" int incagg_relabel_workspace_init(void* workspace, int64_t num_nodes,
                                             incagg_stream_t stream) {
  IA_CHECK_ARG(workspace != nullptr && num_nodes >= 0, "bad workspace arguments");
  if (num_nodes == 0) return INCAGG_OK;
  RelabelWs w = carve_ws(workspace, num_nodes);
  launch(ws_init_kernel, dim3(sm_count() * 8), dim3(256), (size_t)(0), as_stream(stream), w.map, w.first_pos, num_nodes);
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}

extern "C" int incagg_relabel_degree_sum(const int64_t* rowptr, const int64_t* idx, int64_t B,
                                         int64_t num_nodes, int64_t* degsum_out, void* workspace,
                                         incagg_stream_t stream) {
  (void)workspace;
  IA_CHECK_ARG(B >= 0 && degsum_out != nullptr, "bad arguments");
  cudaStream_t st = as_stream(stream);
  IA_CUDA(cudaMemsetAsync(degsum_out, 0, sizeof(int64_t), st));
  if (B == 0) return INCAGG_OK;
  IA_CHECK_ARG(rowptr && idx, "NULL argument");
  const unsigned blocks = blocks_for(B, 256) < (unsigned)(sm_count() * 8) ? blocks_for(B, 256)
                                                                            : (unsigned)(sm_count() * 8);
  launch(degree_sum_kernel, dim3(blocks), dim3(256), (size_t)(0), st, rowptr, idx, B,
                                            reinterpret_cast<unsigned long long*>(degsum_out), num_nodes);
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}

extern "C" int incagg_relabel_one_hop(const int64_t* rowptr, const void* col, int col_width,
                                      const float* val, const int64_t* idx, int64_t B,
                                      int64_t num_nodes, int64_t nnz_b, void* out_rowptr,
                                      void* out_col, int out_width, float* out_val,
                                      int64_t* n_id_out, int64_t* counts_out, void* workspace,
                                      incagg_stream_t stream) {
  int rc = check_relabel_args(rowptr, col, col_width, idx, B, num_nodes, nnz_b, out_width, workspace);
  if (rc != INCAGG_OK) return rc;
  IA_CHECK_ARG(out_rowptr && counts_out, "NULL output");
  IA_CHECK_ARG(B == 0 || n_id_out != nullptr, "n_id_out is NULL");
  IA_CHECK_ARG(nnz_b == 0 || out_col != nullptr, "out_col is NULL");
  IA_CHECK_ARG((val == nullptr) || nnz_b == 0 || out_val != nullptr, "out_val is NULL");
  cudaStream_t st = as_stream(stream);
#define IA_RL(CT, OT)                                                                          \
  return relabel_one_hop_impl<CT, OT>(rowptr, static_cast<const CT*>(col), val, idx, B, num_nodes, \
                                      nnz_b, static_cast<OT*>(out_rowptr), static_cast<OT*>(out_col), \
                                      out_val, n_id_out, counts_out, workspace, st)
  if (col_width == 4 && out_width == 4) IA_RL(int32_t, int32_t);
  if (col_width == 4 && out_width == 8) IA_RL(int32_t, int64_t);
  if (col_width == 8 && out_width == 4) IA_RL(int64_t, int32_t);
  IA_RL(int64_t, int64_t);
#undef IA_RL
}

extern "C" int incagg_relabel_one_hop_within_batch(const int64_t* rowptr, const void* col,
                                                   int col_width, const float* val,
                                                   const int64_t* idx, int64_t B, int64_t num_nodes,
                                                   int64_t nnz_b, void* out_rowptr, void* out_col,
                                                   int out_width, float* out_val,
                                                   int64_t* counts_out, void* workspace,
                                                   incagg_stream_t stream) {
  int rc = check_relabel_args(rowptr, col, col_width, idx, B, num_nodes, nnz_b, out_width, workspace);
  if (rc != INCAGG_OK) return rc;
  IA_CHECK_ARG(out_rowptr && counts_out, "NULL output");
  IA_CHECK_ARG(nnz_b == 0 || out_col != nullptr, "out_col is NULL");
  cudaStream_t st = as_stream(stream);
#define IA_RW(CT, OT)                                                                          \
  return relabel_within_impl<CT, OT>(rowptr, static_cast<const CT*>(col), val, idx, B, num_nodes, \
                                     static_cast<OT*>(out_rowptr), static_cast<OT*>(out_col),    \
                                     out_val, counts_out, workspace, st)
  if (col_width == 4 && out_width == 4) IA_RW(int32_t, int32_t);
  if (col_width == 4 && out_width == 8) IA_RW(int32_t, int64_t);
  if (col_width == 8 && out_width == 4) IA_RW(int64_t, int32_t);
  IA_RW(int64_t, int64_t);
#undef IA_RW
}
