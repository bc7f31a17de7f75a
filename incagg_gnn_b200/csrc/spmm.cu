// CSR SpMM family for the IncAgg-GNN propagation hot path (sm_100a).
//
//   incagg_spmm_csr     out = reduce_e val[e] * X[col[e]]            (sum/mean/min/max)
//   incagg_spmm_delta   out = reduce_e val[e] * (x[col[e]] - M_in[g(col[e])]) + M_ag[g(i)]
//   incagg_spmm_multi   K slabs of X reduced with K different reducers in one launch (PNA)
//   incagg_spmm_minmax_bwd
//
// These replace torch_sparse.matmul / spmm_{sum,mean,min,max} at the call sites listed in
// include/incagg_b200.h.  The work is HBM/L2-bound gather traffic (<= 0.5 flop/B), so the
// design is about bytes in flight, not tensor cores:
//   * one G-lane group per output row (G = 8/16/32 chosen so that G*VEC*NCH covers F with few
//     idle lanes; F=40 -> two rows per warp, F=128 -> one 512 B row segment per warp load);
//   * 128-bit (VEC=4) feature loads when rows are 16 B aligned, 64-bit / 32-bit fallbacks;
//   * col/val of G edges are fetched with one coalesced streaming load and broadcast with
//     group-masked shuffles; the edge loop is unrolled so each lane keeps >= 4 independent
//     128-bit loads in flight;
//   * int32 indices (half the index traffic of the reference's int64 path);
//   * rows longer than LONG_ROW edges are split across the warps of a CTA by a second
//     kernel (degree-bucketed scheduling) and combined in a fixed order -> deterministic.
// Wide sum / mean products (F > 64, 16-byte aligned rows: the GCNII / GraphSAGE / PNA layer widths) run
// through the merge-path kernel `spmm_stream_kernel` instead: the edge array is cut into equal
// pieces, one per resident warp, and every warp streams its piece with a software-pipelined ring of
// D 128-bit row gathers that never drains at row boundaries (see the kernel's comment).
#include <float.h>
#include <stdlib.h>
#include <limits.h>
#include <mutex>
#include <type_traits>

#include "common.cuh"

namespace incagg {

enum { R_SUM = INCAGG_REDUCE_SUM, R_MEAN = INCAGG_REDUCE_MEAN, R_MIN = INCAGG_REDUCE_MIN,
       R_MAX = INCAGG_REDUCE_MAX, R_RUNTIME = 4 };

template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<1> { using type = float; };

template <int VEC>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = __ldg(p);
  }
}
template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  } else {
    *p = v[0];
  }
}
template <int VEC>
__device__ __forceinline__ void store_vec_i(int32_t* p, const int32_t (&v)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<int4*>(p) = make_int4(v[0], v[1], v[2], v[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<int2*>(p) = make_int2(v[0], v[1]);
  } else {
    *p = v[0];
  }
}

// Degree-bucket plan of one CSR structure.  Rows with more than `long_row` edges are not handled by
// a lane group but by whole CTAs: a row of up to `chunk` edges is one work item, a longer ("giant")
// row is split into ceil(deg / chunk) items whose partial results meet in a scratch area.  The plan
// depends only on rowptr, so it is built once per structure and reused by every SpMM over it
// (all layers of a step, forward and - for the transposed structure - backward).
struct SpmmPlan {
  int32_t n_items;    // number of work items
  int32_t n_scratch;  // scratch slots handed out to split rows (may exceed the usable capacity)
  int32_t long_row;
  int32_t chunk;
  int32_t capacity;   // size of the item array
  // merge-path partition of the edge array (spmm_stream_kernel): warp slot w owns the edges
  // [e_base + w * q, e_base + (w + 1) * q) and starts in row wrow[w]
  int32_t n_wslots;
  int32_t q;          // edges per warp slot (a multiple of 32)
  int32_t e_base;     // rowptr[0]
  int32_t e_end;      // rowptr[rows]
  int32_t pad[3];
  // followed by int4 items[capacity]: {row, chunk index, scratch base or -1, number of chunks}
  // followed by int32 wrow[WS_MAX_SLOTS + 1]
};
static_assert(sizeof(SpmmPlan) == 48, "plan header is three int4");
constexpr int WS_WARPS = 8;           // warps per CTA of the stream kernel
constexpr int WS_TILE_F = 128;        // features per tile: 32 lanes x float4
constexpr int WS_MAX_SLOTS = 8192;    // upper bound of warp slots (148 SMs x 32 warps = 4736)
constexpr int WS_MAX_TILES = 8;       // feature tiles the partial-row scratch is sized for (F <= 1024)
constexpr int PART_STRIDE = 512;        // floats per scratch slot (= widest tile, 32 lanes * 4 * 4)
constexpr int PART_SLOTS = 16384;       // scratch slots per device (32 MB values + 32 MB args)
constexpr int PLAN_SCRATCH_CAP = 8192;  // slots one plan may hand out (times the feature tiles)
constexpr int ROW_GRAB = 4;             // short rows (x 32 / G) a warp fetches per trip to the row counter
constexpr int ROW_COUNTERS = 1024;      // feature tiles with a row counter of their own (gridDim.y)

struct SpmmParams {
  const int32_t* rowptr;
  const int32_t* col;
  const float* val;
  const float* X;
  int64_t ldx;
  float* out;
  int64_t ldo;
  int32_t* arg;
  int64_t lda;
  int64_t rows;
  int32_t F;
  // optional output gate: out[r, f] = 0 where gate[r, f] <= 0 (the ReLU mask of the backward pass of
  // a layer whose activation was fused into the producer of x)
  const float* gate;
  int64_t ld_gate;
  // delta extras
  const float* m_in;
  int64_t ld_in;
  const float* m_ag;
  int64_t ld_ag;
  const int64_t* n_id;
  // multi extras: slab k = blockIdx.y / tiles_per_slab uses reducers[k]
  int32_t slab_F;
  int32_t tiles_per_slab;
  int32_t reducers[16];
  // degree-bucket plan (see SpmmPlan) and the scratch area of split rows
  const SpmmPlan* plan;
  const int4* items;        // work items of the plan
  int32_t long_grid;        // CTAs [0, long_grid) walk the plan's items, the rest own short rows
  int32_t n_tiles;          // gridDim.y
  float* part_val;          // [part_slots][PART_STRIDE] partial results of split rows
  int32_t* part_arg;        // same shape, winning edge of min/max partials
  int32_t* part_done;       // [part_slots] arrival counters (left at zero by every call)
  int32_t part_slots;
  // stream kernel: partial sums of rows cut by a warp boundary, arrival counters, mean flag
  int32_t* row_counter;     // [ROW_COUNTERS + 1] next short row per feature tile, CTAs done (zero between calls)
  int32_t dynamic;          // short rows handed out through row_counter (large structures)
  float* ws_part;           // [tiles][n_wslots][2][WS_TILE_F]
  int32_t* ws_done;         // [tiles][n_wslots]
  int32_t mean;
};


template <int REDUCE>
__device__ __forceinline__ float red_init(int op) {
  const int r = (REDUCE == R_RUNTIME) ? op : REDUCE;
  return r == R_MIN ? FLT_MAX : (r == R_MAX ? -FLT_MAX : 0.f);
}

// Accumulate one edge's feature vector into acc (and arg) under reducer `op`.
template <int REDUCE, int VEC, bool ARG>
__device__ __forceinline__ void red_update(int op, float (&acc)[VEC], int32_t (&arg)[VEC], float v,
                                           const float (&x)[VEC], int e) {
  const int r = (REDUCE == R_RUNTIME) ? op : REDUCE;
  if (r == R_SUM || r == R_MEAN) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = fmaf(v, x[i], acc[i]);
  } else if (r == R_MIN) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float t = v * x[i];
      if (t < acc[i]) { acc[i] = t; if (ARG) arg[i] = e; }
    }
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float t = v * x[i];
      if (t > acc[i]) { acc[i] = t; if (ARG) arg[i] = e; }
    }
  }
}

// One G-lane group walks edges [s, e) of one row and accumulates NCH vectors per lane.
// lane_g: lane index inside the group; gmask: shuffle mask of the group.
// Row c of X starts at byte c * (ldx * 4): the product of two 32-bit values (the host checks
// ldx < 2^30), one IMAD.WIDE.U32 per gather instead of a 64-bit multiply-add chain.
template <int REDUCE, int VEC, int G, int NCH, bool DELTA, bool ARG>
__device__ __forceinline__ void walk_edges(const SpmmParams& p, int op, int s, int e, int fbase,
                                           int F, int lane_g, unsigned gmask,
                                           float (&acc)[NCH][VEC], int32_t (&arg)[NCH][VEC]) {
  constexpr int UNROLL = (NCH >= 4) ? 2 : 4;
  const unsigned x_bytes = (unsigned)p.ldx * 4u;
  const unsigned m_bytes = DELTA ? (unsigned)p.ld_in * 4u : 0u;
  const int f0 = fbase + lane_g * VEC;  // column of chunk 0; chunk k is G * VEC columns further
  const char* Xb = reinterpret_cast<const char*>(p.X + f0);
  const char* Mb = DELTA ? reinterpret_cast<const char*>(p.m_in + f0) : nullptr;
  bool fok[NCH];
#pragma unroll
  for (int k = 0; k < NCH; ++k) fok[k] = f0 + k * G * VEC < F;
  for (int base = s; base < e; base += G) {
    const int my_e = base + lane_g;
    int my_c = 0;
    float my_v = 0.f;
    if (my_e < e) {
      my_c = ldg_stream(p.col + my_e);
      my_v = p.val ? ldg_stream(p.val + my_e) : 1.f;
    }
    const int cnt = min(G, e - base);
    int j = 0;
    for (; j + UNROLL <= cnt; j += UNROLL) {
      int c[UNROLL];
      float v[UNROLL];
      float x[UNROLL][NCH][VEC];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        c[u] = __shfl_sync(gmask, my_c, j + u, G);
        v[u] = __shfl_sync(gmask, my_v, j + u, G);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const float* xr = reinterpret_cast<const float*>(Xb + (size_t)(unsigned)c[u] * x_bytes);
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
          if (fok[k]) {
            load_vec<VEC>(xr + k * G * VEC, x[u][k]);
          } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) x[u][k][i] = 0.f;
          }
        }
      }
      if constexpr (DELTA) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const unsigned g = p.n_id ? (unsigned)p.n_id[c[u]] : (unsigned)c[u];
          const float* mr = reinterpret_cast<const float*>(Mb + (size_t)g * m_bytes);
#pragma unroll
          for (int k = 0; k < NCH; ++k) {
            if (fok[k]) {
              float m[VEC];
              load_vec<VEC>(mr + k * G * VEC, m);
#pragma unroll
              for (int i = 0; i < VEC; ++i) x[u][k][i] -= m[i];
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int k = 0; k < NCH; ++k)
          red_update<REDUCE, VEC, ARG>(op, acc[k], arg[k], v[u], x[u][k], base + j + u);
    }
    for (; j < cnt; ++j) {
      const int c = __shfl_sync(gmask, my_c, j, G);
      const float v = __shfl_sync(gmask, my_v, j, G);
      const float* xr = reinterpret_cast<const float*>(Xb + (size_t)(unsigned)c * x_bytes);
      const float* mr = nullptr;
      if constexpr (DELTA) {
        const unsigned g = p.n_id ? (unsigned)p.n_id[c] : (unsigned)c;
        mr = reinterpret_cast<const float*>(Mb + (size_t)g * m_bytes);
      }
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        if (fok[k]) {
          float x[VEC];
          load_vec<VEC>(xr + k * G * VEC, x);
          if constexpr (DELTA) {
            float m[VEC];
            load_vec<VEC>(mr + k * G * VEC, m);
#pragma unroll
            for (int i = 0; i < VEC; ++i) x[i] -= m[i];
          }
          red_update<REDUCE, VEC, ARG>(op, acc[k], arg[k], v, x, base + j);
        }
      }
    }
  }
}

// Epilogue shared by the short-row and long-row kernels.
template <int REDUCE, int VEC, int G, int NCH, bool DELTA, bool ARG>
__device__ __forceinline__ void finish_row(const SpmmParams& p, int op, int64_t row, int deg,
                                           int fbase, int F, int lane_g, float (&acc)[NCH][VEC],
                                           int32_t (&arg)[NCH][VEC]) {
  const int r = (REDUCE == R_RUNTIME) ? op : REDUCE;
  const float inv = (r == R_MEAN) ? 1.f / (float)max(deg, 1) : 1.f;
  const float* ag = nullptr;
  if constexpr (DELTA) {
    const int64_t g = p.n_id ? p.n_id[row] : row;
    ag = p.m_ag + g * p.ld_ag;
  }
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int f = fbase + (k * G + lane_g) * VEC;
    if (f < F) {
      float o[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float a = acc[k][i];
        if (r == R_MEAN) a *= inv;
        if ((r == R_MIN || r == R_MAX) && deg == 0) a = 0.f;
        o[i] = a;
      }
      if constexpr (DELTA) {
        float m[VEC];
        load_vec<VEC>(ag + f, m);
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] += m[i];
      }
      if (p.gate != nullptr) {  // uniform branch
        float gt[VEC];
        load_vec<VEC>(p.gate + row * p.ld_gate + f, gt);
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = (gt[i] > 0.f) ? o[i] : 0.f;
      }
      store_vec<VEC>(p.out + row * p.ldo + f, o);
      if constexpr (ARG) {
        if (p.arg) store_vec_i<VEC>(p.arg + row * p.lda + f, arg[k]);
      }
    }
  }
}

#ifndef SPMM_THREADS_N
#define SPMM_THREADS_N 256
#endif
constexpr int SPMM_THREADS = SPMM_THREADS_N;
// 5 CTAs = 40 warps per SM and <= 51 registers per thread.  Measured per products batch (forward, cold L2):
// 6 CTAs (40 registers: ptxas interleaves the four gathers of a round with their FMAs) 42 us,
// 5 CTAs (48 registers: all four gathers issued first) 33 us, 4 CTAs (56 registers) 40 us.
#ifndef SPMM_MIN_CTAS
#define SPMM_MIN_CTAS (1280 / SPMM_THREADS_N)
#endif

// Feature tile -> (reducer, first feature, feature limit).  For the multi-aggregator launch
// blockIdx.y enumerates (slab, tile-in-slab) and a tile never crosses its slab.
template <int REDUCE, int COVER>
__device__ __forceinline__ void tile_info(const SpmmParams& p, int y, int& op, int& fbase, int& flim) {
  if constexpr (REDUCE == R_RUNTIME) {
    const int slab = y / p.tiles_per_slab;
    op = p.reducers[slab];
    fbase = slab * p.slab_F + (y % p.tiles_per_slab) * COVER;
    flim = min(p.F, (slab + 1) * p.slab_F);
  } else {
    op = REDUCE;
    fbase = y * COVER;
    flim = p.F;
  }
}

// One launch per SpMM.  blockIdx.y = feature tile.
//   CTAs [0, long_grid): walk the plan's work items (long / giant rows), one item per CTA trip:
//     the 8 warps take contiguous 32-aligned slices of the item's edges and are combined through
//     shared memory in warp order (deterministic; min/max keep the first winner because warps own
//     increasing edge ranges).  Items of a split row store their partial in the scratch area; the
//     last one to arrive combines them in chunk order and writes the row, so the result does not
//     depend on scheduling.  Low block indices are scheduled first, so the heavy rows start early.
//   remaining CTAs: one G-lane group per short row.
template <int REDUCE, int VEC, int G, int NCH, bool DELTA, bool ARG>
__global__ void __launch_bounds__(SPMM_THREADS, (NCH == 1 && !ARG) ? SPMM_MIN_CTAS : 1)
spmm_kernel(const SpmmParams p) {
  pdl_prologue();
  constexpr int COVER = G * VEC * NCH;
  const int lane = threadIdx.x & 31;
  const int long_row = p.plan->long_row;
  int op, fbase, flim;
  tile_info<REDUCE, COVER>(p, blockIdx.y, op, fbase, flim);

  if ((int)blockIdx.x >= p.long_grid) {
    // ---- short rows ----
    // Batch-sized structures (a few waves of CTAs): every warp owns 32 / G rows, fixed by its position
    // in the grid.  Large structures (a whole-graph sweep, tens of waves): the CTAs are one resident
    // wave and every warp fetches its next ROW_GRAB * (32 / G) rows from a counter until the rows are
    // used up, so no warp slot idles while a CTA waits for its longest row (rows differ in length by
    // two orders of magnitude).  Measured on the products shape: dynamic 2.36 ms vs 3.03 ms on the
    // 64 M-edge graph, but 47 us vs 35 us on a 0.43 M-edge batch (the trip to the counter adds a
    // dependent latency per grab that two or three grabs per warp do not amortise).  Which warp
    // computes a row does not affect the result.
    constexpr int RPW = 32 / G;  // rows a warp works on at a time
    const int lane_g = threadIdx.x % G;
    const int sub = lane / G;  // group index inside the warp
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (sub * G));
    int32_t* counter = p.row_counter + blockIdx.y;
    const int n_i = p.dynamic ? ROW_GRAB : 1;
    bool first = true;
    for (;;) {
      int base = 0;
      if (p.dynamic) {
        if (lane == 0) base = atomicAdd(counter, ROW_GRAB * RPW);
        base = __shfl_sync(0xffffffffu, base, 0);
      } else {  // static assignment: every warp owns RPW rows, the grid covers all rows
        if (!first) break;
        first = false;
        base = (int)(((int64_t)(blockIdx.x - p.long_grid) * (SPMM_THREADS / 32) + (threadIdx.x >> 5)) * RPW);
      }
      if ((int64_t)base >= p.rows) break;
#pragma unroll 1
      for (int i = 0; i < n_i; ++i) {
        const int64_t row = (int64_t)base + i * RPW + sub;
        if (row >= p.rows) break;  // (uniform per group; groups of a warp do not communicate)
        const int s = __ldg(p.rowptr + row), e = __ldg(p.rowptr + row + 1);
        if (e - s > long_row) continue;  // owned by the plan's items
        float acc[NCH][VEC];
        int32_t arg[NCH][VEC];
#pragma unroll
        for (int k = 0; k < NCH; ++k)
#pragma unroll
          for (int ii = 0; ii < VEC; ++ii) { acc[k][ii] = red_init<REDUCE>(op); arg[k][ii] = -1; }
        walk_edges<REDUCE, VEC, G, NCH, DELTA, ARG>(p, op, s, e, fbase, flim, lane_g, gmask, acc, arg);
        finish_row<REDUCE, VEC, G, NCH, DELTA, ARG>(p, op, row, e - s, fbase, flim, lane_g, acc, arg);
      }
    }
    if (!p.dynamic) return;
    // the last CTA of the launch to get here leaves the counters at zero for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
      const int n_short = (int)(gridDim.x - p.long_grid) * (int)gridDim.y;
      if (atomicAdd(p.row_counter + ROW_COUNTERS, 1) == n_short - 1) {
        for (int t = 0; t <= ROW_COUNTERS; ++t) p.row_counter[t] = 0;
      }
    }
    return;
  }

  // ---- long / giant rows ----
  constexpr int LN = (G == 32) ? NCH : 1;  // vectors per lane with 32 lanes per row segment
  static_assert(COVER <= 32 * VEC * LN, "long-row lanes must cover the tile");
  static_assert(32 * VEC * LN <= PART_STRIDE, "partial slot too small");
  constexpr int WARPS = SPMM_THREADS / 32;
  __shared__ float s_acc[WARPS][LN][32 * VEC];
  __shared__ int32_t s_arg[ARG ? WARPS : 1][LN][32 * VEC];
  const int n_items = min(p.plan->n_items, p.plan->capacity);
  const int chunk = p.plan->chunk;
  const int w = threadIdx.x >> 5;
  const int r = (REDUCE == R_RUNTIME) ? op : REDUCE;
  flim = min(flim, fbase + COVER);
  for (int li = blockIdx.x; li < n_items; li += p.long_grid) {
    const int4 ent = p.items[li];  // {row, chunk index, scratch base, number of chunks}
    const int64_t row = ent.x;
    const int rs = __ldg(p.rowptr + row), re = __ldg(p.rowptr + row + 1);
    const int deg = re - rs;
    // a split row whose scratch slots do not fit is walked whole by its first item
    bool split = ent.z >= 0;
    if (split && (int64_t)(ent.z + ent.w) * p.n_tiles > p.part_slots) {
      if (ent.y != 0) continue;  // (uniform over the CTA)
      split = false;
    }
    const int s = split ? rs + ent.y * chunk : rs;
    const int e = split ? min(re, s + chunk) : re;
    // split on multiples of 32 edges so every warp issues full coalesced index loads
    const int per = (((e - s) + WARPS - 1) / WARPS + 31) & ~31;
    const int ws = min(e, s + w * per), we = min(e, ws + per);
    float acc[LN][VEC];
    int32_t arg[LN][VEC];
#pragma unroll
    for (int k = 0; k < LN; ++k)
#pragma unroll
      for (int i = 0; i < VEC; ++i) { acc[k][i] = red_init<REDUCE>(op); arg[k][i] = -1; }
    walk_edges<REDUCE, VEC, 32, LN, DELTA, ARG>(p, op, ws, we, fbase, flim, lane, 0xffffffffu, acc, arg);
#pragma unroll
    for (int k = 0; k < LN; ++k)
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        s_acc[w][k][lane * VEC + i] = acc[k][i];
        if constexpr (ARG) s_arg[w][k][lane * VEC + i] = arg[k][i];
      }
    __syncthreads();
    if (w == 0) {
      for (int ww = 1; ww < WARPS; ++ww) {
#pragma unroll
        for (int k = 0; k < LN; ++k)
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float t = s_acc[ww][k][lane * VEC + i];
            if (r == R_SUM || r == R_MEAN) {
              acc[k][i] += t;
            } else if ((r == R_MIN && t < acc[k][i]) || (r == R_MAX && t > acc[k][i])) {
              acc[k][i] = t;
              if constexpr (ARG) arg[k][i] = s_arg[ww][k][lane * VEC + i];
            }
          }
      }
      bool write_row = true;
      if (split) {
        const int64_t slot0 = (int64_t)ent.z * p.n_tiles + (int64_t)blockIdx.y * ent.w;
        float* pv = p.part_val + (slot0 + ent.y) * PART_STRIDE;
        int32_t* pa = p.part_arg + (slot0 + ent.y) * PART_STRIDE;
#pragma unroll
        for (int k = 0; k < LN; ++k)
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            pv[(k * 32 + lane) * VEC + i] = acc[k][i];
            if constexpr (ARG) pa[(k * 32 + lane) * VEC + i] = arg[k][i];
          }
        __threadfence();
        int prev = 0;
        if (lane == 0) prev = atomicAdd(p.part_done + slot0, 1);
        prev = __shfl_sync(0xffffffffu, prev, 0);
        write_row = (prev == ent.w - 1);
        if (write_row) {  // last arriver: combine all partials in chunk order
          __threadfence();
#pragma unroll
          for (int k = 0; k < LN; ++k)
#pragma unroll
            for (int i = 0; i < VEC; ++i) { acc[k][i] = red_init<REDUCE>(op); arg[k][i] = -1; }
          for (int c = 0; c < ent.w; ++c) {
            const float* qv = p.part_val + (slot0 + c) * PART_STRIDE;
            const int32_t* qa = p.part_arg + (slot0 + c) * PART_STRIDE;
#pragma unroll
            for (int k = 0; k < LN; ++k)
#pragma unroll
              for (int i = 0; i < VEC; ++i) {
                const float t = __ldcg(qv + (k * 32 + lane) * VEC + i);
                if (r == R_SUM || r == R_MEAN) {
                  acc[k][i] += t;
                } else if ((r == R_MIN && t < acc[k][i]) || (r == R_MAX && t > acc[k][i])) {
                  acc[k][i] = t;
                  if constexpr (ARG) arg[k][i] = __ldcg(qa + (k * 32 + lane) * VEC + i);
                }
              }
          }
          if (lane == 0) p.part_done[slot0] = 0;  // leave the counters clean for the next call
        }
      }
      if (write_row)
        finish_row<REDUCE, VEC, 32, LN, DELTA, ARG>(p, op, row, deg, fbase, flim, lane, acc, arg);
    }
    __syncthreads();
  }
}

// ---- merge-path kernel for wide sum / mean products ---------------------------------------------
//
// The row-per-warp kernel above spends ~30 instructions per edge, leaves the memory system idle at
// every row end (a 26-edge row is six rounds of four gathers plus a serial remainder, and a new row
// starts with two dependent index loads), and rows of very different lengths make the tail of a
// batch-sized launch long.  Here the EDGE array is the unit of scheduling (merge-path SpMM): warp
// slot w owns the `q` consecutive edges [e_base + w q, e_base + (w + 1) q) whatever rows they belong
// to, the grid is one resident wave, so every warp does the same amount of work and finishes at the
// same time.  Inside its piece a warp
//   * stages (col, val) pairs in a 128-entry shared-memory ring, fetched in coalesced blocks of 32
//     edges three blocks ahead; a consumer reads its pair with one broadcast LDS.64;
//   * keeps two register buffers of D independent 128-bit gathers of X rows (one 512-byte row segment
//     per warp instruction): while one buffer is consumed the other is in flight, and a consumed
//     buffer is refilled at once - across row boundaries, so the pipeline never drains;
//   * walks the rows its edges belong to through a 64-entry shared-memory window of rowptr (refilled
//     32 rows ahead) and writes a row as soon as its last edge is consumed.
// Rows cut by a piece boundary: every piece that holds a part of the row stores its partial sum
// (head part = slot 0, tail part = slot 1 of the warp) when it has finished its piece, bumps the
// arrival counter of the row's first piece, and the last one to arrive adds the partials IN PIECE
// ORDER and writes the row - wait-free, and independent of scheduling (deterministic).  Empty rows
// are written (as zeros, plus M_ag for the delta form) by the piece that passes over them.
struct WsHeader {
  int n_w, q, e_base, e_end, capacity;
};

__device__ __forceinline__ void ws_finish(const SpmmParams& p, int row, int deg, float4 a, int f, bool fok) {
  if (!fok) return;
  if (p.mean) {
    const float inv = 1.f / (float)max(deg, 1);
    a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
  }
  if (p.m_ag != nullptr) {
    const float4 m = __ldg(reinterpret_cast<const float4*>(p.m_ag + (int64_t)row * p.ld_ag + f));
    a.x += m.x; a.y += m.y; a.z += m.z; a.w += m.w;
  }
  if (p.gate != nullptr) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(p.gate + (int64_t)row * p.ld_gate + f));
    a.x = g.x > 0.f ? a.x : 0.f; a.y = g.y > 0.f ? a.y : 0.f;
    a.z = g.z > 0.f ? a.z : 0.f; a.w = g.w > 0.f ? a.w : 0.f;
  }
  *reinterpret_cast<float4*>(p.out + (int64_t)row * p.ldo + f) = a;
}

// Part of a row that is cut by a piece boundary.  The partial sum is already in the warp's slot
// (slot 0 = head part of a row that started in an earlier piece, slot 1 = tail / middle part); here,
// at the end of the piece, it is published: the arrival counter of the row's first piece is bumped
// and the last piece to arrive adds the partials in piece order and writes the row.
__device__ __noinline__ void ws_publish(const SpmmParams& p, int w, int tile, int row, int f, bool fok) {
  const int lane = threadIdx.x & 31;
  const int4 h1 = __ldg(reinterpret_cast<const int4*>(p.plan) + 1);  // capacity, n_wslots, q, e_base
  const int n_w = h1.y, q = h1.z, e_base = h1.w;
  const int rs = __ldg(p.rowptr + row), re = __ldg(p.rowptr + row + 1);
  const int first = (rs - e_base) / q, last = (re - 1 - e_base) / q;
  const int64_t tbase = (int64_t)tile * n_w;
  __threadfence();
  int prev = 0;
  if (lane == 0) prev = atomicAdd(p.ws_done + tbase + first, 1);
  prev = __shfl_sync(0xffffffffu, prev, 0);
  if (prev != last - first) return;
  __threadfence();
  float4 s = __ldcg(reinterpret_cast<const float4*>(p.ws_part + ((tbase + first) * 2 + 1) * WS_TILE_F + lane * 4));
  for (int ww = first + 1; ww <= last; ++ww) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(p.ws_part + ((tbase + ww) * 2) * WS_TILE_F + lane * 4));
    s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
  }
  if (lane == 0) p.ws_done[tbase + first] = 0;  // clean for the next call
  ws_finish(p, row, re - rs, s, f, fok);
}

template <int D, int MINB, bool DELTA>
__global__ void __launch_bounds__(WS_WARPS * 32, MINB)
spmm_stream_kernel(const __grid_constant__ SpmmParams p) {
  pdl_prologue();
  __shared__ int s_c[WS_WARPS][128];    // col of 4 index blocks of 32 edges
  __shared__ float s_v[WS_WARPS][128];  // val
  __shared__ int s_rp[WS_WARPS][64];    // rowptr window
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int w = blockIdx.x * WS_WARPS + wi;
  const int f = blockIdx.y * WS_TILE_F + lane * 4;
  const bool fok = f < p.F;
  const int rows = (int)p.rows;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  int e0, e1;
  float* my_part;  // this warp's two partial-row slots
  {
    const int4 h1 = __ldg(reinterpret_cast<const int4*>(p.plan) + 1);  // capacity, n_wslots, q, e_base
    const int4 h2 = __ldg(reinterpret_cast<const int4*>(p.plan) + 2);  // e_end
    if (w >= h1.y) return;
    if (h1.w == h2.x) {  // a structure without edges: every row is empty
      for (int r = w; r < rows; r += h1.y) ws_finish(p, r, 0, zero4, f, fok);
      return;
    }
    e0 = h1.w + w * h1.z;
    if (e0 >= h2.x) return;
    e1 = min(e0 + h1.z, h2.x);
    my_part = p.ws_part + (((int64_t)blockIdx.y * h1.y + w) * 2) * WS_TILE_F + lane * 4;
  }
  int* sc = s_c[wi];
  float* sv = s_v[wi];
  int* rpw = s_rp[wi];

  // index blocks: block b = edges [e0 + 32 b, e0 + 32 b + 32), ring position (32 b + lane) & 127
  auto ld_blk = [&](int b, int& c, float& v) {
    const int idx = e0 + b * 32 + lane;
    c = 0;
    v = 0.f;
    if (idx < e1) {
      c = ldg_stream(p.col + idx);
      v = p.val ? ldg_stream(p.val + idx) : 1.f;
    }
  };
  int r, rbase, rp_pre, cn;
  float vn;
  {
    int c1, c2;
    float v1, v2;
    ld_blk(0, cn, vn);
    ld_blk(1, c1, v1);
    ld_blk(2, c2, v2);
    const int capacity = __ldg(reinterpret_cast<const int*>(p.plan) + 4);
    r = __ldg(reinterpret_cast<const int*>(p.items + capacity) + w);  // wrow[w]: the row that holds edge e0
    // rowptr window: the ring holds rowptr[rbase .. rbase + 63], rp_pre the next 32 entries
    rbase = r;
    const int rp_a = __ldg(p.rowptr + min(rbase + lane, rows));
    const int rp_b = __ldg(p.rowptr + min(rbase + 32 + lane, rows));
    rp_pre = __ldg(p.rowptr + min(rbase + 64 + lane, rows));
    sc[lane] = cn; sv[lane] = vn;
    sc[32 + lane] = c1; sv[32 + lane] = v1;
    sc[64 + lane] = c2; sv[64 + lane] = v2;
    ld_blk(3, cn, vn);  // stays in registers until block 0 has been consumed
    rpw[(rbase + lane) & 63] = rp_a;
    rpw[(rbase + 32 + lane) & 63] = rp_b;
  }
  __syncwarp();

  // row c of X / M_in starts at byte c * row_bytes (< 2^32, checked by the host): one IMAD.WIDE.U32
  // (lanes past F in the last tile gather column 0 instead - unconditional loads - and never store)
  const char* Xb = reinterpret_cast<const char*>(p.X + (fok ? f : 0));
  const char* Mb = DELTA ? reinterpret_cast<const char*>(p.m_in + (fok ? f : 0)) : nullptr;
  const unsigned x_bytes = (unsigned)p.ldx * 4u, m_bytes = DELTA ? (unsigned)p.ld_in * 4u : 0u;
  // Two register buffers of D gathers each: while buffer A is consumed, buffer B is in flight, and A
  // is refilled as soon as it has been consumed (batches, not a rolling ring: the hardware tracks
  // outstanding loads with a few counting scoreboards, so "wait for the oldest of D loads" would wait
  // for all of them).
  float4 xa[D], xb[D];
  float4 ma[DELTA ? D : 1], mb[DELTA ? D : 1];
  auto gather = [&](int pos, float4& xo, float4& mo) {  // gather of the edge at ring position `pos`
    const unsigned c = (unsigned)sc[pos & 127];
    xo = __ldg(reinterpret_cast<const float4*>(Xb + (size_t)c * x_bytes));
    if constexpr (DELTA) mo = __ldg(reinterpret_cast<const float4*>(Mb + (size_t)c * m_bytes));
  };
#pragma unroll
  for (int u = 0; u < D; ++u) {
    xa[u] = zero4; xb[u] = zero4;
    if constexpr (DELTA) { ma[u] = zero4; mb[u] = zero4; }
  }
  // prologue: the first D gathers
#pragma unroll
  for (int u = 0; u < D; ++u)
    if (e0 + u < e1) gather(u, xa[u], ma[DELTA ? u : 0]);
  if (w == 0)  // empty rows in front of the first edge
    for (int rr = 0; rr < r; ++rr) ws_finish(p, rr, 0, zero4, f, fok);
  int row_start = rpw[r & 63], row_end = rpw[(r + 1) & 63];
  const bool head_cut = row_start < e0;  // the first row started in an earlier piece
  float4 acc = zero4;
  auto next_row = [&]() {
    ++r;
    if (r - rbase >= 32) {  // slide the window: entries [rbase, rbase + 32) are no longer needed
      rpw[(rbase + 64 + lane) & 63] = rp_pre;
      rbase += 32;
      rp_pre = __ldg(p.rowptr + min(rbase + 64 + lane, rows));
      __syncwarp();
    }
    row_start = row_end;
    row_end = rpw[(r + 1) & 63];
  };
  auto row_done = [&]() {  // the last edge of row r has been consumed (the row ends inside this piece)
    if (row_start >= e0) ws_finish(p, r, row_end - row_start, acc, f, fok);
    else *reinterpret_cast<float4*>(my_part) = acc;  // head part, published at the end of the piece
    acc = zero4;
    next_row();
  };
  // LAST: the piece ends within the edges this iteration touches, so its end is tested per edge.
  auto issue_group = [&](int jg, float4 (&xo)[D], float4 (&mo)[DELTA ? D : 1], auto last_tag) {
    constexpr bool LAST = decltype(last_tag)::value;
#pragma unroll
    for (int u = 0; u < D; ++u)
      if (!LAST || jg + u < e1) gather(jg + u - e0, xo[u], mo[DELTA ? u : 0]);
  };
  auto consume_group = [&](int jg, const float4 (&xi)[D], const float4 (&mi)[DELTA ? D : 1], auto last_tag) {
    constexpr bool LAST = decltype(last_tag)::value;
#pragma unroll
    for (int u = 0; u < D; ++u) {
      const int jj = jg + u;
      if (!LAST || jj < e1) {  // uniform
        while (jj >= row_end) row_done();
        const float v = sv[(jj - e0) & 127];
        if constexpr (DELTA) {
          acc.x = fmaf(v, xi[u].x - mi[u].x, acc.x); acc.y = fmaf(v, xi[u].y - mi[u].y, acc.y);
          acc.z = fmaf(v, xi[u].z - mi[u].z, acc.z); acc.w = fmaf(v, xi[u].w - mi[u].w, acc.w);
        } else {
          acc.x = fmaf(v, xi[u].x, acc.x); acc.y = fmaf(v, xi[u].y, acc.y);
          acc.z = fmaf(v, xi[u].z, acc.z); acc.w = fmaf(v, xi[u].w, acc.w);
        }
      }
    }
  };
  auto do_pair = [&](int j, auto last_tag) {
    issue_group(j + D, xb, mb, last_tag);
    consume_group(j, xa, ma, last_tag);
    issue_group(j + 2 * D, xa, ma, last_tag);
    consume_group(j + D, xb, mb, last_tag);
  };
  for (int j = e0; j < e1; j += 2 * D) {
    const int off = j - e0;
    if ((off & 31) == 0 && off != 0) {
      // block off/32 - 1 is consumed: its ring slot takes the block held in registers (three ahead)
      const int b = (off >> 5) + 2;
      sc[(b * 32 + lane) & 127] = cn;
      sv[(b * 32 + lane) & 127] = vn;
      ld_blk(b + 1, cn, vn);
      __syncwarp();
    }
    if (j + 3 * D <= e1) do_pair(j, std::false_type{});
    else do_pair(j, std::true_type{});
  }
  // the row of the last edge: complete if it lies inside the piece, else a tail (or middle) part
  const bool tail_inside = row_end <= e1 && row_start >= e0;
  const int t_row = r;
  if (tail_inside) ws_finish(p, r, row_end - row_start, acc, f, fok);
  else *reinterpret_cast<float4*>(my_part + (row_start >= e0 ? WS_TILE_F : 0)) = acc;
  // empty rows that follow, up to the first row the next piece starts in
  if (row_end <= e1) {
    while (r + 1 < rows) {
      next_row();
      if (row_end > e1) break;
      ws_finish(p, r, 0, zero4, f, fok);
    }
  }
  // publish the parts of cut rows (the head part only if that row also ended inside this piece;
  // a row that covers the whole piece is the "tail" case above with slot 0)
  if (head_cut || !tail_inside) {
    const int capacity = __ldg(reinterpret_cast<const int*>(p.plan) + 4);
    const int r0 = __ldg(reinterpret_cast<const int*>(p.items + capacity) + w);
    if (head_cut && (r0 != t_row || tail_inside)) ws_publish(p, w, blockIdx.y, r0, f, fok);
    if (!tail_inside) ws_publish(p, w, blockIdx.y, t_row, f, fok);
  }
}

// Measured alternatives of this kernel that are not shipped (profiles/r02_spmm_variants.md; the code is in
// the history: two edges per warp instruction with 256-bit gathers per half-warp, commit 07d9d4e; source
// rows landed in a shared-memory ring by cp.async.bulk, df70b15, or by cp.async, c7a49a7): all of them are
// bound by the per-warp instruction stream, none beats the row kernel on plain products.

// Plan construction: one thread per row appends the row's work items.
__global__ void spmm_plan_kernel(const int32_t* __restrict__ rowptr, int64_t rows, SpmmPlan* plan,
                                 int4* __restrict__ items, int long_row, int chunk, int capacity,
                                 int n_wslots) {
  pdl_prologue();
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int e_base = rowptr[0], e_end = rowptr[rows];
  const int64_t nnz = (int64_t)e_end - e_base;
  if (chunk <= 0) {
    // adaptive chunk: small enough that one CTA trip is a few microseconds on batch-sized
    // structures, large enough that the split rows of a whole graph fit the scratch area
    chunk = 256;
    while (chunk < 4096 && (int64_t)chunk * (PLAN_SCRATCH_CAP / 2) < nnz) chunk <<= 1;
  }
  // merge-path partition: q edges per warp slot, a multiple of 32 (aligned index blocks)
  int64_t q64 = (nnz + n_wslots - 1) / n_wslots;
  q64 = (q64 + 31) / 32 * 32;
  if (q64 < 32) q64 = 32;
  const int q = (int)q64;
  if (row == 0) {
    plan->long_row = long_row;
    plan->chunk = chunk;
    plan->capacity = capacity;
    plan->n_wslots = n_wslots;
    plan->q = q;
    plan->e_base = e_base;
    plan->e_end = e_end;
  }
  if (row <= n_wslots) {
    // wrow[w] = the row that holds edge e_base + w q: the largest r with rowptr[r] <= that edge
    int* wrow = reinterpret_cast<int*>(items + capacity);
    const int64_t tgt = (int64_t)e_base + row * q64;
    int r = (int)rows;
    if (tgt < e_end) {
      int lo = 0, hi = (int)rows;  // invariant: rowptr[lo] <= tgt < rowptr[hi]
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (rowptr[mid] <= tgt) lo = mid; else hi = mid;
      }
      r = lo;
    }
    wrow[row] = r;
  }
  if (row >= rows) return;
  const int deg = rowptr[row + 1] - rowptr[row];
  if (deg <= long_row) return;
  int n = (deg + chunk - 1) / chunk, sbase = -1;
  if (n > 1) {
    sbase = atomicAdd(&plan->n_scratch, n);
    if (sbase + n > PLAN_SCRATCH_CAP) { sbase = -1; n = 1; }
  }
  const int slot = atomicAdd(&plan->n_items, n);
  if (slot + n <= capacity)
    for (int c = 0; c < n; ++c) items[slot + c] = make_int4((int)row, c, sbase, n);
}

__global__ void minmax_bwd_kernel(const int32_t* __restrict__ col, const float* __restrict__ val,
                                  const int32_t* __restrict__ arg, int64_t lda,
                                  const float* __restrict__ grad_out, int64_t ldg,
                                  float* __restrict__ grad_x, int64_t ldx, int64_t rows, int F) {
  pdl_prologue();
  const int64_t total = rows * F;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / F;
    const int f = (int)(t - i * F);
    const int e = arg[i * lda + f];
    if (e >= 0) {
      const float v = val ? val[e] : 1.f;
      atomicAdd(grad_x + (int64_t)col[e] * ldx + f, v * grad_out[i * ldg + f]);
    }
  }
}

// ---- host-side dispatch ------------------------------------------------------
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
// Degree buckets: rows up to LONG_ROW edges -> one lane group; up to CHUNK edges -> one CTA; longer
// rows -> ceil(deg / CHUNK) CTAs, CHUNK = 256 .. 4096 by structure size.  (Tunable for experiments
// through the environment.)
static int long_row_edges() { static int v = env_int("INCAGG_SPMM_LONG_ROW", 64); return v; }
static int chunk_edges() { static int v = env_int("INCAGG_SPMM_CHUNK", 0); return v; }  // 0 = adaptive

// Stream (merge-path) kernel configuration (gathers per buffer x CTAs of 8 warps per SM):
// variant 0 = 8 x 2, 1 = 4 x 4, 2 = 4 x 3, 3 = 2 x 5, 4 = 2 x 6, 5 = 16 x 1.  The number of warp slots is
// one resident wave.
// Which kernel serves wide sum / mean products.  -2 (default) = by measurement on the products shape
// (profiles/r02_spmm_variants.md): the row kernel for plain products (35 us vs 35-41 us per batch,
// 2.4-3.0 ms vs 4.7 ms on the whole graph), the merge-path kernel (variant 2) for the incremental-
// aggregation delta form (44 us vs 49 us); -1 = row kernel everywhere; >= 0 = that merge-path variant.
static int stream_variant() {
  static int dflt = env_int("INCAGG_SPMM_STREAM", -2);
  return incagg::tune_get(INCAGG_TUNE_SPMM_STREAM_VARIANT, dflt);
}
static int stream_variant_for(bool delta) {
  const int v = stream_variant();
  if (v == -2) return delta ? 2 : -1;
  return v;
}
static int stream_min_f() { return incagg::tune_get(INCAGG_TUNE_SPMM_STREAM_MIN_F, 65); }
static int stream_ctas_per_sm(int variant) {
  return variant == 1 ? 4 : (variant == 2 ? 3 : (variant == 3 ? 5 : (variant == 4 ? 6 : (variant == 5 ? 1 : 2))));
}
static int stream_wslots() {
  const int v = stream_variant();
  int n = sm_count() * stream_ctas_per_sm(v == -2 ? 2 : v) * WS_WARPS;
  return n > WS_MAX_SLOTS ? WS_MAX_SLOTS : n;
}

// Upper bound of the work items of a structure: one per row longer than LONG_ROW plus the extra
// chunks of split rows.  nnz < 0 = unknown.
static int64_t plan_capacity(int64_t rows, int64_t nnz) {
  int64_t long_rows = rows;
  if (nnz >= 0) {
    const int64_t by_edges = nnz / (long_row_edges() + 1) + 1;
    if (by_edges < long_rows) long_rows = by_edges;
  }
  return long_rows + PLAN_SCRATCH_CAP + 8;
}
static size_t plan_bytes_for(int64_t capacity) {
  return sizeof(SpmmPlan) + sizeof(int4) * (size_t)capacity + sizeof(int32_t) * (WS_MAX_SLOTS + 1);
}

static int build_plan(const int32_t* rowptr, int64_t rows, void* plan, int64_t capacity,
                      cudaStream_t st) {
  IA_CHECK_ARG(capacity < 0x7fffffff, "plan too large");
  IA_CUDA(cudaMemsetAsync(plan, 0, sizeof(SpmmPlan), st));
  if (rows == 0) return INCAGG_OK;
  SpmmPlan* hdr = static_cast<SpmmPlan*>(plan);
  int4* items = reinterpret_cast<int4*>(hdr + 1);
  const int n_w = stream_wslots();
  const int64_t threads = rows > n_w + 1 ? rows : n_w + 1;
  launch(spmm_plan_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), (size_t)(0), st, rowptr, rows, hdr,
         items, long_row_edges(), chunk_edges(), (int)capacity, n_w);
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}

struct SpmmScratch {
  float* part_val = nullptr;
  int32_t* part_arg = nullptr;
  int32_t* part_done = nullptr;
  int32_t* row_counter = nullptr;
  float* ws_part = nullptr;     // stream kernel: [WS_MAX_TILES][WS_MAX_SLOTS][2][WS_TILE_F]
  int32_t* ws_done = nullptr;   // [WS_MAX_TILES][WS_MAX_SLOTS]
  void* plan = nullptr;  // temporary plan of calls that pass none
  int64_t plan_capacity = 0;
};
// Per-device scratch of split rows, shared by all host threads (allocated once / the temporary plan
// grown on demand).  It is reused by every call, so SpMM launches that use it must be ordered on the
// device: the Python layer serialises SpMM calls issued on different streams (ops._order_spmm).
static int get_scratch(SpmmScratch** out, int64_t want_plan_capacity) {
  static SpmmScratch scratch[16];
  static std::mutex mu;
  int dev = 0;
  IA_CUDA(cudaGetDevice(&dev));
  IA_CHECK_ARG(dev >= 0 && dev < 16, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lock(mu);
  SpmmScratch& s = scratch[dev];
  if (s.part_val == nullptr) {
    IA_CUDA(cudaMalloc(&s.part_val, sizeof(float) * (size_t)PART_SLOTS * PART_STRIDE));
    IA_CUDA(cudaMalloc(&s.part_arg, sizeof(int32_t) * (size_t)PART_SLOTS * PART_STRIDE));
    IA_CUDA(cudaMalloc(&s.part_done, sizeof(int32_t) * (size_t)PART_SLOTS));
    IA_CUDA(cudaMemset(s.part_done, 0, sizeof(int32_t) * (size_t)PART_SLOTS));
    IA_CUDA(cudaMalloc(&s.row_counter, sizeof(int32_t) * (ROW_COUNTERS + 1)));
    IA_CUDA(cudaMemset(s.row_counter, 0, sizeof(int32_t) * (ROW_COUNTERS + 1)));
    IA_CUDA(cudaMalloc(&s.ws_part, sizeof(float) * (size_t)WS_MAX_TILES * WS_MAX_SLOTS * 2 * WS_TILE_F));
    IA_CUDA(cudaMalloc(&s.ws_done, sizeof(int32_t) * (size_t)WS_MAX_TILES * WS_MAX_SLOTS));
    IA_CUDA(cudaMemset(s.ws_done, 0, sizeof(int32_t) * (size_t)WS_MAX_TILES * WS_MAX_SLOTS));
  }
  if (want_plan_capacity > s.plan_capacity) {
    int64_t cap = s.plan_capacity ? s.plan_capacity : (1 << 16);
    while (cap < want_plan_capacity) cap <<= 1;
    if (s.plan) IA_CUDA(cudaFree(s.plan));  // synchronises the device: earlier users are done
    s.plan = nullptr;
    s.plan_capacity = 0;
    IA_CUDA(cudaMalloc(&s.plan, plan_bytes_for(cap)));
    s.plan_capacity = cap;
  }
  *out = &s;
  return INCAGG_OK;
}

// Wide sum / mean products through the merge-path kernel.  Returns 1 when the call does not qualify
// (the caller falls through to the row kernel), an error code, or INCAGG_OK after the launch.
static int try_stream(SpmmParams& p, int reduce, int vec, cudaStream_t st) {
  const int variant = stream_variant_for(p.m_in != nullptr);
  if (variant < 0) return 1;
  if (vec != 4 || p.F < stream_min_f() || p.F > WS_TILE_F * WS_MAX_TILES) return 1;
  if ((reduce != R_SUM && reduce != R_MEAN) || p.arg != nullptr || p.n_id != nullptr) return 1;
  if (p.rows >= 0x7fffffff || p.ldx >= (1ll << 30) || p.ld_in >= (1ll << 30)) return 1;
  SpmmScratch* sc = nullptr;
  const bool own_plan = (p.plan == nullptr);
  int rc = get_scratch(&sc, own_plan ? plan_capacity(p.rows, -1) : 0);
  if (rc != INCAGG_OK) return rc;
  if (own_plan) {
    rc = build_plan(p.rowptr, p.rows, sc->plan, sc->plan_capacity, st);
    if (rc != INCAGG_OK) return rc;
    p.plan = static_cast<const SpmmPlan*>(sc->plan);
  }
  p.items = reinterpret_cast<const int4*>(p.plan + 1);
  p.ws_part = sc->ws_part;
  p.ws_done = sc->ws_done;
  p.mean = (reduce == R_MEAN);
  const int n_w = stream_wslots();
  const int tiles = (p.F + WS_TILE_F - 1) / WS_TILE_F;
  dim3 grid((unsigned)((n_w + WS_WARPS - 1) / WS_WARPS), (unsigned)tiles);
  const bool delta = p.m_in != nullptr;
#define IA_STREAM(D_, MINB_)                                                                         \
  do {                                                                                               \
    if (delta) launch(spmm_stream_kernel<((D_) > 1 ? (D_) / 2 : 1), MINB_, true>, grid, dim3(WS_WARPS * 32), (size_t)(0), st, p); \
    else launch(spmm_stream_kernel<D_, MINB_, false>, grid, dim3(WS_WARPS * 32), (size_t)(0), st, p);   \
  } while (0)
  if (variant == 1) IA_STREAM(4, 4);
  else if (variant == 2) IA_STREAM(4, 3);
  else if (variant == 3) IA_STREAM(2, 5);
  else if (variant == 4) IA_STREAM(2, 6);
  else if (variant == 5) IA_STREAM(16, 1);
  else IA_STREAM(8, 2);
#undef IA_STREAM
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}

template <int REDUCE, int VEC, int G, int NCH, bool DELTA, bool ARG>
static int launch_cfg(SpmmParams& p, int n_tiles, int64_t items_bound, cudaStream_t st) {
  constexpr int GROUPS = SPMM_THREADS / G;
  const int64_t blocks = (p.rows + GROUPS - 1) / GROUPS;
  if (blocks == 0) return INCAGG_OK;
  SpmmScratch* sc = nullptr;
  const bool own_plan = (p.plan == nullptr);
  int rc = get_scratch(&sc, own_plan ? plan_capacity(p.rows, -1) : 0);
  if (rc != INCAGG_OK) return rc;
  if (own_plan) {  // no plan given: build a temporary one (one memset + one small launch)
    rc = build_plan(p.rowptr, p.rows, sc->plan, sc->plan_capacity, st);
    if (rc != INCAGG_OK) return rc;
    p.plan = static_cast<const SpmmPlan*>(sc->plan);
    items_bound = p.rows + PLAN_SCRATCH_CAP;
  }
  p.items = reinterpret_cast<const int4*>(p.plan + 1);
  p.part_val = sc->part_val;
  p.part_arg = sc->part_arg;
  p.part_done = sc->part_done;
  p.part_slots = PART_SLOTS;
  p.n_tiles = n_tiles;
  p.row_counter = sc->row_counter;
  IA_CHECK_ARG(n_tiles <= ROW_COUNTERS, "too many feature tiles (%d)", n_tiles);
  IA_CHECK_ARG(p.ldx < (1ll << 30) && p.ld_in < (1ll << 30), "row stride too large");
  // CTAs that walk the plan's items: never more than the items can be, at most 8 per SM
  int64_t lg = (int64_t)env_int("INCAGG_SPMM_LONG_CTAS", 8) * sm_count();
  if (items_bound < lg) lg = items_bound;
  if (lg < 1) lg = 1;
  p.long_grid = (int)lg;
  // CTAs of the short rows: one resident wave, rows are handed out dynamically
  static int ctas_per_sm = 0;  // (per template instance)
  if (ctas_per_sm == 0) {
    int n = 0;
    IA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, spmm_kernel<REDUCE, VEC, G, NCH, DELTA, ARG>,
                                                          SPMM_THREADS, 0));
    ctas_per_sm = n > 0 ? n : 1;
  }
  // dynamic row hand-out pays off from ~8 waves of CTAs on (INCAGG_SPMM_DYNAMIC=0/1 forces it)
  static int force_dyn = env_int("INCAGG_SPMM_DYNAMIC", -1);
  int64_t n_short = (int64_t)ctas_per_sm * sm_count();
  p.dynamic = force_dyn >= 0 ? force_dyn : (blocks > 8 * n_short);
  if (!p.dynamic || blocks < n_short) n_short = blocks;
  dim3 grid((unsigned)(n_short + lg), (unsigned)n_tiles);
  launch(spmm_kernel<REDUCE, VEC, G, NCH, DELTA, ARG>, dim3(grid), dim3(SPMM_THREADS), (size_t)(0), st, p);
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}

// Pick (VEC, G, NCH) for a feature width and alignment.
template <int REDUCE, bool DELTA, bool ARG>
static int dispatch_shape(SpmmParams& p, int width, int vec, int tiles_mult, int64_t items_bound,
                          cudaStream_t st) {
  // `width` = number of features one tile row must cover (F, or slab_F for multi)
  const int nvec = (width + vec - 1) / vec;
  auto tiles = [&](int cover) { return tiles_mult * ((width + cover - 1) / cover); };
#define IA_CFG(V, G_, N_)                                                    \
  do {                                                                       \
    if (REDUCE == R_RUNTIME) p.tiles_per_slab = (width + V * G_ * N_ - 1) / (V * G_ * N_); \
    return launch_cfg<REDUCE, V, G_, N_, DELTA, ARG>(p, tiles(V * G_ * N_), items_bound, st); \
  } while (0)
  if (vec == 4) {
    if (nvec <= 8) IA_CFG(4, 8, 1);
    if (nvec <= 16) IA_CFG(4, 16, 1);
    if (nvec <= 32) IA_CFG(4, 32, 1);
    if (nvec <= 64) IA_CFG(4, 32, 2);
    IA_CFG(4, 32, 4);
  } else if (vec == 2) {
    if (nvec <= 32) IA_CFG(2, 32, 1);
    IA_CFG(2, 32, 4);
  } else {
    if (nvec <= 8) IA_CFG(1, 8, 1);
    if (nvec <= 32) IA_CFG(1, 32, 1);
    IA_CFG(1, 32, 4);
  }
#undef IA_CFG
}

static int pick_vec(const SpmmParams& p, int width) {
  bool a16 = aligned16(p.X) && aligned16(p.out) && p.ldx % 4 == 0 && p.ldo % 4 == 0 && width % 4 == 0;
  bool a8 = aligned8(p.X) && aligned8(p.out) && p.ldx % 2 == 0 && p.ldo % 2 == 0 && width % 2 == 0;
  if (p.arg) {
    a16 = a16 && aligned16(p.arg) && p.lda % 4 == 0;
    a8 = a8 && aligned8(p.arg) && p.lda % 2 == 0;
  }
  if (p.gate) {
    a16 = a16 && aligned16(p.gate) && p.ld_gate % 4 == 0;
    a8 = a8 && aligned8(p.gate) && p.ld_gate % 2 == 0;
  }
  if (p.m_in) {
    a16 = a16 && aligned16(p.m_in) && aligned16(p.m_ag) && p.ld_in % 4 == 0 && p.ld_ag % 4 == 0;
    a8 = a8 && aligned8(p.m_in) && aligned8(p.m_ag) && p.ld_in % 2 == 0 && p.ld_ag % 2 == 0;
  }
  return a16 ? 4 : (a8 ? 2 : 1);
}

static int check_common(const int32_t* rowptr, const int32_t* col, const float* X, float* out,
                        int64_t ldx, int64_t ldo, int64_t rows, int32_t F) {
  IA_CHECK_ARG(rows >= 0 && F >= 0, "negative size (rows=%lld, F=%d)", (long long)rows, F);
  if (rows == 0 || F == 0) return INCAGG_OK;
  IA_CHECK_ARG(rowptr != nullptr, "rowptr is NULL");
  IA_CHECK_ARG(out != nullptr, "out is NULL");
  IA_CHECK_ARG(ldx >= F && ldo >= F, "leading dimension smaller than F");
  (void)col; (void)X;
  return INCAGG_OK;
}

}  // namespace incagg

using namespace incagg;

extern "C" size_t incagg_spmm_plan_bytes(int64_t rows, int64_t nnz) {
  if (rows < 0) return 0;
  return plan_bytes_for(plan_capacity(rows, nnz));
}

extern "C" int incagg_spmm_plan(const int32_t* rowptr, int64_t rows, int64_t nnz, void* plan,
                                size_t plan_bytes, incagg_stream_t stream) {
  IA_CHECK_ARG(rows >= 0, "negative size");
  IA_CHECK_ARG(plan != nullptr && plan_bytes >= incagg_spmm_plan_bytes(rows, nnz), "plan buffer too small");
  IA_CHECK_ARG(rows == 0 || rowptr != nullptr, "rowptr is NULL");
  IA_CHECK_ARG((reinterpret_cast<uintptr_t>(plan) & 15) == 0, "plan buffer must be 16-byte aligned");
  // (the item capacity must be the one plan_bytes was computed for: wrow[] lies behind the items)
  return build_plan(rowptr, rows, plan, plan_capacity(rows, nnz), as_stream(stream));
}

static int spmm_csr_impl(int reduce, const int32_t* rowptr, const int32_t* col, const float* val,
                         const float* X, int64_t ldx, float* out, int64_t ldo, int32_t* arg_out, int64_t lda,
                         int64_t rows, int32_t F, const void* plan, const float* gate, int64_t ld_gate,
                         incagg_stream_t stream) {
  int rc = check_common(rowptr, col, X, out, ldx, ldo, rows, F);
  if (rc != INCAGG_OK) return rc;
  if (rows == 0 || F == 0) return INCAGG_OK;
  IA_CHECK_ARG(reduce >= 0 && reduce <= 3, "unknown reducer %d", reduce);
  IA_CHECK_ARG(arg_out == nullptr || lda >= F, "lda smaller than F");
  IA_CHECK_ARG(gate == nullptr || ld_gate >= F, "ld_gate smaller than F");
  SpmmParams p{};
  p.rowptr = rowptr; p.col = col; p.val = val; p.X = X; p.ldx = ldx; p.out = out; p.ldo = ldo;
  p.arg = (reduce == R_MIN || reduce == R_MAX) ? arg_out : nullptr;
  p.lda = lda; p.rows = rows; p.F = F;
  p.gate = gate; p.ld_gate = ld_gate;
  p.plan = static_cast<const SpmmPlan*>(plan);
  const int vec = pick_vec(p, F);
  cudaStream_t st = as_stream(stream);
  const int64_t ib = INT64_MAX;
  rc = try_stream(p, reduce, vec, st);
  if (rc != 1) return rc;
  switch (reduce) {
    case R_SUM: return dispatch_shape<R_SUM, false, false>(p, F, vec, 1, ib, st);
    case R_MEAN: return dispatch_shape<R_MEAN, false, false>(p, F, vec, 1, ib, st);
    case R_MIN:
      return p.arg ? dispatch_shape<R_MIN, false, true>(p, F, vec, 1, ib, st)
                   : dispatch_shape<R_MIN, false, false>(p, F, vec, 1, ib, st);
    default:
      return p.arg ? dispatch_shape<R_MAX, false, true>(p, F, vec, 1, ib, st)
                   : dispatch_shape<R_MAX, false, false>(p, F, vec, 1, ib, st);
  }
}

extern "C" int incagg_spmm_csr(int reduce, const int32_t* rowptr, const int32_t* col,
                               const float* val, const float* X, int64_t ldx, float* out,
                               int64_t ldo, int32_t* arg_out, int64_t lda, int64_t rows, int32_t F,
                               const void* plan, incagg_stream_t stream) {
  return spmm_csr_impl(reduce, rowptr, col, val, X, ldx, out, ldo, arg_out, lda, rows, F, plan, nullptr, 0,
                       stream);
}

extern "C" int incagg_spmm_csr_gated(int reduce, const int32_t* rowptr, const int32_t* col,
                                     const float* val, const float* X, int64_t ldx, float* out,
                                     int64_t ldo, int64_t rows, int32_t F, const void* plan,
                                     const float* gate, int64_t ld_gate, incagg_stream_t stream) {
  IA_CHECK_ARG(reduce == R_SUM || reduce == R_MEAN, "gated SpMM supports sum / mean (got %d)", reduce);
  IA_CHECK_ARG(gate != nullptr, "gate is NULL");
  return spmm_csr_impl(reduce, rowptr, col, val, X, ldx, out, ldo, nullptr, 0, rows, F, plan, gate, ld_gate,
                       stream);
}

extern "C" int incagg_spmm_delta(int reduce, const int32_t* rowptr, const int32_t* col,
                                 const float* val, const float* x, int64_t ldx, const float* m_in,
                                 int64_t ld_in, const float* m_ag, int64_t ld_ag,
                                 const int64_t* n_id, float* out, int64_t ldo, int64_t rows,
                                 int32_t F, const void* plan, incagg_stream_t stream) {
  int rc = check_common(rowptr, col, x, out, ldx, ldo, rows, F);
  if (rc != INCAGG_OK) return rc;
  if (rows == 0 || F == 0) return INCAGG_OK;
  IA_CHECK_ARG(reduce == R_SUM || reduce == R_MEAN, "delta supports sum/mean only (got %d)", reduce);
  IA_CHECK_ARG(m_in != nullptr && m_ag != nullptr, "M_in / M_ag is NULL");
  IA_CHECK_ARG(ld_in >= F && ld_ag >= F, "history leading dimension smaller than F");
  SpmmParams p{};
  p.rowptr = rowptr; p.col = col; p.val = val; p.X = x; p.ldx = ldx; p.out = out; p.ldo = ldo;
  p.rows = rows; p.F = F; p.m_in = m_in; p.ld_in = ld_in; p.m_ag = m_ag; p.ld_ag = ld_ag;
  p.n_id = n_id;
  p.plan = static_cast<const SpmmPlan*>(plan);
  const int vec = pick_vec(p, F);
  cudaStream_t st = as_stream(stream);
  rc = try_stream(p, reduce, vec, st);
  if (rc != 1) return rc;
  if (reduce == R_SUM) return dispatch_shape<R_SUM, true, false>(p, F, vec, 1, INT64_MAX, st);
  return dispatch_shape<R_MEAN, true, false>(p, F, vec, 1, INT64_MAX, st);
}

static int spmm_multi_impl(const int32_t* rowptr, const int32_t* col, const float* val, const float* X,
                           int64_t ldx, float* out, int64_t ldo, int32_t* arg_out, int64_t lda, int64_t rows,
                           int32_t F, int32_t K, const int32_t* reducers, const void* plan,
                           incagg_stream_t stream) {
  IA_CHECK_ARG(K >= 1 && K <= 16, "K must be in [1, 16] (got %d)", K);
  IA_CHECK_ARG(reducers != nullptr, "reducers is NULL");
  int rc = check_common(rowptr, col, X, out, ldx, ldo, rows, F * K);
  if (rc != INCAGG_OK) return rc;
  if (rows == 0 || F == 0) return INCAGG_OK;
  IA_CHECK_ARG(arg_out == nullptr || lda >= (int64_t)F * K, "lda smaller than K * F");
  SpmmParams p{};
  p.rowptr = rowptr; p.col = col; p.val = val; p.X = X; p.ldx = ldx; p.out = out; p.ldo = ldo;
  p.rows = rows; p.F = F * K; p.slab_F = F;
  p.arg = arg_out; p.lda = lda;
  for (int k = 0; k < K; ++k) {
    IA_CHECK_ARG(reducers[k] >= 0 && reducers[k] <= 3, "unknown reducer %d", reducers[k]);
    p.reducers[k] = reducers[k];
  }
  p.plan = static_cast<const SpmmPlan*>(plan);
  const int vec = pick_vec(p, F);  // slab starts k*F keep the alignment of F
  if (arg_out != nullptr) return dispatch_shape<R_RUNTIME, false, true>(p, F, vec, K, INT64_MAX, as_stream(stream));
  return dispatch_shape<R_RUNTIME, false, false>(p, F, vec, K, INT64_MAX, as_stream(stream));
}

extern "C" int incagg_spmm_multi(const int32_t* rowptr, const int32_t* col, const float* val,
                                 const float* X, int64_t ldx, float* out, int64_t ldo, int64_t rows,
                                 int32_t F, int32_t K, const int32_t* reducers, const void* plan,
                                 incagg_stream_t stream) {
  return spmm_multi_impl(rowptr, col, val, X, ldx, out, ldo, nullptr, 0, rows, F, K, reducers, plan, stream);
}

extern "C" int incagg_spmm_multi_arg(const int32_t* rowptr, const int32_t* col, const float* val,
                                     const float* X, int64_t ldx, float* out, int64_t ldo, int32_t* arg_out,
                                     int64_t lda, int64_t rows, int32_t F, int32_t K, const int32_t* reducers,
                                     const void* plan, incagg_stream_t stream) {
  IA_CHECK_ARG(arg_out != nullptr, "arg_out is NULL");
  return spmm_multi_impl(rowptr, col, val, X, ldx, out, ldo, arg_out, lda, rows, F, K, reducers, plan, stream);
}

extern "C" int incagg_spmm_minmax_bwd(const int32_t* col, const float* val, const int32_t* arg,
                                      int64_t lda, const float* grad_out, int64_t ldg,
                                      float* grad_x, int64_t ldx, int64_t rows, int32_t F,
                                      incagg_stream_t stream) {
  IA_CHECK_ARG(rows >= 0 && F >= 0, "negative size");
  if (rows == 0 || F == 0) return INCAGG_OK;
  IA_CHECK_ARG(col && arg && grad_out && grad_x, "NULL argument");
  const int64_t total = rows * F;
  const int threads = 256;
  const int64_t want = (total + threads - 1) / threads;
  const int blocks = (int)(want < (int64_t)sm_count() * 16 ? want : (int64_t)sm_count() * 16);
  launch(minmax_bwd_kernel, dim3(blocks), dim3(threads), (size_t)(0), as_stream(stream), col, val, arg, lda, grad_out, ldg,
                                                               grad_x, ldx, rows, F);
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}
