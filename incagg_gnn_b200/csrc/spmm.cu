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
#include <float.h>
#include <stdlib.h>
#include <limits.h>

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
  int32_t pad[3];
  // followed by int4 items[capacity]: {row, chunk index, scratch base or -1, number of chunks}
};
constexpr int PART_STRIDE = 512;        // floats per scratch slot (= widest tile, 32 lanes * 4 * 4)
constexpr int PART_SLOTS = 16384;       // scratch slots per device (32 MB values + 32 MB args)
constexpr int PLAN_SCRATCH_CAP = 8192;  // slots one plan may hand out (times the feature tiles)

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
template <int REDUCE, int VEC, int G, int NCH, bool DELTA, bool ARG>
__device__ __forceinline__ void walk_edges(const SpmmParams& p, int op, int s, int e, int fbase,
                                           int F, int lane_g, unsigned gmask,
                                           float (&acc)[NCH][VEC], int32_t (&arg)[NCH][VEC]) {
  constexpr int UNROLL = (NCH >= 4) ? 2 : 4;
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
        const float* xr = p.X + (int64_t)c[u] * p.ldx;
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
          const int f = fbase + (k * G + lane_g) * VEC;
          if (f < F) {
            load_vec<VEC>(xr + f, x[u][k]);
          } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) x[u][k][i] = 0.f;
          }
        }
      }
      if constexpr (DELTA) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const int64_t g = p.n_id ? p.n_id[c[u]] : (int64_t)c[u];
          const float* mr = p.m_in + g * p.ld_in;
#pragma unroll
          for (int k = 0; k < NCH; ++k) {
            const int f = fbase + (k * G + lane_g) * VEC;
            if (f < F) {
              float m[VEC];
              load_vec<VEC>(mr + f, m);
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
      const float* xr = p.X + (int64_t)c * p.ldx;
      const float* mr = nullptr;
      if constexpr (DELTA) {
        const int64_t g = p.n_id ? p.n_id[c] : (int64_t)c;
        mr = p.m_in + g * p.ld_in;
      }
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const int f = fbase + (k * G + lane_g) * VEC;
        if (f < F) {
          float x[VEC];
          load_vec<VEC>(xr + f, x);
          if constexpr (DELTA) {
            float m[VEC];
            load_vec<VEC>(mr + f, m);
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
#ifndef SPMM_MIN_CTAS
#define SPMM_MIN_CTAS (1536 / SPMM_THREADS_N)
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
    constexpr int GROUPS = SPMM_THREADS / G;
    const int lane_g = threadIdx.x % G;
    const int sub = lane / G;  // group index inside the warp
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (sub * G));
    const int64_t row = (int64_t)(blockIdx.x - p.long_grid) * GROUPS + threadIdx.x / G;
    if (row >= p.rows) return;
    const int s = __ldg(p.rowptr + row), e = __ldg(p.rowptr + row + 1);
    if (e - s > long_row) return;  // owned by the plan's items
    float acc[NCH][VEC];
    int32_t arg[NCH][VEC];
#pragma unroll
    for (int k = 0; k < NCH; ++k)
#pragma unroll
      for (int i = 0; i < VEC; ++i) { acc[k][i] = red_init<REDUCE>(op); arg[k][i] = -1; }
    walk_edges<REDUCE, VEC, G, NCH, DELTA, ARG>(p, op, s, e, fbase, flim, lane_g, gmask, acc, arg);
    finish_row<REDUCE, VEC, G, NCH, DELTA, ARG>(p, op, row, e - s, fbase, flim, lane_g, acc, arg);
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

// Plan construction: one thread per row appends the row's work items.
__global__ void spmm_plan_kernel(const int32_t* __restrict__ rowptr, int64_t rows, SpmmPlan* plan,
                                 int4* __restrict__ items, int long_row, int chunk, int capacity) {
  pdl_prologue();
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (chunk <= 0) {
    // adaptive chunk: small enough that one CTA trip is a few microseconds on batch-sized
    // structures, large enough that the split rows of a whole graph fit the scratch area
    const int64_t nnz = (int64_t)rowptr[rows] - rowptr[0];
    chunk = 256;
    while (chunk < 4096 && (int64_t)chunk * (PLAN_SCRATCH_CAP / 2) < nnz) chunk <<= 1;
  }
  if (row == 0) {
    plan->long_row = long_row;
    plan->chunk = chunk;
    plan->capacity = capacity;
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

static int build_plan(const int32_t* rowptr, int64_t rows, void* plan, int64_t capacity,
                      cudaStream_t st) {
  IA_CHECK_ARG(capacity < 0x7fffffff, "plan too large");
  IA_CUDA(cudaMemsetAsync(plan, 0, sizeof(SpmmPlan), st));
  if (rows == 0) return INCAGG_OK;
  SpmmPlan* hdr = static_cast<SpmmPlan*>(plan);
  int4* items = reinterpret_cast<int4*>(hdr + 1);
  launch(spmm_plan_kernel, dim3((unsigned)((rows + 255) / 256)), dim3(256), (size_t)(0), st, rowptr, rows, hdr, items,
                                                                  long_row_edges(), chunk_edges(),
                                                                  (int)capacity);
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}

struct SpmmScratch {
  float* part_val = nullptr;
  int32_t* part_arg = nullptr;
  int32_t* part_done = nullptr;
  void* plan = nullptr;  // temporary plan of calls that pass none
  int64_t plan_capacity = 0;
};
// Per-device scratch (allocated once / grown on demand, reused by every call: kernels of one stream
// serialise).  Concurrent SpMM calls on different streams of one device must not share a thread.
static int get_scratch(SpmmScratch** out, int64_t want_plan_capacity) {
  static thread_local SpmmScratch scratch[16];
  int dev = 0;
  IA_CUDA(cudaGetDevice(&dev));
  IA_CHECK_ARG(dev >= 0 && dev < 16, "device ordinal %d out of range", dev);
  SpmmScratch& s = scratch[dev];
  if (s.part_val == nullptr) {
    IA_CUDA(cudaMalloc(&s.part_val, sizeof(float) * (size_t)PART_SLOTS * PART_STRIDE));
    IA_CUDA(cudaMalloc(&s.part_arg, sizeof(int32_t) * (size_t)PART_SLOTS * PART_STRIDE));
    IA_CUDA(cudaMalloc(&s.part_done, sizeof(int32_t) * (size_t)PART_SLOTS));
    IA_CUDA(cudaMemset(s.part_done, 0, sizeof(int32_t) * (size_t)PART_SLOTS));
  }
  if (want_plan_capacity > s.plan_capacity) {
    int64_t cap = s.plan_capacity ? s.plan_capacity : (1 << 16);
    while (cap < want_plan_capacity) cap <<= 1;
    if (s.plan) IA_CUDA(cudaFree(s.plan));  // synchronises the device: earlier users are done
    s.plan = nullptr;
    s.plan_capacity = 0;
    IA_CUDA(cudaMalloc(&s.plan, sizeof(SpmmPlan) + sizeof(int4) * (size_t)cap));
    s.plan_capacity = cap;
  }
  *out = &s;
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
  // CTAs that walk the plan's items: never more than the items can be, at most 8 per SM
  int64_t lg = (int64_t)env_int("INCAGG_SPMM_LONG_CTAS", 8) * sm_count();
  if (items_bound < lg) lg = items_bound;
  if (lg < 1) lg = 1;
  p.long_grid = (int)lg;
  IA_CHECK_ARG(blocks + lg <= 0x7fffffff, "too many rows for one launch");
  dim3 grid((unsigned)(blocks + lg), (unsigned)n_tiles);
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
  return sizeof(SpmmPlan) + sizeof(int4) * (size_t)plan_capacity(rows, nnz);
}

extern "C" int incagg_spmm_plan(const int32_t* rowptr, int64_t rows, int64_t nnz, void* plan,
                                size_t plan_bytes, incagg_stream_t stream) {
  IA_CHECK_ARG(rows >= 0, "negative size");
  IA_CHECK_ARG(plan != nullptr && plan_bytes >= incagg_spmm_plan_bytes(rows, nnz), "plan buffer too small");
  IA_CHECK_ARG(rows == 0 || rowptr != nullptr, "rowptr is NULL");
  IA_CHECK_ARG((reinterpret_cast<uintptr_t>(plan) & 15) == 0, "plan buffer must be 16-byte aligned");
  return build_plan(rowptr, rows, plan, (int64_t)((plan_bytes - sizeof(SpmmPlan)) / sizeof(int4)),
                    as_stream(stream));
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
  if (reduce == R_SUM) return dispatch_shape<R_SUM, true, false>(p, F, vec, 1, INT64_MAX, st);
  return dispatch_shape<R_MEAN, true, false>(p, F, vec, 1, INT64_MAX, st);
}

extern "C" int incagg_spmm_multi(const int32_t* rowptr, const int32_t* col, const float* val,
                                 const float* X, int64_t ldx, float* out, int64_t ldo, int64_t rows,
                                 int32_t F, int32_t K, const int32_t* reducers, const void* plan,
                                 incagg_stream_t stream) {
  IA_CHECK_ARG(K >= 1 && K <= 16, "K must be in [1, 16] (got %d)", K);
  IA_CHECK_ARG(reducers != nullptr, "reducers is NULL");
  int rc = check_common(rowptr, col, X, out, ldx, ldo, rows, F * K);
  if (rc != INCAGG_OK) return rc;
  if (rows == 0 || F == 0) return INCAGG_OK;
  SpmmParams p{};
  p.rowptr = rowptr; p.col = col; p.val = val; p.X = X; p.ldx = ldx; p.out = out; p.ldo = ldo;
  p.rows = rows; p.F = F * K; p.slab_F = F;
  for (int k = 0; k < K; ++k) {
    IA_CHECK_ARG(reducers[k] >= 0 && reducers[k] <= 3, "unknown reducer %d", reducers[k]);
    p.reducers[k] = reducers[k];
  }
  p.plan = static_cast<const SpmmPlan*>(plan);
  const int vec = pick_vec(p, F);  // slab starts k*F keep the alignment of F
  return dispatch_shape<R_RUNTIME, false, false>(p, F, vec, K, INT64_MAX, as_stream(stream));
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
