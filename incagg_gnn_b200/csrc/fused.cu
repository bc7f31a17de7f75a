// Small fused kernels around the GEMMs of the training step (sm_100a): what PyTorch would issue as
// ~20 tiny launches (mask cast, sums, log-softmax, NLL, their backward) or as a slow strided
// reduction becomes one or two launches each.
//
//   incagg_relu_bwd_colsum   gm = g * (y > 0)  and  colsum[c] = sum_r gm[r, c]   (ReLU backward of the
//                            first Linear fused with its bias gradient; reference gcn2.py:87 lins[0])
//   incagg_masked_ce         mean cross-entropy over the rows selected by a mask, and d loss / d logits
//                            (reference main.py:80  criterion(out[train_mask], y[train_mask]))
#include <float.h>

#include "common.cuh"

namespace incagg {

constexpr int CS_THREADS = 256;
constexpr int CS_ROWS_PER_BLOCK = 64;   // (256: 340 blocks for the layer-0 gradient = 2.3 per SM, a long unbalanced tail)

// Each block owns CS_ROWS_PER_BLOCK rows; thread t owns column chunk (t % cvec) of row group
// (t / cvec), walks its rows with a fixed stride and the block combines through shared memory in a
// fixed order -> partial[block][cols].  V = 4: float4 along the row (cols % 4 == 0, 16-byte aligned
// rows); V = 1: scalar columns (any layout, e.g. the 47 logits of the classifier head).
// add (nullable): a second gradient that is added to the first add_rows rows of g before the mask (the
// x_0 gradients collected by the layers' GEMM epilogues, nn.X0GradSink).
template <int V>
__global__ void __launch_bounds__(CS_THREADS)
relu_bwd_colsum_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ y, int64_t ldy,
                       int64_t rows, int cols, float* __restrict__ gm, int64_t ldo,
                       const float* __restrict__ add, int64_t ldadd, int64_t add_rows,
                       float* __restrict__ partial) {
  pdl_prologue();
  extern __shared__ float sm[];  // [groups][cols]
  const int cvec = cols / V;
  const int groups = CS_THREADS / cvec;
  const int grp = threadIdx.x / cvec, cv = threadIdx.x % cvec;
  const int64_t r0 = (int64_t)blockIdx.x * CS_ROWS_PER_BLOCK;
  const int64_t r1 = min(rows, r0 + CS_ROWS_PER_BLOCK);
  if (grp < groups) {
    if constexpr (V == 4) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int64_t r = r0 + grp; r < r1; r += groups) {
        float4 v = *reinterpret_cast<const float4*>(g + r * ldg + cv * 4);
        if (add && r < add_rows) {
          const float4 a = *reinterpret_cast<const float4*>(add + r * ldadd + cv * 4);
          v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        }
        if (y) {
          const float4 m = *reinterpret_cast<const float4*>(y + r * ldy + cv * 4);
          v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f;
          v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
        }
        if (gm && (y || add)) *reinterpret_cast<float4*>(gm + r * ldo + cv * 4) = v;
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(sm + grp * cols + cv * 4) = acc;
    } else {
      float acc = 0.f;
      for (int64_t r = r0 + grp; r < r1; r += groups) {
        float v = g[r * ldg + cv];
        if (add && r < add_rows) v += add[r * ldadd + cv];
        if (y) v = y[r * ldy + cv] > 0.f ? v : 0.f;
        if (gm && (y || add)) gm[r * ldo + cv] = v;
        acc += v;
      }
      sm[grp * cols + cv] = acc;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += CS_THREADS) {
    float s = 0.f;
    for (int k = 0; k < groups; ++k) s += sm[k * cols + c];
    partial[(int64_t)blockIdx.x * cols + c] = s;
  }
}

// out[c] = sum_b partial[b][c]: a block of 8 warps owns 32 columns, warp w sums the blocks w, w + 8, ...
// (independent coalesced loads), the eight sums are combined in warp order: a fixed association.
// (One thread per column walking all the blocks took 17 us for the 340 blocks of the layer-0 gradient.)
constexpr int CF_WARPS = 32;
__global__ void __launch_bounds__(CF_WARPS * 32)
colsum_finish_kernel(const float* __restrict__ partial, int nblocks, int cols, float* __restrict__ out,
                     int accumulate) {
  pdl_prologue();
  __shared__ float part[CF_WARPS][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (c < cols) {
    const float* src = partial + c;
    int b = w;
    for (; b + 3 * CF_WARPS < nblocks; b += 4 * CF_WARPS) {
      const float a0 = src[(int64_t)b * cols], a1 = src[(int64_t)(b + CF_WARPS) * cols];
      const float a2 = src[(int64_t)(b + 2 * CF_WARPS) * cols], a3 = src[(int64_t)(b + 3 * CF_WARPS) * cols];
      s += a0; s += a1; s += a2; s += a3;
    }
    for (; b < nblocks; b += CF_WARPS) s += src[(int64_t)b * cols];
  }
  part[w][lane] = s;
  __syncthreads();
  if (w == 0 && c < cols) {
    float t = part[0][lane];
#pragma unroll
    for (int i = 1; i < CF_WARPS; ++i) t += part[i][lane];
    out[c] = accumulate ? out[c] + t : t;
  }
}

// ---- masked cross-entropy ---------------------------------------------------------------------------
__global__ void mask_count_kernel(const uint8_t* __restrict__ mask, int64_t n, float* __restrict__ count) {
  pdl_prologue();
  __shared__ int s_tot;
  if (threadIdx.x == 0) s_tot = 0;
  __syncthreads();
  int c = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) c += mask[i] ? 1 : 0;
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_tot, c);  // integer: order-independent
  __syncthreads();
  if (threadIdx.x == 0) *count = (float)s_tot;
}

// One warp per row: log-sum-exp, loss_i = w_i (lse - z_y), dlogits = w_i / max(n,1) (softmax - onehot).
// Per-block loss partials (fixed order inside the block) -> partial[block].
constexpr int CE_THREADS = 256;
__global__ void __launch_bounds__(CE_THREADS)
masked_ce_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ y,
                 const uint8_t* __restrict__ mask, int64_t rows, int C, const float* __restrict__ count,
                 float* __restrict__ dlogits, int64_t ldd, float* __restrict__ partial) {
  pdl_prologue();
  __shared__ float s_loss[CE_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * (CE_THREADS / 32) + w;
  const float inv_n = 1.f / fmaxf(*count, 1.f);
  float loss = 0.f;
  if (row < rows) {
    const float* z = logits + row * ld;
    float mx = -FLT_MAX;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += __expf(z[c] - mx);
    se = warp_sum(se);
    const float lse = mx + __logf(se);
    const bool on = mask[row] != 0;
    const int64_t t = y[row];
    if (on) loss = lse - z[t];
    const float wgt = on ? inv_n : 0.f;
    for (int c = lane; c < C; c += 32) {
      const float p = __expf(z[c] - lse);
      dlogits[row * ldd + c] = wgt * (p - (c == t ? 1.f : 0.f));
    }
  }
  if (lane == 0) s_loss[w] = loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < CE_THREADS / 32; ++k) s += s_loss[k];
    partial[blockIdx.x] = s;
  }
}

__global__ void ce_finish_kernel(const float* __restrict__ partial, int nblocks, const float* __restrict__ count,
                                 float* __restrict__ out /* [3]: loss sum, mean loss, count */,
                                 double* __restrict__ running /* nullable [2]: += loss sum, += count */) {
  pdl_prologue();
  // one block, fixed order
  __shared__ double s[256];
  double acc = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) acc += (double)partial[b];
  s[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < blockDim.x; ++k) t += s[k];
    const float n = *count;
    out[0] = (float)t;
    out[1] = (float)(t / (double)fmaxf(n, 1.f));
    out[2] = n;
    if (running) { running[0] += (double)(float)t; running[1] += (double)n; }
  }
}

}  // namespace incagg

using namespace incagg;

extern "C" size_t incagg_colsum_workspace_bytes(int64_t rows, int32_t cols) {
  if (rows <= 0 || cols <= 0) return 0;
  return sizeof(float) * (size_t)((rows + CS_ROWS_PER_BLOCK - 1) / CS_ROWS_PER_BLOCK) * (size_t)cols;
}

extern "C" int incagg_relu_bwd_colsum_ex(const float* g, int64_t ldg, const float* y, int64_t ldy, int64_t rows,
                                         int32_t cols, float* gm, int64_t ldo, const float* add, int64_t ldadd,
                                         int64_t add_rows, float* colsum, int accumulate, void* workspace,
                                         size_t workspace_bytes, incagg_stream_t stream) {
  IA_CHECK_ARG(rows >= 0 && cols > 0, "bad size");
  IA_CHECK_ARG(colsum != nullptr, "colsum is NULL");
  cudaStream_t st = as_stream(stream);
  if (rows == 0) {
    if (!accumulate) IA_CUDA(cudaMemsetAsync(colsum, 0, sizeof(float) * (size_t)cols, st));
    return INCAGG_OK;
  }
  IA_CHECK_ARG(g != nullptr, "g is NULL");
  IA_CHECK_ARG(add == nullptr || (add_rows >= 0 && add_rows <= rows && ldadd >= cols), "bad addend");
  IA_CHECK_ARG((y == nullptr && add == nullptr) || gm != nullptr, "gm is NULL");
  IA_CHECK_ARG(workspace != nullptr && workspace_bytes >= incagg_colsum_workspace_bytes(rows, cols),
               "workspace too small");
  const int nblocks = (int)((rows + CS_ROWS_PER_BLOCK - 1) / CS_ROWS_PER_BLOCK);
  float* partial = static_cast<float*>(workspace);
  if (add == nullptr) add_rows = 0;
  // float4 path: cols % 4 == 0 and every operand 16-byte aligned with ld % 4 == 0; else scalar columns
  const bool vec = cols % 4 == 0 && cols <= 1024 && ldg % 4 == 0 && aligned16(g) &&
                   (y == nullptr || (ldy % 4 == 0 && aligned16(y))) &&
                   (gm == nullptr || (ldo % 4 == 0 && aligned16(gm))) &&
                   (add == nullptr || (ldadd % 4 == 0 && aligned16(add)));
  if (vec) {
    const int groups = CS_THREADS / (cols / 4);
    launch(relu_bwd_colsum_kernel<4>, dim3(nblocks), dim3(CS_THREADS), (size_t)(sizeof(float) * (size_t)groups * cols), st,
        g, ldg, y, ldy, rows, cols, gm, ldo, add, ldadd, add_rows, partial);
  } else {
    IA_CHECK_ARG(cols <= CS_THREADS, "unaligned / ragged layouts support at most 256 columns");
    const int groups = CS_THREADS / cols;
    launch(relu_bwd_colsum_kernel<1>, dim3(nblocks), dim3(CS_THREADS), (size_t)(sizeof(float) * (size_t)groups * cols), st,
        g, ldg, y, ldy, rows, cols, gm, ldo, add, ldadd, add_rows, partial);
  }
  IA_LAUNCH_CHECK();
  launch(colsum_finish_kernel, dim3((cols + 31) / 32), dim3(CF_WARPS * 32), (size_t)(0), st, partial, nblocks, cols, colsum,
         accumulate ? 1 : 0);
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}

extern "C" int incagg_relu_bwd_colsum(const float* g, int64_t ldg, const float* y, int64_t ldy, int64_t rows,
                                      int32_t cols, float* gm, int64_t ldo, float* colsum, void* workspace,
                                      size_t workspace_bytes, incagg_stream_t stream) {
  return incagg_relu_bwd_colsum_ex(g, ldg, y, ldy, rows, cols, gm, ldo, nullptr, 0, 0, colsum, 0, workspace,
                                   workspace_bytes, stream);
}

extern "C" size_t incagg_masked_ce_workspace_bytes(int64_t rows) {
  if (rows <= 0) return 16;
  return sizeof(float) * (size_t)((rows + CE_THREADS / 32 - 1) / (CE_THREADS / 32)) + 16;
}

extern "C" int incagg_masked_ce(const float* logits, int64_t ld, const int64_t* y, const uint8_t* mask,
                                int64_t rows, int32_t C, float* dlogits, int64_t ldd, float* out3,
                                void* workspace, size_t workspace_bytes, incagg_stream_t stream) {
  IA_CHECK_ARG(rows >= 0 && C > 0, "bad size");
  IA_CHECK_ARG(out3 != nullptr, "out is NULL");
  IA_CHECK_ARG(workspace != nullptr && workspace_bytes >= incagg_masked_ce_workspace_bytes(rows),
               "workspace too small");
  cudaStream_t st = as_stream(stream);
  // out3 = {loss sum, mean loss, count}
  float* count = out3 + 2;
  if (rows == 0) {
    IA_CUDA(cudaMemsetAsync(out3, 0, 3 * sizeof(float), st));
    return INCAGG_OK;
  }
  IA_CHECK_ARG(logits && y && mask && dlogits, "NULL argument");
  launch(mask_count_kernel, dim3(1), dim3(1024), (size_t)(0), st, mask, rows, count);
  IA_LAUNCH_CHECK();
  const int nblocks = (int)((rows + CE_THREADS / 32 - 1) / (CE_THREADS / 32));
  float* partial = static_cast<float*>(workspace);
  launch(masked_ce_kernel, dim3(nblocks), dim3(CE_THREADS), (size_t)(0), st, logits, ld, y, mask, rows, C, count, dlogits, ldd, partial);
  IA_LAUNCH_CHECK();
  launch(ce_finish_kernel, dim3(1), dim3(256), (size_t)(0), st, partial, nblocks, count, out3, (double*)nullptr);
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}

// The same three kernels as separate calls, so that a training step can keep only the middle one on its
// critical path: the count depends on the mask alone (side stream, while the forward pass runs) and the
// loss value is not needed by the backward pass (side stream, while it runs).
extern "C" int incagg_mask_count(const uint8_t* mask, int64_t rows, float* count, incagg_stream_t stream) {
  IA_CHECK_ARG(rows >= 0 && count != nullptr && (rows == 0 || mask != nullptr), "bad argument");
  launch(mask_count_kernel, dim3(1), dim3(1024), (size_t)(0), as_stream(stream), mask, rows, count);
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}

extern "C" int incagg_masked_ce_rows(const float* logits, int64_t ld, const int64_t* y, const uint8_t* mask,
                                     int64_t rows, int32_t C, const float* count, float* dlogits, int64_t ldd,
                                     void* workspace, size_t workspace_bytes, incagg_stream_t stream) {
  IA_CHECK_ARG(rows >= 0 && C > 0 && count != nullptr, "bad argument");
  IA_CHECK_ARG(workspace != nullptr && workspace_bytes >= incagg_masked_ce_workspace_bytes(rows),
               "workspace too small");
  if (rows == 0) return INCAGG_OK;
  IA_CHECK_ARG(logits && y && mask && dlogits, "NULL argument");
  const int nblocks = (int)((rows + CE_THREADS / 32 - 1) / (CE_THREADS / 32));
  launch(masked_ce_kernel, dim3(nblocks), dim3(CE_THREADS), (size_t)(0), as_stream(stream), logits, ld, y, mask, rows,
         C, count, dlogits, ldd, static_cast<float*>(workspace));
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}

extern "C" int incagg_masked_ce_finish(const void* workspace, int64_t rows, const float* count, float* out3,
                                       double* acc, incagg_stream_t stream) {
  IA_CHECK_ARG(rows >= 0 && count != nullptr && out3 != nullptr && (rows == 0 || workspace != nullptr), "bad argument");
  const int nblocks = (int)((rows + CE_THREADS / 32 - 1) / (CE_THREADS / 32));
  launch(ce_finish_kernel, dim3(1), dim3(256), (size_t)(0), as_stream(stream), static_cast<const float*>(workspace),
         nblocks, count, out3, acc);
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}

// ---- fused Adam over flat buffers ---------------------------------------------------------------------
// torch.optim.Adam (main.py:196-201: two parameter groups that differ in weight decay only) issues
// ~12 multi-tensor launches per step; with all parameters / gradients / moments in flat buffers the
// update is one launch.  Same arithmetic as torch's (non-amsgrad, L2 weight decay added to the gradient):
//   g += wd p ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
//   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The step counter lives on the device (CUDA-graph friendly): the kernel reads step[0] + 1 and a
// single thread of the last block stores it back after the grid has read it (arrival counter).
namespace incagg {
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, int64_t n_decay, float lr, float b1, float b2,
                            float eps, float wd_first, float wd_rest, float* step, unsigned int* arrivals) {
  pdl_prologue();
  const float t = step[0] + 1.f;
  const float bc1 = 1.f - powf(b1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, t));
  const float step_size = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = g[i];
    const float wd = i < n_decay ? wd_first : wd_rest;
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float m0 = m[i];
    const float mi = m0 + (1.f - b1) * (gi - m0);          // torch: exp_avg.lerp_(grad, 1 - beta1)
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
  }
  // every block has read `step` before the last arriver overwrites it
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(arrivals, 1u);
    if (prev == gridDim.x - 1) {
      step[0] = t;
      *arrivals = 0u;
    }
  }
}
}  // namespace incagg

extern "C" int incagg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                int64_t n_first_group, float lr, float beta1, float beta2, float eps,
                                float wd_first_group, float wd_rest, float* step_dev, void* arrivals_dev,
                                incagg_stream_t stream) {
  IA_CHECK_ARG(n >= 0 && n_first_group >= 0 && n_first_group <= n, "bad sizes");
  if (n == 0) return INCAGG_OK;
  IA_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && step_dev && arrivals_dev, "NULL argument");
  const int threads = 256;
  int64_t want = (n + threads - 1) / threads;
  const int blocks = (int)(want < (int64_t)sm_count() * 4 ? want : (int64_t)sm_count() * 4);
  launch(incagg::adam_kernel, dim3(blocks), dim3(threads), (size_t)(0), as_stream(stream), 
      params, grads, exp_avg, exp_avg_sq, n, n_first_group, lr, beta1, beta2, eps, wd_first_group, wd_rest,
      step_dev, static_cast<unsigned int*>(arrivals_dev));
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}


// ---- gradient all-reduce over NVLink peer memory fused with the Adam update ---------------------------
//
// Data-parallel step on W GPUs of one NVSwitch box: the gradient buffers are small (GCNII 5 x 128:
// 0.2 M floats), so a ring / tree collective is all latency.  One kernel per rank does the whole
// exchange and the optimizer update:
//   1. block b copies its slice of the local gradient into a staging buffer (two buffers, selected by
//      the parity of the step number) and publishes "slice b of step t is staged" in every peer's
//      signal array with a system-scope release store;
//   2. it waits until every peer has published the same slice (acquire loads on its own signal array,
//      which the peers write over NVLink), then reads the peers' staged slices straight out of their
//      HBM (volatile loads: the lines of step t-2 may still sit in this SM's L1);
//   3. the contributions are added IN RANK ORDER (bit-identical sums on every rank, so the replicas
//      stay identical without a broadcast), divided by W and fed to the Adam update of the slice.
// There is no barrier at the end: a staging buffer is rewritten at step t + 2, which a rank can only
// reach after every peer has entered step t + 1, i.e. (stream order) finished reading step t.
// The blocks of one launch do not depend on each other (each slice has its own signals), so no
// co-residency of the grid is assumed; ranks run on different GPUs (one process per GPU).
// A wait that lasts longer than ~10 s sets INCAGG_DEVERR_PEER_TIMEOUT and proceeds instead of hanging.
namespace incagg {
constexpr int AR_MAX_RANKS = 16;
constexpr int AR_BLOCKS = 32;
struct ArPeers {
  float* stage[AR_MAX_RANKS];     // [2][n] staging buffers of every rank (own + peers, IPC-mapped)
  int32_t* signal[AR_MAX_RANKS];  // [AR_BLOCKS][AR_MAX_RANKS] signal arrays of every rank
};

__device__ __forceinline__ void st_release_sys(int32_t* p, int32_t v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int32_t ld_acquire_sys(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
allreduce_adam_kernel(const ArPeers peers, int rank, int world, float* __restrict__ g, float* __restrict__ p,
                      float* __restrict__ m, float* __restrict__ v, int64_t n, int64_t n_decay, float lr,
                      float b1, float b2, float eps, float wd_first, float wd_rest, float* step,
                      unsigned int* arrivals, int32_t* err) {
  pdl_prologue();
  const float t = step[0] + 1.f;
  const int epoch = (int)t;
  const int64_t par = (int64_t)(epoch & 1) * n;
  // slice of this block (multiples of 4 elements)
  int64_t per = (n + gridDim.x - 1) / gridDim.x;
  per = (per + 3) / 4 * 4;
  const int64_t lo = min(n, (int64_t)blockIdx.x * per), hi = min(n, lo + per);
  float* mine = peers.stage[rank] + par;
  if ((n % 4 == 0) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
    for (int64_t i = lo + 4 * (int64_t)threadIdx.x; i < hi; i += 4 * (int64_t)blockDim.x)
      *reinterpret_cast<float4*>(mine + i) = *reinterpret_cast<const float4*>(g + i);
  } else {
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) mine[i] = g[i];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
    st_release_sys(peers.signal[threadIdx.x] + blockIdx.x * AR_MAX_RANKS + rank, epoch);
    const int32_t* flag = peers.signal[rank] + blockIdx.x * AR_MAX_RANKS + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < epoch) {
      if (clock64() - t0 > (1ll << 34)) {  // ~10 s: a peer is gone; do not hang the GPU
        if (err) atomicOr(err, INCAGG_DEVERR_PEER_TIMEOUT);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  const float bc1 = 1.f - powf(b1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, t));
  const float step_size = lr / bc1;
  const float inv_w = 1.f / (float)world;
  // four elements per thread and trip; the W peer loads of a trip are issued back to back (volatile
  // 128-bit loads over NVLink) before any of them is used
  auto adam = [&](int64_t i, float gi) {
    g[i] = gi;  // the averaged gradient stays readable (p.grad)
    float pi = p[i];
    const float wd = i < n_decay ? wd_first : wd_rest;
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float m0 = m[i];
    const float mi = m0 + (1.f - b1) * (gi - m0);
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
  };
  const bool vec_ok = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0);
  if (vec_ok) {
    for (int64_t i = lo + 4 * (int64_t)threadIdx.x; i < hi; i += 4 * (int64_t)blockDim.x) {
      float4 c[AR_MAX_RANKS];
#pragma unroll
      for (int r = 0; r < AR_MAX_RANKS; ++r) {
        if (r < world) {
          c[r] = (r == rank) ? *reinterpret_cast<const float4*>(g + i)
                             : __ldcv(reinterpret_cast<const float4*>(peers.stage[r] + par + i));
        }
      }
      float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < AR_MAX_RANKS; ++r) {
        if (r < world) { sum.x += c[r].x; sum.y += c[r].y; sum.z += c[r].z; sum.w += c[r].w; }
      }
      adam(i, sum.x * inv_w);
      adam(i + 1, sum.y * inv_w);
      adam(i + 2, sum.z * inv_w);
      adam(i + 3, sum.w * inv_w);
    }
  } else {
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      float gi = 0.f;
      for (int r = 0; r < world; ++r) gi += (r == rank) ? g[i] : __ldcv(peers.stage[r] + par + i);
      adam(i, gi * inv_w);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(arrivals, 1u);
    if (prev == gridDim.x - 1) {
      step[0] = t;
      *arrivals = 0u;
    }
  }
}
}  // namespace incagg

extern "C" int incagg_allreduce_adam_blocks(void) { return incagg::AR_BLOCKS; }
extern "C" int incagg_allreduce_adam_max_ranks(void) { return incagg::AR_MAX_RANKS; }

extern "C" int incagg_allreduce_adam_step(void* const* stage_ptrs, void* const* signal_ptrs, int rank, int world,
                                          float* grads, float* params, float* exp_avg, float* exp_avg_sq,
                                          int64_t n, int64_t n_first_group, float lr, float beta1, float beta2,
                                          float eps, float wd_first_group, float wd_rest, float* step_dev,
                                          void* arrivals_dev, incagg_stream_t stream) {
  using namespace incagg;
  IA_CHECK_ARG(world >= 1 && world <= AR_MAX_RANKS && rank >= 0 && rank < world, "bad rank / world size");
  IA_CHECK_ARG(n >= 0 && n_first_group >= 0 && n_first_group <= n, "bad sizes");
  if (n == 0) return INCAGG_OK;
  IA_CHECK_ARG(stage_ptrs && signal_ptrs && grads && params && exp_avg && exp_avg_sq && step_dev && arrivals_dev,
               "NULL argument");
  ArPeers peers;
  for (int r = 0; r < AR_MAX_RANKS; ++r) {
    peers.stage[r] = r < world ? static_cast<float*>(stage_ptrs[r]) : nullptr;
    peers.signal[r] = r < world ? static_cast<int32_t*>(signal_ptrs[r]) : nullptr;
    IA_CHECK_ARG(r >= world || (peers.stage[r] && peers.signal[r]), "rank %d: NULL staging / signal buffer", r);
  }
  launch(allreduce_adam_kernel, dim3(AR_BLOCKS), dim3(256), (size_t)(0), as_stream(stream), peers, rank, world, grads,
         params, exp_avg, exp_avg_sq, n, n_first_group, lr, beta1, beta2, eps, wd_first_group, wd_rest, step_dev,
         static_cast<unsigned int*>(arrivals_dev), device_error_word());
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}
