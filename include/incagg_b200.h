/*
 * incagg_b200.h — C ABI of the B200-native IncAgg-GNN propagation hot path.
 *
 * This is the drop-in boundary.  Every entry point takes plain
 * pointers and sizes (no torch types), a cudaStream_t passed as `void*`, returns
 * an `int` status (0 = ok, <0 = error, text via incagg_last_error()) and never
 * throws.  All work is enqueued on `stream`; nothing here synchronises the
 * device or the stream unless the comment says so.
 *
 * Reference interfaces replaced (paths relative to the reference repo):
 *   torch_sparse.matmul / SparseTensor.__matmul__ / torch_geometric.utils.spmm
 *       call sites: torch_geometric_autoscale/models/gcn.py:143,164,241,262,296,361,386,403
 *                   models/gcn2.py:130,142,255,305,336,454,477,499
 *                   models/appnp.py:85,89,122,130,152,253,284,306
 *                   models/graphsage.py:634,683,898,928,952   models/pna.py:75
 *   History.pull / History.push            torch_geometric_autoscale/history.py:33-65
 *   read_async / write_async / synchronize csrc/async.cpp:13-48, csrc/cuda/async_cuda.cu:12-165
 *   relabel_one_hop                        csrc/relabel.cpp:10-24, csrc/cpu/relabel_cpu.cpp:3-108
 *   relabel_one_hop_within_batch           csrc/relabel.cpp:27-38, csrc/cpu/relabel_cpu.cpp:111-214
 *
 * Index conventions: batch-local CSR structures (what the SpMM kernels read)
 * use int32 rowptr/col; the global graph handed to relabel uses int64 rowptr and
 * int32 or int64 col; node ids (`idx`, `n_id`) are int64 as in the reference.
 * Feature matrices are fp32 row-major with an explicit leading dimension (in
 * elements) so that history-table slices can be read in place.
 */
#ifndef INCAGG_B200_H
#define INCAGG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* incagg_stream_t; /* a cudaStream_t */

/* status codes */
#define INCAGG_OK 0
#define INCAGG_ERR_INVALID (-1)     /* bad argument (what the reference AT_ASSERTMs) */
#define INCAGG_ERR_CUDA (-2)        /* CUDA runtime error, text in incagg_last_error() */
#define INCAGG_ERR_UNSUPPORTED (-3) /* valid request this build does not implement */

/* reducers: torch_sparse.matmul(reduce=...) */
#define INCAGG_REDUCE_SUM 0
#define INCAGG_REDUCE_MEAN 1
#define INCAGG_REDUCE_MIN 2
#define INCAGG_REDUCE_MAX 3

/* ---- library --------------------------------------------------------- */
int incagg_version(void);
/* Thread-local text of the last error returned on this thread ("" if none). */
const char* incagg_last_error(void);
/* Number of CUDA kernels this library has launched in this process (every entry point counts
 * its own launches; DMA copies issued through cudaMemcpyAsync are not kernels and not counted). */
int64_t incagg_launch_count(void);
/* Experiment knobs of the kernels (not part of the reference-facing surface; the defaults are what the
 * product runs with).  Plans built before a knob that changes the warp partition was set must be
 * rebuilt. */
#define INCAGG_TUNE_SPMM_STREAM_VARIANT 0 /* merge-path SpMM: -2 auto (delta form only, variant 2), -1 off, else gathers per buffer x warps/SM: 0 = 8 x 16, 1 = 4 x 32, 2 = 4 x 24, 3 = 2 x 40, 4 = 2 x 48, 5 = 16 x 8 */
#define INCAGG_TUNE_SPMM_STREAM_MIN_F 1   /* smallest feature width routed to the merge-path kernel (65) */
#define INCAGG_TUNE_GEMM_DUAL_M_CTAS 2    /* CTA budget of an M-concatenated split-K GEMM (weight gradients that run beside the backward chain); default: the SM count */
#define INCAGG_TUNE_GEMM_BN64_MIN_TILES 3 /* 64-wide n-tiles (two CTAs per SM) for problems with at least this many m-tiles; default: never */
#define INCAGG_TUNE_COUNT 8
int incagg_tune_set(int key, int value);
/* Device-side error word of the current device.  The reference raises on an index outside its table
 * (index_select, emb[n_id] = x); kernels cannot raise, so they record the fact: a gather writes zeros
 * for the row, a scatter / relabel skips it, and the bit below is set.  Reading synchronises with the
 * device; `reset` != 0 clears the word.  (ops.check_device_errors() raises RuntimeError on non-zero.) */
#define INCAGG_DEVERR_ROW_INDEX 1 /* gather / scatter / sharded gather: row id outside the table */
#define INCAGG_DEVERR_NODE_ID 2   /* relabel: batch node id outside [0, num_nodes) */
#define INCAGG_DEVERR_PEER_TIMEOUT 4 /* fused all-reduce: a peer rank did not show up within ~10 s */
int incagg_device_errors(int32_t* out, int reset);
/* SM count and compute capability of the current device. */
int incagg_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Enable loads / stores from kernels of the current device to memory of `peer_device` (NVLink P2P;
 * the history shards of other ranks are mapped through CUDA IPC by the host side).  Idempotent. */
int incagg_enable_peer_access(int peer_device);

/* ---- CSR SpMM -------------------------------------------------------- */
/*
 * out[i, 0:F] = reduce_{e in [rowptr[i], rowptr[i+1])} val[e] * X[col[e], 0:F]
 * Replaces torch_sparse.matmul(adj_t, X, reduce) (SURVEY §8a "CSR SpMM").
 *   val      nullable (treated as 1.0)
 *   reduce   SUM / MEAN (sum / max(deg,1)) / MIN / MAX (empty row -> 0)
 *   arg_out  nullable; for MIN/MAX receives, per output element, the edge
 *            position e that won (first winner in CSR order), -1 for empty rows;
 *            leading dimension lda (elements)
 * The transposed call (backward, grad_X = A^T grad_out) is the same function on
 * the CSR of A^T (incagg_csr_transpose).
 *
 * `plan` (nullable) is the degree-bucket plan of this CSR structure built by incagg_spmm_plan: rows
 * longer than 64 edges are walked by whole CTAs, rows longer than 2048 edges by several CTAs whose
 * partials are combined in a fixed order (deterministic).  With plan == NULL a temporary plan is
 * built inside the call (one extra small launch); structures that are reused (all layers of a step,
 * forward and transposed backward) should build it once.  One kernel launch per call.
 */
size_t incagg_spmm_plan_bytes(int64_t rows, int64_t nnz /* -1 if unknown */);
int incagg_spmm_plan(const int32_t* rowptr, int64_t rows, int64_t nnz, void* plan /* 16-B aligned */,
                     size_t plan_bytes, incagg_stream_t stream);
int incagg_spmm_csr(int reduce, const int32_t* rowptr, const int32_t* col, const float* val,
                    const float* X, int64_t ldx, float* out, int64_t ldo, int32_t* arg_out,
                    int64_t lda, int64_t rows, int32_t F, const void* plan, incagg_stream_t stream);

/*
 * incagg_spmm_csr (sum / mean) whose epilogue zeroes out[r, f] where gate[r, f] <= 0.  Backward of a
 * layer whose ReLU was fused into the producer of X (the GEMM epilogue of the previous layer):
 * grad_X = (A^T grad_out) * [X > 0] in one launch instead of the SpMM plus torch's threshold_backward
 * (reference: the `.relu_()` after every conv, gcn2.py:139, and its autograd backward).
 */
int incagg_spmm_csr_gated(int reduce, const int32_t* rowptr, const int32_t* col, const float* val,
                          const float* X, int64_t ldx, float* out, int64_t ldo, int64_t rows, int32_t F,
                          const void* plan, const float* gate, int64_t ld_gate, incagg_stream_t stream);

/*
 * Fused incremental-aggregation update (reference: gcn.py:241, gcn2.py:255,
 * appnp.py:122, graphsage.py:634):
 *     out[i] = reduce_e val[e] * (x[col[e]] - M_in[g(col[e])]) + M_ag[g(i)]
 * where g(r) = n_id[r] when n_id != NULL (M_in / M_ag are whole history tables
 * indexed by global node id; this fuses the History pull) and g(r) = r otherwise
 * (M_in / M_ag are the already-pulled [B, >=F] slices).  reduce is SUM or MEAN
 * (MEAN divides the delta term by max(in-batch degree, 1), graphsage.py:634).
 */
int incagg_spmm_delta(int reduce, const int32_t* rowptr, const int32_t* col, const float* val,
                      const float* x, int64_t ldx, const float* m_in, int64_t ld_in,
                      const float* m_ag, int64_t ld_ag, const int64_t* n_id, float* out,
                      int64_t ldo, int64_t rows, int32_t F, const void* plan, incagg_stream_t stream);

/*
 * Backward of MIN/MAX: grad_X[col[arg[i,f]], f] += val[arg[i,f]] * grad_out[i,f]
 * (atomic scatter; grad_X must be zero-filled by the caller).
 */
int incagg_spmm_minmax_bwd(const int32_t* col, const float* val, const int32_t* arg, int64_t lda,
                           const float* grad_out, int64_t ldg, float* grad_x, int64_t ldx,
                           int64_t rows, int32_t F, incagg_stream_t stream);

/*
 * Fused multi-aggregator propagate for PNA (reference pna.py:66-84 runs K separate
 * matmul(reduce=aggr) passes).  X is [n_src, K*F]; slab k (columns k*F..(k+1)*F)
 * is reduced with reducers[k] (host array of K INCAGG_REDUCE_* codes, K <= 16);
 * the CSR structure is read once.  No arg output (forward / inference use).
 */
int incagg_spmm_multi(const int32_t* rowptr, const int32_t* col, const float* val, const float* X,
                      int64_t ldx, float* out, int64_t ldo, int64_t rows, int32_t F, int32_t K,
                      const int32_t* reducers, const void* plan, incagg_stream_t stream);
/* The same launch, also returning the winning edge of every min / max slab column in arg_out
 * ([rows, K*F] int32, -1 for empty rows; columns of sum / mean slabs are -1): what the backward pass
 * of a training step routes the min / max gradients with (incagg_spmm_minmax_bwd). */
int incagg_spmm_multi_arg(const int32_t* rowptr, const int32_t* col, const float* val, const float* X,
                          int64_t ldx, float* out, int64_t ldo, int32_t* arg_out, int64_t lda, int64_t rows,
                          int32_t F, int32_t K, const int32_t* reducers, const void* plan,
                          incagg_stream_t stream);

/* ---- dense feature transform on the tensor cores ----------------------- */
/*
 * D[M,N] = alpha * op(A)[M,K] * op(B)[K,N] + beta * Cin + bias[n]   (+ ReLU if relu != 0), fp32 in/out.
 * Replaces the cuBLAS SGEMMs behind torch.nn.Linear / torch.addmm on the path (gcn2.py:87,149;
 * GCN2Conv addmm; GCNConv.lin gcn.py:63; SAGEConv.lin_l/lin_r; appnp.py:79-83).  tcgen05.mma kind::tf32
 * with fp32 accumulation in TMEM and the error-compensated 3xTF32 operand split, so results stay
 * within ~1e-6 relative of an fp32 GEMM.
 *   transA = 0: A is [M,K] row-major (lda); 1: A is stored [K,M] row-major
 *   transB = 0: B is [K,N] row-major (ldb); 1: B is stored [N,K] row-major (a Linear weight)
 *   Cin / bias nullable.  relu is a flag word: bit 0 = ReLU; bit 2 (value 4) = Cin is a ReLU-backward gate
 *   instead of an addend, D = Cin > 0 ? alpha op(A) op(B) : 0 (the input gradient of a Linear whose input
 *   was a ReLU output, gcn2.py:149; no bias, beta ignored).
 *   workspace (nullable) enables deterministic split-K for long reductions
 *   (weight gradients); incagg_gemm_workspace_bytes gives a sufficient size.
 */
size_t incagg_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K);
int incagg_gemm_tf32x3(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A,
                       int64_t lda, const float* B, int64_t ldb, float alpha, const float* Cin,
                       int64_t ldcin, float beta, const float* bias, int relu, float* D, int64_t ldd,
                       void* workspace, size_t workspace_bytes, incagg_stream_t stream);

/*
 * Two GEMMs that share an operand in ONE launch (the dense half of a GCNII layer, reference
 * GCN2Conv: out = (1-b)((1-a)h + a x0) + b((1-a) h W1 + a x0 W2), and its gradients):
 *   mode 1 (K-concatenation)  D  = alpha * (A·(scaleB B) + A2·(scaleB2 B2)) + beta Cin + beta2 Cin2
 *                             A [M,K], A2 [M,K2]: forward  [h | x0] · [c1 W1 ; c2 W2]
 *   mode 2 (N-concatenation)  D  = alpha  * A·(scaleB  B ) + beta  Cin
 *                             D2 = alpha2 * A·(scaleB2 B2) + beta2 Cin2   (shared A): input gradients
 *                             (relu is a flag word here: bit 0 = ReLU; bit 1 = D2 += instead of D2 =, which
 *                             accumulates the x_0 gradient of the layers in the epilogue - every GCNII layer
 *                             reads x_0, gcn2.py:121)
 *   mode 3 (M-concatenation)  D  = alpha  * op(A )·B + beta Cin ,  D2 = alpha2 * op(A2)·B + beta2 Cin2
 *                             (shared B; split-K with the workspace): weight gradients  [h | x0]^T · g,
 *                             accumulated in place into the gradient buffers when Cin == D
 * Same layout flags, accuracy and epilogue rules as incagg_gemm_tf32x3; M2 = M, N2 = N.
 */
int incagg_gemm_tf32x3_dual(int mode, int transA, int transB, int64_t M, int64_t N, int64_t K, int64_t K2,
                            const float* A, int64_t lda, const float* A2, int64_t lda2, const float* B,
                            int64_t ldb, const float* B2, int64_t ldb2, float alpha, float alpha2,
                            float scaleB, float scaleB2, const float* Cin, int64_t ldcin, float beta,
                            const float* Cin2, int64_t ldcin2, float beta2, int relu, float* D, int64_t ldd,
                            float* D2, int64_t ldd2, void* workspace, size_t workspace_bytes,
                            incagg_stream_t stream);

/*
 * `count` (<= 16) independent GEMMs of ONE shape in one launch:  D[g] = alpha[g] * op(A[g]) op(B[g]) + beta * D[g]
 * (A, lda, B, ldb, alpha, D, ldd: HOST arrays of `count` entries; the operands are device pointers).  The
 * weight gradients of all GCNII layers of a step, [h_l | x_0]^T g_l (2 L problems of 128 x 128 x B,
 * GCN2Conv's two weights per layer, gcn2.py:121), are one split-K launch beside the end of the backward
 * pass instead of L launches that compete with its chain for the SMs.
 */
int incagg_gemm_tf32x3_group(int count, int transA, int transB, int64_t M, int64_t N, int64_t K,
                             const float* const* A, const int64_t* lda, const float* const* B,
                             const int64_t* ldb, const float* alpha, float beta, float* const* D,
                             const int64_t* ldd, void* workspace, size_t workspace_bytes,
                             incagg_stream_t stream);

/* ---- small fused kernels of the training step ------------------------------ */
/*
 * ReLU backward fused with the bias gradient of the Linear in front of it (gcn2.py:87 lins[0]):
 *   gm[r, c] = y[r, c] > 0 ? g[r, c] : 0 ;  colsum[c] = sum_r gm[r, c]   (fixed order: deterministic)
 * y == NULL: plain column sums of g (gm unused).  float4 path when cols % 4 == 0 (<= 1024) and every
 * operand is 16-byte aligned with ld % 4 == 0; any other layout runs scalar columns (cols <= 256, e.g. the
 * 47 logits of the classifier head, gcn2.py:149 lins[1]).
 */
size_t incagg_colsum_workspace_bytes(int64_t rows, int32_t cols);
int incagg_relu_bwd_colsum(const float* g, int64_t ldg, const float* y, int64_t ldy, int64_t rows,
                           int32_t cols, float* gm, int64_t ldo, float* colsum, void* workspace,
                           size_t workspace_bytes, incagg_stream_t stream);
/*
 * Extended form: `add` (nullable, [add_rows, cols]) is added to the first add_rows rows of g before the mask
 * (the x_0 gradients that the GCNII layers' GEMM epilogues collected, gcn2.py:121); accumulate != 0:
 * colsum += instead of colsum = (the bias gradient lands in the flat gradient buffer directly).
 * gm is required when y or add is given.
 */
int incagg_relu_bwd_colsum_ex(const float* g, int64_t ldg, const float* y, int64_t ldy, int64_t rows,
                              int32_t cols, float* gm, int64_t ldo, const float* add, int64_t ldadd,
                              int64_t add_rows, float* colsum, int accumulate, void* workspace,
                              size_t workspace_bytes, incagg_stream_t stream);
/*
 * Mean cross-entropy over the rows whose mask byte is non-zero (main.py:80
 * criterion(out[train_mask], y[train_mask])) and its gradient:
 *   out3 = {sum_i w_i CE_i, that sum / max(n, 1), n}, dlogits[i, c] = w_i / max(n,1) (softmax - onehot).
 */
size_t incagg_masked_ce_workspace_bytes(int64_t rows);
int incagg_masked_ce(const float* logits, int64_t ld, const int64_t* y, const uint8_t* mask, int64_t rows,
                     int32_t C, float* dlogits, int64_t ldd, float* out3, void* workspace,
                     size_t workspace_bytes, incagg_stream_t stream);
/*
 * The same computation as three calls, for a training step that keeps only the gradient on its critical
 * path (main.py:80-82: the loss VALUE is bookkeeping, `total_loss += loss * n`):
 *   incagg_mask_count        count[0] = n = number of non-zero mask bytes (depends on the batch alone)
 *   incagg_masked_ce_rows    dlogits as above from a count computed earlier + per-block loss partials in
 *                            the workspace (incagg_masked_ce_workspace_bytes)
 *   incagg_masked_ce_finish  out3 = {loss sum, mean, n} from the partials; acc (nullable, double[2]):
 *                            acc[0] += loss sum, acc[1] += n  (the running epoch loss, main.py:82)
 */
int incagg_mask_count(const uint8_t* mask, int64_t rows, float* count, incagg_stream_t stream);
int incagg_masked_ce_rows(const float* logits, int64_t ld, const int64_t* y, const uint8_t* mask, int64_t rows,
                          int32_t C, const float* count, float* dlogits, int64_t ldd, void* workspace,
                          size_t workspace_bytes, incagg_stream_t stream);
int incagg_masked_ce_finish(const void* workspace, int64_t rows, const float* count, float* out3, double* acc,
                            incagg_stream_t stream);

/*
 * One Adam step (torch.optim.Adam arithmetic, main.py:196-201) over flat fp32 buffers of n elements:
 * the first n_first_group elements use weight decay wd_first_group, the rest wd_rest.  step_dev is a
 * device float holding the number of steps taken (incremented by the call); arrivals_dev a zeroed
 * device uint32 scratch.  One launch.
 */
int incagg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                     int64_t n_first_group, float lr, float beta1, float beta2, float eps,
                     float wd_first_group, float wd_rest, float* step_dev, void* arrivals_dev,
                     incagg_stream_t stream);

/*
 * Data-parallel step on `world` GPUs of one NVSwitch box (SURVEY.md 8e "gradients are all-reduced"; the
 * reference itself is single-GPU): gradient all-reduce over NVLink peer memory FUSED with the Adam
 * update, one launch per rank, no NCCL call.  stage_ptrs[r] / signal_ptrs[r]: rank r's staging buffer
 * (2 * n floats) and signal array (incagg_allreduce_adam_blocks() * incagg_allreduce_adam_max_ranks()
 * int32, zero-initialised), the peers' mapped into this process through CUDA IPC.  Contributions are
 * added in rank order and divided by `world`: every rank computes bit-identical parameters.  `grads`
 * holds the averaged gradient afterwards.  All ranks must call it once per step, in step order, each on
 * its own GPU.  Other arguments as incagg_adam_step.
 */
int incagg_allreduce_adam_blocks(void);
int incagg_allreduce_adam_max_ranks(void);
int incagg_allreduce_adam_step(void* const* stage_ptrs, void* const* signal_ptrs, int rank, int world,
                               float* grads, float* params, float* exp_avg, float* exp_avg_sq, int64_t n,
                               int64_t n_first_group, float lr, float beta1, float beta2, float eps,
                               float wd_first_group, float wd_rest, float* step_dev, void* arrivals_dev,
                               incagg_stream_t stream);

/* ---- CSR transpose (CSC view for the backward SpMM) ------------------- */
/*
 * Counting-sort transpose of a [rows x cols] CSR with nnz entries.
 *   t_rowptr [cols+1], t_col [nnz] (source row of each entry), t_val [nnz] (nullable
 *   iff val is NULL), t_perm [nnz] nullable (original edge position of each entry).
 * Entries inside a transposed row are ordered by original edge position
 * (deterministic), so repeated runs reduce in the same order.
 * workspace: at least incagg_csr_transpose_workspace_bytes(rows, cols, nnz) bytes.
 */
size_t incagg_csr_transpose_workspace_bytes(int64_t rows, int64_t cols, int64_t nnz);
int incagg_csr_transpose(const int32_t* rowptr, const int32_t* col, const float* val, int64_t rows,
                         int64_t cols, int64_t nnz, int32_t* t_rowptr, int32_t* t_col, float* t_val,
                         int32_t* t_perm, void* workspace, size_t workspace_bytes,
                         incagg_stream_t stream);

/* ---- history gather / scatter / slice copies -------------------------- */
/*
 * dst[i, :] = src[idx[i], :] for i < n.  Replaces History.pull (history.py:38) and the
 * CPU index_select + bounce-buffer H2D of read_async (async_cuda.cu:95-110): `src` may be
 * device memory or pinned (page-locked, UVA-mapped) host memory, `dst` device memory.
 * Rows are `row_bytes` bytes (any size: 16/8/4/2/1-byte vectors by alignment); leading dimensions in bytes.  idx is a
 * device-accessible int64 array.  Also the collate feature gather x[n_id] (loader.py:188-190).
 */
int incagg_gather_rows(const void* src, int64_t src_ld_bytes, int64_t src_rows, const int64_t* idx,
                       int64_t n, void* dst, int64_t dst_ld_bytes, int64_t row_bytes,
                       incagg_stream_t stream);
/*
 * The same gather out of a table that is sharded by contiguous row ranges over `num_shards` <= 16
 * memories: shard s holds global rows [bounds[s], bounds[s+1]) at shard_ptrs[s] (host arrays).  Used
 * for multi-GPU history pulls: the local shard plus the peers' HBM shards mapped into this process
 * (CUDA IPC); remote rows are fetched by loads over NVLink inside the packing kernel.  Indices outside
 * [bounds[0], bounds[num_shards]) are skipped.
 */
int incagg_gather_rows_sharded(const void* const* shard_ptrs, const int64_t* bounds, int num_shards,
                               int64_t src_ld_bytes, const int64_t* idx, int64_t n, void* dst,
                               int64_t dst_ld_bytes, int64_t row_bytes, incagg_stream_t stream);
/* dst[idx[i], :] = src[i, :]  (History.push index branch, history.py:58). Out-of-range
 * indices are skipped and counted in *oob_count (device int32, nullable). */
int incagg_scatter_rows(const void* src, int64_t src_ld_bytes, const int64_t* idx, int64_t n,
                        void* dst, int64_t dst_ld_bytes, int64_t dst_rows, int64_t row_bytes,
                        incagg_stream_t stream);
/*
 * Slice copies with HOST offset/count arrays of length k (they are host tensors in the
 * reference too):
 *   direction 0 (pull, read_async async_cuda.cu:68-92):  dst[d : d+c_i] = src[offset_i : offset_i+c_i], d += c_i
 *   direction 1 (push, write_async async_cuda.cu:139-163, History.push history.py:60-65):
 *                                                        dst[offset_i : +c_i] = src[s : s+c_i], s += c_i
 * Either side may be pinned host memory (copies then run on the DMA engines).  Bounds are
 * checked against src_rows / dst_rows as the reference does ("Invalid index").
 */
int incagg_copy_slices(const void* src, int64_t src_ld_bytes, int64_t src_rows, void* dst,
                       int64_t dst_ld_bytes, int64_t dst_rows, const int64_t* offset,
                       const int64_t* count, int64_t k, int64_t row_bytes, int direction,
                       incagg_stream_t stream);

/* ---- relabel ---------------------------------------------------------- */
/*
 * GPU relabel_one_hop / relabel_one_hop_within_batch, bit-exact with
 * csrc/cpu/relabel_cpu.cpp (first-seen halo order, last duplicate of idx wins).
 *
 * The workspace is a direct-address table over global node ids (2 x int32 per node
 * plus scan scratch) that is initialised once with incagg_relabel_workspace_init and
 * left clean by every call.  It must not be shared by concurrent calls.
 *
 *   rowptr      [num_nodes+1] int64 (device)       col  [nnz] int32 or int64 (col_width 4/8)
 *   val         [nnz] fp32, nullable               idx  [B] int64 (device)
 *   nnz_b       sum of degrees of idx rows (host value; incagg_relabel_degree_sum computes it)
 *   out_rowptr  [B+1]      out_col [nnz_b]   (index width out_width 4 or 8)
 *   out_val     [nnz_b] nullable iff val NULL
 *   n_id_out    capacity B + min(nnz_b, num_nodes); receives idx followed by the halo ids
 *   counts_out  device int64[2]: {H (number of halo ids), nnz_out}
 * The non-bipartite padding of out_rowptr (relabel_cpu.cpp:98-101,208-211) is a host-side
 * concat done by the binding once H is known.
 */
size_t incagg_relabel_workspace_bytes(int64_t num_nodes);
int incagg_relabel_workspace_init(void* workspace, int64_t num_nodes, incagg_stream_t stream);
/* degsum_out: device int64[1] = sum_i (rowptr[idx[i]+1] - rowptr[idx[i]]) */
int incagg_relabel_degree_sum(const int64_t* rowptr, const int64_t* idx, int64_t B,
                              int64_t num_nodes, int64_t* degsum_out, void* workspace,
                              incagg_stream_t stream);
int incagg_relabel_one_hop(const int64_t* rowptr, const void* col, int col_width, const float* val,
                           const int64_t* idx, int64_t B, int64_t num_nodes, int64_t nnz_b,
                           void* out_rowptr, void* out_col, int out_width, float* out_val,
                           int64_t* n_id_out, int64_t* counts_out, void* workspace,
                           incagg_stream_t stream);
int incagg_relabel_one_hop_within_batch(const int64_t* rowptr, const void* col, int col_width,
                                        const float* val, const int64_t* idx, int64_t B,
                                        int64_t num_nodes, int64_t nnz_b, void* out_rowptr,
                                        void* out_col, int out_width, float* out_val,
                                        int64_t* counts_out, void* workspace,
                                        incagg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* INCAGG_B200_H */
