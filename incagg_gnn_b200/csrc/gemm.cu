// Dense feature transforms X·W on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   D[M,N] = alpha * op(A)[M,K] · op(B)[K,N] + beta * Cin + bias[n]      (optional ReLU)
//
// Replaces the cuBLAS SGEMMs behind torch.nn.Linear / torch.addmm on the propagation path
// (reference call sites: gcn2.py:87,149 lins, GCN2Conv addmm, gcn.py:63 GCNConv.lin, graphsage.py
// lin_l / lin_r, appnp.py:79-83).  The reference computes them in fp32 (allow_tf32 = False), and
// parity is 1e-5 relative, which plain TF32 (10-bit mantissa, ~1e-3) cannot hold.  The kernel
// therefore runs the error-compensated 3xTF32 scheme: every fp32 operand is split on the way into
// shared memory into  hi = rna_tf32(x)  and  lo = rna_tf32(x - hi)  and three tcgen05.mma
// (kind::tf32, fp32 accumulation in TMEM) are issued per k-step:
//        acc += A_lo·B_hi ;  acc += A_hi·B_lo ;  acc += A_hi·B_hi
// (the dropped lo·lo term is ~2^-22 relative).
//
// Structure of one CTA (256 threads, one 128 x BN output tile, BN = 64 / 128):
//   * operand tiles are 128 (or BN) rows x 32 k-elements = one 128-byte swizzle row per operand row,
//     stored K-major in the canonical SWIZZLE_128B layout that the UMMA shared-memory descriptor
//     expects (16-byte chunk index XOR row-in-atom).  All four source layouts (A as [M,K] or [K,M],
//     B as [N,K] or [K,N], row-major) are handled by the loader, which transposes on the way in, so
//     forward (x·W^T), input-gradient (g·W) and weight-gradient (x^T·g) GEMMs are the same kernel;
//   * the global loads of tile k+1 are in flight (registers) while tile k is split, stored and
//     multiplied; the four layout combinations are separate template instances and the ragged-edge
//     loader is out of line, because the kernel body runs once per CTA and instruction fetch of a
//     large straight-line body was the first bottleneck found (ncu: 55 % of stalls "no instruction");
//   * a 3-stage ring of smem tiles: thread 0 issues the 12 MMAs of a stage and commits them to the
//     stage's mbarrier (tcgen05.commit), which frees the stage for the loader;
//   * epilogue: 8 warps, each tcgen05.ld's 32 lanes x BN/2 columns; Cin is fetched before the
//     accumulator is waited for; alpha / beta / bias / ReLU fused, vectorised stores.
//   * split-K (weight gradients: K = number of batch rows) writes fp32 partials that a second small
//     kernel sums in a fixed order (deterministic) and finishes with the same epilogue.
#include "common.cuh"

namespace incagg {

constexpr int G_BM = 128;
constexpr int G_BK = 32;     // 32 tf32 = 128 bytes = one swizzle row
constexpr int G_THREADS = 256;

enum { DUAL_NONE = 0, DUAL_K = 1, DUAL_N = 2, DUAL_M = 3, DUAL_NC = 4, DUAL_GROUP = 5 };  // NC: internal, see gemm_nc_kernel
// DUAL_GROUP: up to G_MAX_GROUP independent problems of one shape in one launch (incagg_gemm_tf32x3_group)
constexpr int G_MAX_GROUP = 16;
struct GemmGroup {
  const float* A[G_MAX_GROUP];
  const float* B[G_MAX_GROUP];
  float* D[G_MAX_GROUP];
  int64_t lda[G_MAX_GROUP], ldb[G_MAX_GROUP], ldd[G_MAX_GROUP];
  float alpha[G_MAX_GROUP];
};

struct GemmParams {
  const float* A; int64_t lda; int transA;   // transA = 0: A is [M,K] row-major; 1: stored [K,M]
  const float* B; int64_t ldb; int transB;   // transB = 0: B is [K,N] row-major; 1: stored [N,K]
  const float* Cin; int64_t ldcin;
  const float* bias;
  float* D; int64_t ldd;
  float* partial;                            // split-K partial sums [splits][Mpad][N] (NULL if 1 split)
  int64_t M, N, K;
  float alpha, beta;
  int relu;
  int acc2;                                  // DUAL_N: D2 += (instead of =): gradient accumulation in the epilogue
  int gate;                                  // single GEMM: Cin is a ReLU-backward gate, D = Cin > 0 ? alpha acc + bias : 0
  int kb_per_split;                          // k-blocks handled by one blockIdx.z
  // second operand set (fused GCNII layer GEMMs, see incagg_gemm_tf32x3_dual)
  int dual;
  const float* A2; int64_t lda2;             // DUAL_K / DUAL_M
  const float* B2; int64_t ldb2;             // DUAL_K / DUAL_N
  const float* Cin2; int64_t ldcin2; float beta2;   // DUAL_K: D += beta2 * Cin2; DUAL_N: Cin/beta of set 2
  float* D2; int64_t ldd2; float alpha2;     // DUAL_N / DUAL_M
  float scaleB, scaleB2;                     // B (B2) is multiplied by this before the hi/lo split
  int kb1;                                   // DUAL_K: k-blocks of the first segment (K = K1 + K2)
  int64_t K2;
  int tiles1;                                // DUAL_N / DUAL_M: tiles (y resp. x) of the first set
  int64_t Mpad;                              // rows of one split-K partial slab
  int64_t ldp;                               // row stride of a partial slab (N; 256 for DUAL_NC)
  int group_n;                               // DUAL_GROUP: number of problems; m-tile index / tiles1 = problem
  GemmGroup grp;
};

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] · B[smem], kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor of a K-major SWIZZLE_128B tile (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4 (= 1, unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B between 8-row groups)
//   [46,48) version = 1 (Blackwell)   [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @4, a/b_format TF32 = 2
// @7 / @10, a/b K-major = 0 @15 / @16, N >> 3 @17, M >> 4 @24.
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Round to TF32 (10 explicit mantissa bits), ties away from zero = cvt.rna.tf32.f32.  ptxas expands
// the cvt into four instructions (it special-cases Inf / NaN); on the bit pattern it is one add and
// one mask, and Inf / NaN keep their class (0x7f800000 + 0x1000 masks back to Inf).
__device__ __forceinline__ float to_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// ---- tile loader ----------------------------------------------------------------------------------
// A tile is ROWS "output" rows (m or n) x 32 k.  TRANS = the source is contiguous along the output
// dimension ([K, rows] row-major) instead of along k ([rows, K]).  Each of the 256 threads owns
// ROWS/32 float4 of a tile.  (The source layouts are template parameters and the ragged-edge loader is
// kept out of line: the kernel body is executed once per CTA, so its size in instructions matters.)
template <int ROWS>
struct TileRegs {
  float4 v[ROWS / 32];
};

__device__ __noinline__ float4 ldg4_guarded(const float* p, int valid) {
  // valid = number of in-range elements starting at p (<= 0: none); no alignment assumed
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid > 0) r.x = __ldg(p);
  if (valid > 1) r.y = __ldg(p + 1);
  if (valid > 2) r.z = __ldg(p + 2);
  if (valid > 3) r.w = __ldg(p + 3);
  return r;
}
// edge tiles: a whole in-range, aligned float4 is still one vector load; only the straddling ones
// (and unaligned sources) take the scalar path
__device__ __forceinline__ float4 ldg4_edge(const float* p, int valid, bool vec_ok) {
  if (valid >= 4 && vec_ok) return __ldg(reinterpret_cast<const float4*>(p));
  if (valid <= 0) return make_float4(0.f, 0.f, 0.f, 0.f);
  return ldg4_guarded(p, valid);
}

// Thread -> element mapping.
//   !TRANS: chunk c = tid & 7 (16 bytes of the 128-byte k-row), row r = (tid >> 3) + 32 i.  A quarter
//           warp writes the 8 chunks of one row: the XOR swizzle keeps them on distinct banks.
//   TRANS : a warp owns a 16 (k) x 8 (rows) patch: kq = lane & 3, cq = (lane >> 2) & 3 -> k = 16 kh +
//           4 cq + kq, mh = lane >> 4 -> one float4 along the rows (two lanes read one 32-byte sector
//           of a k-row); patch index ub = warp + 8 i, kh = ub & 1, row group mg = 2 (ub >> 1) + mh.
//           Element j of the float4 goes to tile row 4 mg + j at 16-byte chunk (k >> 2) ^ (row & 7) =
//           4 (kh ^ mh) + (cq ^ j): for a fixed j the 32 lanes hit 32 different banks, so the four
//           scalar stores use compile-time register indices and no conflicts.
template <int ROWS, bool TRANS>
__device__ __forceinline__ void load_tile(const float* __restrict__ src, int64_t ld, int64_t row0,
                                          int64_t rows_total, int64_t k0, int64_t k_end, bool vec_ok,
                                          TileRegs<ROWS>& t) {
  const int tid = threadIdx.x;
  // whole tile in range and 16-byte aligned: plain vector loads (uniform branch)
  const bool fast = vec_ok && (row0 + ROWS <= rows_total) && (k0 + G_BK <= k_end);
  if constexpr (!TRANS) {
    const int64_t k = k0 + (tid & 7) * 4;
    const float* base = src + (row0 + (tid >> 3)) * ld + k;
    if (fast) {
#pragma unroll
      for (int i = 0; i < ROWS / 32; ++i) t.v[i] = __ldg(reinterpret_cast<const float4*>(base + (int64_t)(32 * i) * ld));
    } else {
      const int kvalid = (int)min((int64_t)4, k_end - k);
#pragma unroll
      for (int i = 0; i < ROWS / 32; ++i) {
        const int64_t r = row0 + (tid >> 3) + 32 * i;
        t.v[i] = ldg4_edge(base + (int64_t)(32 * i) * ld, (r < rows_total) ? kvalid : 0, vec_ok);
      }
    }
  } else {
    const int lane = tid & 31, warp = tid >> 5;
    const int kl = (lane & 15), mh = lane >> 4;
    if (fast) {
#pragma unroll
      for (int i = 0; i < ROWS / 32; ++i) {
        const int ub = warp + 8 * i;
        t.v[i] = __ldg(reinterpret_cast<const float4*>(src + (k0 + (ub & 1) * 16 + kl) * ld + row0 + ((ub >> 1) * 2 + mh) * 4));
      }
    } else {
#pragma unroll
      for (int i = 0; i < ROWS / 32; ++i) {
        const int ub = warp + 8 * i;
        const int64_t k = k0 + (ub & 1) * 16 + kl;
        const int64_t r = row0 + ((ub >> 1) * 2 + mh) * 4;
        t.v[i] = ldg4_edge(src + k * ld + r, (k < k_end) ? (int)min((int64_t)4, rows_total - r) : 0, vec_ok);
      }
    }
  }
}

// Split into hi / lo TF32 parts and store into the two swizzled K-major tiles (shared-window addresses).
template <int ROWS, bool TRANS>
__device__ __forceinline__ void store_tile(const TileRegs<ROWS>& t, uint32_t hi, uint32_t lo, float scale = 1.f) {
  const int tid = threadIdx.x;
  if constexpr (!TRANS) {
    const int c = tid & 7, r0 = tid >> 3;  // r & 7 = r0 & 7 for every i
    const uint32_t off0 = (uint32_t)(r0 * 128 + ((c ^ (r0 & 7)) << 4));
#pragma unroll
    for (int i = 0; i < ROWS / 32; ++i) {
      float4 x = t.v[i];
      x.x *= scale; x.y *= scale; x.z *= scale; x.w *= scale;
      float4 h, l;
      h.x = to_tf32(x.x); h.y = to_tf32(x.y); h.z = to_tf32(x.z); h.w = to_tf32(x.w);
      l.x = to_tf32(x.x - h.x); l.y = to_tf32(x.y - h.y); l.z = to_tf32(x.z - h.z); l.w = to_tf32(x.w - h.w);
      sts128(hi + off0 + i * 32 * 128, h);
      sts128(lo + off0 + i * 32 * 128, l);
    }
  } else {
    const int lane = tid & 31, warp = tid >> 5;
    const int kq = lane & 3, cq = (lane >> 2) & 3, mh = lane >> 4;
#pragma unroll
    for (int i = 0; i < ROWS / 32; ++i) {
      const int ub = warp + 8 * i;
      const int kc = (ub & 1) * 4 + cq;                 // 16-byte chunk of k within the 128-byte row
      const int rbase = ((ub >> 1) * 2 + mh) * 4;       // rbase & 7 = 4 mh
      const float xs[4] = {t.v[i].x, t.v[i].y, t.v[i].z, t.v[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xv = xs[j] * scale;
        const float h = to_tf32(xv);
        const float l = to_tf32(xv - h);
        const uint32_t off = (uint32_t)((rbase + j) * 128 + ((kc ^ (mh * 4 + j)) << 4) + kq * 4);
        sts32(hi + off, h);
        sts32(lo + off, l);
      }
    }
  }
}

// TA: A is stored [K, M] (contiguous along m).  TBK: B is stored [K, N] (contiguous along n), i.e.
// transB == 0 of the C ABI; both make the loader transpose on the way into shared memory.
// BN = 128: three stages, one CTA per SM.  BN = 64: two stages (96 KB), two CTAs per SM - the small
// problems of this path (16 K rows) are latency-bound single waves, and a second resident CTA
// overlaps one CTA's loads / splits with the other's MMAs and epilogue.
template <int BN> struct GemmCfg { static constexpr int STAGES = (BN == 64) ? 2 : 3; static constexpr int MIN_CTAS = (BN == 64) ? 2 : 1; };

template <int BN, bool TA, bool TBK>
__global__ void __launch_bounds__(G_THREADS, GemmCfg<BN>::MIN_CTAS)
gemm_tf32x3_kernel(const __grid_constant__ GemmParams p) {
  pdl_trigger();  // the wait comes after the on-chip set-up (barriers, TMEM), before the first global access
  constexpr int G_STAGES = GemmCfg<BN>::STAGES;
  extern __shared__ __align__(1024) char smem_raw[];
  // 1024-byte alignment of every tile (the swizzle is a function of the absolute smem address)
  char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = G_BM * 128, B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  __shared__ uint64_t mma_done[G_STAGES];
  __shared__ uint64_t acc_ready;
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  // operand set of this CTA (uniform): DUAL_N switches B / D / Cin by n-tile, DUAL_M switches A / D by m-tile
  const float* Aptr = p.A; int64_t lda = p.lda; int64_t Mrows = p.M;
  const float* Bptr = p.B; int64_t ldb = p.ldb;
  const float* Cin = p.Cin; int64_t ldcin = p.ldcin;
  float* Dptr = p.D; int64_t ldd = p.ldd;
  float alpha_e = p.alpha, beta_e = p.beta, scaleB = p.scaleB;
  int bx = blockIdx.x, by = blockIdx.y;
  bool acc_d = false;                        // add the old contents of D (second output of DUAL_N with acc2)
  if (p.dual == DUAL_N && by >= p.tiles1) {
    acc_d = p.acc2 != 0;
    by -= p.tiles1; Bptr = p.B2; ldb = p.ldb2; Dptr = p.D2; ldd = p.ldd2; alpha_e = p.alpha2; beta_e = p.beta2;
    Cin = p.Cin2; ldcin = p.ldcin2; scaleB = p.scaleB2;
  }
  if (p.dual == DUAL_GROUP) {   // problem `set` of the group; D accumulates in place when beta != 0
    const int set = bx / p.tiles1;
    bx -= set * p.tiles1;
    Aptr = p.grp.A[set]; lda = p.grp.lda[set]; Bptr = p.grp.B[set]; ldb = p.grp.ldb[set];
    Dptr = p.grp.D[set]; ldd = p.grp.ldd[set]; alpha_e = p.grp.alpha[set];
    Cin = Dptr; ldcin = ldd;
  }
  if (p.dual == DUAL_M && bx >= p.tiles1) {
    bx -= p.tiles1; Aptr = p.A2; lda = p.lda2; Dptr = p.D2; ldd = p.ldd2; alpha_e = p.alpha2;
    Cin = p.Cin2; ldcin = p.ldcin2; beta_e = p.beta2;
  }
  const int64_t m0 = (int64_t)bx * G_BM, n0 = (int64_t)by * BN;
  const int64_t num_kb_total = (p.dual == DUAL_K) ? (p.kb1 + (p.K2 + G_BK - 1) / G_BK) : (p.K + G_BK - 1) / G_BK;
  const int64_t kb_lo = (int64_t)blockIdx.z * p.kb_per_split;
  const int64_t kb_hi = min(num_kb_total, kb_lo + p.kb_per_split);
  const int num_kb = (int)(kb_hi - kb_lo);

  if (tid == 0) {
    for (int s = 0; s < G_STAGES; ++s) mbar_init(&mma_done[s], 1);
    mbar_init(&acc_ready, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_slot, BN);  // BN fp32 accumulator columns x 128 lanes

  const bool vecA = (lda % 4 == 0) && aligned16(Aptr) &&
                    (p.dual != DUAL_K || ((p.lda2 % 4 == 0) && aligned16(p.A2)));
  const bool vecB = (ldb % 4 == 0) && aligned16(Bptr) &&
                    (p.dual != DUAL_K || ((p.ldb2 % 4 == 0) && aligned16(p.B2)));
  constexpr uint32_t IDESC = umma_idesc(G_BM, BN);

  // k-block -> operand segment (DUAL_K: the first kb1 blocks read (A, B), the rest (A2, B2))
  auto load_kb = [&](int64_t kbg, TileRegs<G_BM>& a, TileRegs<BN>& b) {
    if (p.dual == DUAL_K && kbg >= p.kb1) {
      load_tile<G_BM, TA>(p.A2, p.lda2, m0, Mrows, (kbg - p.kb1) * G_BK, p.K2, vecA, a);
      load_tile<BN, TBK>(p.B2, p.ldb2, n0, p.N, (kbg - p.kb1) * G_BK, p.K2, vecB, b);
    } else {
      load_tile<G_BM, TA>(Aptr, lda, m0, Mrows, kbg * G_BK, p.K, vecA, a);
      load_tile<BN, TBK>(Bptr, ldb, n0, p.N, kbg * G_BK, p.K, vecB, b);
    }
  };
  // the global loads of k-block kb+1 are in flight while kb is split, stored and multiplied
  TileRegs<G_BM> ra;
  TileRegs<BN> rb;
  pdl_wait();
  if (num_kb > 0) load_kb(kb_lo, ra, rb);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_slot;
  const uint32_t smem_base = smem_u32(smem);

#pragma unroll 1
  for (int kb = 0; kb < num_kb; ++kb) {
    const int s = kb % G_STAGES;
    const uint32_t st = smem_base + (uint32_t)(s * STAGE_BYTES);
    if (kb >= G_STAGES) mbar_wait(&mma_done[s], (uint32_t)(((kb / G_STAGES) - 1) & 1));  // stage free?
    store_tile<G_BM, TA>(ra, st, st + A_BYTES);
    store_tile<BN, TBK>(rb, st + 2 * A_BYTES, st + 2 * A_BYTES + B_BYTES,
                        (p.dual == DUAL_K && kb_lo + kb >= p.kb1) ? p.scaleB2 : scaleB);
    if (kb + 1 < num_kb) load_kb(kb_lo + kb + 1, ra, rb);
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      if (lane == 0) {
        tc_fence_after();
        const uint32_t a_hi = st, a_lo = a_hi + A_BYTES;
        const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
        const uint64_t d_ahi = umma_desc(a_hi), d_alo = umma_desc(a_lo);
        const uint64_t d_bhi = umma_desc(b_hi), d_blo = umma_desc(b_lo);
#pragma unroll
        for (int k = 0; k < G_BK / 8; ++k) {  // UMMA_K = 8 tf32 = 32 bytes: start address += 32 >> 4
          const uint64_t ko = (uint64_t)(k * 2);
          umma_tf32(tmem_acc, d_alo + ko, d_bhi + ko, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
          umma_tf32(tmem_acc, d_ahi + ko, d_blo + ko, IDESC, 1u);
          umma_tf32(tmem_acc, d_ahi + ko, d_bhi + ko, IDESC, 1u);
        }
        umma_commit(&mma_done[s]);  // arrives when the MMAs that read this stage have finished
        if (kb == num_kb - 1) umma_commit(&acc_ready);
      }
      __syncwarp();
    }
  }

  // ---- epilogue: warp w reads TMEM lanes 32 (w & 3) .. +31 (its rows), columns half (w >> 2) ----
  constexpr int CW = BN / 2;             // columns per warp
  const int64_t m = m0 + (warp & 3) * 32 + lane;
  const int cbase = (warp >> 2) * CW;
  const bool splitk = p.partial != nullptr;
  // split-K partial slab row: the global m-tile index (blockIdx.x), so DUAL_M sets do not collide
  float* drow = splitk ? p.partial + ((int64_t)blockIdx.z * p.Mpad + (int64_t)blockIdx.x * G_BM + (warp & 3) * 32 + lane) * p.ldp
                       : Dptr + m * ldd;
  const int64_t ldd_eff = splitk ? p.ldp : ldd;
  const bool vecD = (ldd_eff % 4 == 0) && aligned16(splitk ? (const void*)p.partial : (const void*)Dptr);
  const bool gate = p.gate != 0;             // (DUAL_NONE only: Cin is this CTA's gate)
  if (gate) beta_e = 1.f;                    // the gate values are kept as they are
  const bool use_cin = !splitk && Cin && beta_e != 0.f;
  const bool vecC = use_cin && (ldcin % 4 == 0) && aligned16(Cin);
  const bool full_n = (n0 + BN <= p.N);
  // Cin of this thread's row segment is fetched before waiting for the accumulator
  // (pre-scaled by beta; DUAL_K adds beta2 * Cin2)
  float4 cin[CW / 4];
#pragma unroll
  for (int i = 0; i < CW / 4; ++i) cin[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (use_cin && m < Mrows) {
    const float* cp = Cin + m * ldcin + n0 + cbase;
#pragma unroll
    for (int i = 0; i < CW / 4; ++i) {
      const float4 t = (vecC && full_n) ? __ldg(reinterpret_cast<const float4*>(cp + i * 4))
                                        : ldg4_guarded(cp + i * 4, (int)min((int64_t)4, p.N - (n0 + cbase + i * 4)));
      cin[i] = make_float4(beta_e * t.x, beta_e * t.y, beta_e * t.z, beta_e * t.w);
    }
  }
  if (!splitk && acc_d && m < Mrows) {        // D2 += : its old row segment joins the Cin terms
    const float* cp = Dptr + m * ldd + n0 + cbase;
    const bool v2 = vecD && full_n;
#pragma unroll
    for (int i = 0; i < CW / 4; ++i) {
      const float4 t = v2 ? *reinterpret_cast<const float4*>(cp + i * 4)
                          : ldg4_guarded(cp + i * 4, (int)min((int64_t)4, p.N - (n0 + cbase + i * 4)));
      cin[i].x += t.x; cin[i].y += t.y; cin[i].z += t.z; cin[i].w += t.w;
    }
  }
  if (!splitk && p.dual == DUAL_K && p.Cin2 && p.beta2 != 0.f && m < Mrows) {
    const float* cp = p.Cin2 + m * p.ldcin2 + n0 + cbase;
    const bool v2 = (p.ldcin2 % 4 == 0) && aligned16(p.Cin2) && full_n;
#pragma unroll
    for (int i = 0; i < CW / 4; ++i) {
      const float4 t = v2 ? __ldg(reinterpret_cast<const float4*>(cp + i * 4))
                          : ldg4_guarded(cp + i * 4, (int)min((int64_t)4, p.N - (n0 + cbase + i * 4)));
      cin[i].x += p.beta2 * t.x; cin[i].y += p.beta2 * t.y; cin[i].z += p.beta2 * t.z; cin[i].w += p.beta2 * t.w;
    }
  }
  if (num_kb > 0) mbar_wait(&acc_ready, 0);
  tc_fence_after();
  const float alpha = alpha_e;
  const bool do_relu = !splitk && p.relu;
  const float* bias = splitk ? nullptr : p.bias;
#pragma unroll
  for (int c0 = 0; c0 < CW; c0 += 32) {
    float v[32];
    if (num_kb > 0) {
      tmem_ld32(tmem_acc + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cbase + c0), v);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;
    }
    if (m < Mrows) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const int64_t n = n0 + cbase + c0 + i;
        const float4 ci = cin[(c0 + i) / 4];
        float o[4];
        if (splitk) {
          o[0] = v[i]; o[1] = v[i + 1]; o[2] = v[i + 2]; o[3] = v[i + 3];
        } else if (gate) {   // ReLU backward: the gradient passes where the forward output was positive
          o[0] = ci.x > 0.f ? alpha * v[i] : 0.f;
          o[1] = ci.y > 0.f ? alpha * v[i + 1] : 0.f;
          o[2] = ci.z > 0.f ? alpha * v[i + 2] : 0.f;
          o[3] = ci.w > 0.f ? alpha * v[i + 3] : 0.f;
        } else {
          o[0] = alpha * v[i] + ci.x;
          o[1] = alpha * v[i + 1] + ci.y;
          o[2] = alpha * v[i + 2] + ci.z;
          o[3] = alpha * v[i + 3] + ci.w;
        }
        if (full_n && vecD) {
          if (bias) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n));
            o[0] += bv.x; o[1] += bv.y; o[2] += bv.z; o[3] += bv.w;
          }
          if (do_relu) {
            o[0] = fmaxf(o[0], 0.f); o[1] = fmaxf(o[1], 0.f); o[2] = fmaxf(o[2], 0.f); o[3] = fmaxf(o[3], 0.f);
          }
          *reinterpret_cast<float4*>(drow + n) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
          for (int j = 0; j < 4; ++j)
            if (n + j < p.N) {
              float x = o[j];
              if (bias) x += __ldg(bias + n + j);
              if (do_relu) x = fmaxf(x, 0.f);
              drow[n + j] = x;
            }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_acc, BN);
}


// ---- N-concatenated pair: D = alpha op(A) (sB B) + beta Cin,  D2 = alpha2 op(A) (sB2 B2) + beta2 Cin2 ----
// One CTA owns a 128-row tile of the shared operand A and BOTH outputs: the B tile is 256 rows
// (rows 0..127 from B, 128..255 from B2; N <= 128 each), the accumulator 128 x 256 in TMEM, the MMA
// shape 128 x 256 x 8.  A is split into hi / lo once instead of once per output, and the grid has
// half the CTAs (the 16 K-row problems fit one wave).
template <bool TA, bool TBK>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_nc_kernel(const __grid_constant__ GemmParams p) {
  pdl_trigger();  // the wait comes after the on-chip set-up (barriers, TMEM), before the first global access
  constexpr int STAGES = 2, BN = 256, HALF = 128;
  extern __shared__ __align__(1024) char smem_raw[];
  char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = G_BM * 128, B_BYTES = BN * 128, BH_BYTES = HALF * 128;
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  __shared__ uint64_t mma_done[STAGES];
  __shared__ uint64_t acc_ready;
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int64_t m0 = (int64_t)blockIdx.x * G_BM;
  const int64_t num_kb_total = (p.K + G_BK - 1) / G_BK;
  const int64_t kb_lo = (int64_t)blockIdx.z * p.kb_per_split;
  const int64_t kb_hi = min(num_kb_total, kb_lo + p.kb_per_split);
  const int num_kb = (int)(kb_hi - kb_lo);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&mma_done[s], 1);
    mbar_init(&acc_ready, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_slot, BN);

  const bool vecA = (p.lda % 4 == 0) && aligned16(p.A);
  const bool vecB = (p.ldb % 4 == 0) && aligned16(p.B), vecB2 = (p.ldb2 % 4 == 0) && aligned16(p.B2);
  constexpr uint32_t IDESC = umma_idesc(G_BM, BN);

  TileRegs<G_BM> ra;
  TileRegs<HALF> rb, rb2;
  auto load_kb = [&](int64_t kbg) {
    load_tile<G_BM, TA>(p.A, p.lda, m0, p.M, kbg * G_BK, p.K, vecA, ra);
    load_tile<HALF, TBK>(p.B, p.ldb, 0, p.N, kbg * G_BK, p.K, vecB, rb);
    load_tile<HALF, TBK>(p.B2, p.ldb2, 0, p.N, kbg * G_BK, p.K, vecB2, rb2);
  };
  pdl_wait();
  if (num_kb > 0) load_kb(kb_lo);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_slot;
  const uint32_t smem_base = smem_u32(smem);

#pragma unroll 1
  for (int kb = 0; kb < num_kb; ++kb) {
    const int s = kb % STAGES;
    const uint32_t st = smem_base + (uint32_t)(s * STAGE_BYTES);
    if (kb >= STAGES) mbar_wait(&mma_done[s], (uint32_t)(((kb / STAGES) - 1) & 1));
    const uint32_t b_hi = st + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
    store_tile<G_BM, TA>(ra, st, st + A_BYTES);
    store_tile<HALF, TBK>(rb, b_hi, b_lo, p.scaleB);
    store_tile<HALF, TBK>(rb2, b_hi + BH_BYTES, b_lo + BH_BYTES, p.scaleB2);
    if (kb + 1 < num_kb) load_kb(kb_lo + kb + 1);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      if (lane == 0) {
        tc_fence_after();
        const uint64_t d_ahi = umma_desc(st), d_alo = umma_desc(st + A_BYTES);
        const uint64_t d_bhi = umma_desc(b_hi), d_blo = umma_desc(b_lo);
#pragma unroll
        for (int k = 0; k < G_BK / 8; ++k) {
          const uint64_t ko = (uint64_t)(k * 2);
          umma_tf32(tmem_acc, d_alo + ko, d_bhi + ko, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
          umma_tf32(tmem_acc, d_ahi + ko, d_blo + ko, IDESC, 1u);
          umma_tf32(tmem_acc, d_ahi + ko, d_bhi + ko, IDESC, 1u);
        }
        umma_commit(&mma_done[s]);
        if (kb == num_kb - 1) umma_commit(&acc_ready);
      }
      __syncwarp();
    }
  }

  // ---- epilogue: warps 0-3 write output 1 (TMEM columns 0..127), warps 4-7 output 2 (128..255);
  //      warp w reads TMEM lanes 32 (w & 3) .. +31 = its rows ----
  const int set = warp >> 2;
  const int64_t m = m0 + (warp & 3) * 32 + lane;
  const bool splitk = p.partial != nullptr;
  const float* Cin = set ? p.Cin2 : p.Cin;
  const int64_t ldcin = set ? p.ldcin2 : p.ldcin;
  float* Dptr = set ? p.D2 : p.D;
  const int64_t ldd = set ? p.ldd2 : p.ldd;
  const float alpha = set ? p.alpha2 : p.alpha, beta = set ? p.beta2 : p.beta;
  const bool use_cin = !splitk && Cin && beta != 0.f;
  const bool acc_d = set && p.acc2;          // D2 += (the split-K reduce kernel does it otherwise)
  const bool full_n = (p.N == HALF);
  const bool vecD = full_n && (splitk || ((ldd % 4 == 0) && aligned16(Dptr)));
  const bool vecC = use_cin && full_n && (ldcin % 4 == 0) && aligned16(Cin);
  float* prow = splitk ? p.partial + ((int64_t)blockIdx.z * p.Mpad + m) * p.ldp + set * HALF : nullptr;
  // D2 += : the old values of the next 32 columns are fetched one iteration ahead (the first batch before
  // the accumulator is waited for).  Read in place between the stores they would be 32 dependent round
  // trips per thread: a load may not pass an earlier store to the same buffer, and each is consumed at once.
  // Cin likewise: one batch of eight 16-byte loads in flight ahead of its use instead of one L2 round trip
  // per 32 columns after the accumulator has arrived.
  const bool acc_vec = acc_d && !splitk && vecD && m < p.M;
  const bool cin_vec = use_cin && vecC && vecD && m < p.M;
  float4 old[8], cnx[8];
  auto load_old = [&](int c0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) old[i] = *reinterpret_cast<const float4*>(Dptr + m * ldd + c0 + 4 * i);
  };
  auto load_cin = [&](int c0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) cnx[i] = __ldg(reinterpret_cast<const float4*>(Cin + m * ldcin + c0 + 4 * i));
  };
  if (cin_vec) load_cin(0);
  if (acc_vec) load_old(0);
  if (num_kb > 0) mbar_wait(&acc_ready, 0);
  tc_fence_after();
#pragma unroll 1
  for (int c0 = 0; c0 < HALF; c0 += 32) {
    float v[32];
    if (num_kb > 0) {
      tmem_ld32(tmem_acc + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(set * HALF + c0), v);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;
    }
    if (m >= p.M) continue;
    if (splitk) {  // raw partial sums, slab row m, columns set * 128 + n
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(prow + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else if (vecD) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 o = make_float4(alpha * v[i], alpha * v[i + 1], alpha * v[i + 2], alpha * v[i + 3]);
        if (use_cin) {
          const float* cp = Cin + m * ldcin + c0 + i;
          const float4 t = vecC ? cnx[i / 4] : make_float4(cp[0], cp[1], cp[2], cp[3]);
          o.x += beta * t.x; o.y += beta * t.y; o.z += beta * t.z; o.w += beta * t.w;
        }
        if (acc_d) {
          const float4 t = old[i / 4];
          o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
        }
        if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4*>(Dptr + m * ldd + c0 + i) = o;
      }
      if (cin_vec && c0 + 32 < HALF) load_cin(c0 + 32);
      if (acc_vec && c0 + 32 < HALF) load_old(c0 + 32);
    } else {
      for (int i = 0; i < 32; ++i) {
        const int64_t n = c0 + i;
        if (n >= p.N) break;
        float x = alpha * v[i];
        if (use_cin) x += beta * Cin[m * ldcin + n];
        if (acc_d) x += Dptr[m * ldd + n];
        if (p.relu) x = fmaxf(x, 0.f);
        Dptr[m * ldd + n] = x;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_acc, BN);
}

// D = alpha * sum_z partial[z] + beta * Cin + bias (+ReLU).  A block of 8 warps owns 32 consecutive
// outputs of one row segment: warp w sums the partials z = w, w + 8, ... (independent coalesced loads),
// the eight sums are combined through shared memory in warp order - a fixed association, so the result
// is deterministic.  Partial slabs are [Mpad][N] with Mpad = gridDim.x * 128 rows; for DUAL_M the rows
// of the second output start at tiles1 * 128.
constexpr int RED_WARPS = 8;
__global__ void __launch_bounds__(RED_WARPS * 32)
gemm_splitk_reduce_kernel(const __grid_constant__ GemmParams p, int splits, int n_chunks) {
  pdl_prologue();
  __shared__ float part[RED_WARPS][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t sets = (p.dual == DUAL_M || p.dual == DUAL_NC) ? 2 : (p.dual == DUAL_GROUP ? p.group_n : 1);
  const int64_t items = sets * p.M * n_chunks;             // one item = 32 columns of one output row
  const int64_t slab = p.Mpad * p.ldp;
  for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
    const int64_t set = it / (p.M * n_chunks);
    const int64_t j = it - set * p.M * n_chunks;
    const int64_t m = j / n_chunks;
    const int64_t n = (j - m * n_chunks) * 32 + lane;
    // DUAL_M: the second output's rows follow the first's; DUAL_NC: its columns start at 128
    const int64_t prow = m + ((p.dual == DUAL_M || p.dual == DUAL_GROUP) ? set * (int64_t)p.tiles1 * G_BM : 0);
    float acc = 0.f;
    if (n < p.N) {
      const float* src = p.partial + prow * p.ldp + n + ((set && p.dual == DUAL_NC) ? 128 : 0);
      int z = w;
      for (; z + 3 * RED_WARPS < splits; z += 4 * RED_WARPS) {
        const float a0 = src[(int64_t)z * slab], a1 = src[(int64_t)(z + RED_WARPS) * slab];
        const float a2 = src[(int64_t)(z + 2 * RED_WARPS) * slab], a3 = src[(int64_t)(z + 3 * RED_WARPS) * slab];
        acc += a0; acc += a1; acc += a2; acc += a3;
      }
      for (; z < splits; z += RED_WARPS) acc += src[(int64_t)z * slab];
    }
    part[w][lane] = acc;
    __syncthreads();
    if (w == 0 && n < p.N) {
      float t = part[0][lane];
#pragma unroll
      for (int i = 1; i < RED_WARPS; ++i) t += part[i][lane];
      if (p.dual == DUAL_GROUP) {
        float* d = p.grp.D[set] + m * p.grp.ldd[set] + n;
        float x = p.grp.alpha[set] * t;
        if (p.beta != 0.f) x += p.beta * *d;
        *d = x;
      } else if (set == 0) {
        float x = p.alpha * t;
        if (p.gate) x = p.Cin[m * p.ldcin + n] > 0.f ? x : 0.f;
        else if (p.Cin && p.beta != 0.f) x += p.beta * p.Cin[m * p.ldcin + n];
        if (p.bias) x += p.bias[n];
        if (p.relu) x = fmaxf(x, 0.f);
        p.D[m * p.ldd + n] = x;
      } else {
        float x = p.alpha2 * t;
        if (p.Cin2 && p.beta2 != 0.f) x += p.beta2 * p.Cin2[m * p.ldcin2 + n];
        if (p.acc2 && p.dual == DUAL_NC) x += p.D2[m * p.ldd2 + n];
        if (p.relu && p.dual == DUAL_NC) x = fmaxf(x, 0.f);
        p.D2[m * p.ldd2 + n] = x;
      }
    }
    __syncthreads();
  }
}

template <int BN, bool TA, bool TBK>
static int launch_gemm_t(const GemmParams& p, int splits, cudaStream_t st) {
  constexpr int STAGE_BYTES = 2 * G_BM * 128 + 2 * BN * 128;
  constexpr int SMEM = GemmCfg<BN>::STAGES * STAGE_BYTES + 1024;
  static bool attr_set[16] = {false};
  if (first_use_on_device(attr_set)) {
    IA_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<BN, TA, TBK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  }
  unsigned gx = (unsigned)((p.M + G_BM - 1) / G_BM), gy = (unsigned)((p.N + BN - 1) / BN);
  if (p.dual == DUAL_M) gx *= 2;
  if (p.dual == DUAL_GROUP) gx *= (unsigned)p.group_n;
  if (p.dual == DUAL_N) gy *= 2;
  dim3 grid(gx, gy, (unsigned)splits);
  launch(gemm_tf32x3_kernel<BN, TA, TBK>, dim3(grid), dim3(G_THREADS), (size_t)(SMEM), st, p);
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}

template <bool TA, bool TBK>
static int launch_nc_t(const GemmParams& p, int splits, cudaStream_t st) {
  constexpr int SMEM = 2 * (2 * G_BM * 128 + 2 * 256 * 128) + 1024;
  static bool attr_set[16] = {false};
  if (first_use_on_device(attr_set)) {
    IA_CUDA(cudaFuncSetAttribute(gemm_nc_kernel<TA, TBK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  }
  dim3 grid((unsigned)((p.M + G_BM - 1) / G_BM), 1, (unsigned)splits);
  launch(gemm_nc_kernel<TA, TBK>, dim3(grid), dim3(G_THREADS), (size_t)(SMEM), st, p);
  IA_LAUNCH_CHECK();
  return INCAGG_OK;
}
static int launch_nc(const GemmParams& p, int splits, cudaStream_t st) {
  const bool ta = p.transA != 0, tbk = p.transB == 0;
  if (ta) return tbk ? launch_nc_t<true, true>(p, splits, st) : launch_nc_t<true, false>(p, splits, st);
  return tbk ? launch_nc_t<false, true>(p, splits, st) : launch_nc_t<false, false>(p, splits, st);
}

template <int BN>
static int launch_gemm(const GemmParams& p, int splits, cudaStream_t st) {
  const bool ta = p.transA != 0, tbk = p.transB == 0;
  if (ta) return tbk ? launch_gemm_t<BN, true, true>(p, splits, st) : launch_gemm_t<BN, true, false>(p, splits, st);
  return tbk ? launch_gemm_t<BN, false, true>(p, splits, st) : launch_gemm_t<BN, false, false>(p, splits, st);
}

static int run_gemm(GemmParams& p, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  static const bool no_nc = getenv("INCAGG_GEMM_NO_NC") != nullptr;   // A/B switch
  // N-concatenated pairs with N <= 128 run in the pair kernel (one CTA owns both outputs).  (Running
  // the M-concatenated weight gradients there as the transposed problem measured slower: 26 vs 24 us.)
  if (!no_nc && p.dual == DUAL_N && p.N <= 128 && p.bias == nullptr) {
    p.dual = DUAL_NC;
  }
  const bool nc = p.dual == DUAL_NC;
  static const int force_bn = getenv("INCAGG_GEMM_BN") ? atoi(getenv("INCAGG_GEMM_BN")) : 0;
  // 64-wide n-tiles only for narrow outputs: on the 16 K-row problems of this path, splitting 128
  // columns over two co-resident CTAs measured slower (the A tile is split into hi / lo twice)
  const int64_t mt = (p.M + G_BM - 1) / G_BM;
  const bool many_tiles = p.dual == DUAL_NONE && mt >= (int64_t)tune_get(INCAGG_TUNE_GEMM_BN64_MIN_TILES, 1 << 30);
  const int bn = nc ? 256 : ((p.N <= 64 || force_bn == 64 || many_tiles) ? 64 : 128);
  const int64_t nt = nc ? 1 : (p.N + bn - 1) / bn;
  p.tiles1 = (int)(p.dual == DUAL_N ? nt : mt);
  const int64_t gx = mt * (p.dual == DUAL_M ? 2 : (p.dual == DUAL_GROUP ? p.group_n : 1)), gy = nt * (p.dual == DUAL_N ? 2 : 1);
  p.Mpad = gx * G_BM;
  p.ldp = nc ? 256 : p.N;
  if (p.dual == DUAL_K) p.kb1 = (int)((p.K + G_BK - 1) / G_BK);
  const int64_t num_kb = (p.K + G_BK - 1) / G_BK + (p.dual == DUAL_K ? (p.K2 + G_BK - 1) / G_BK : 0);
  // split-K when the output has few tiles and the reduction is long (weight gradients)
  const int64_t tiles = gx * gy;
  int splits = 1;
  if (num_kb >= 16 && tiles < sm_count() && p.dual != DUAL_K && p.dual != DUAL_N && workspace != nullptr) {
    // one wave: (tiles x splits) CTAs <= SM count (one CTA per SM is resident)
    const int64_t budget = p.dual == DUAL_M ? (int64_t)tune_get(INCAGG_TUNE_GEMM_DUAL_M_CTAS, sm_count()) : (int64_t)sm_count();
    int64_t want = budget / tiles;
    if (want > num_kb / 4) want = num_kb / 4;
    if (want > 128) want = 128;
    const int64_t fit = (int64_t)(workspace_bytes / (sizeof(float) * (size_t)p.Mpad * (size_t)p.ldp));
    if (want > fit) want = fit;
    if (want > 1) splits = (int)want;
  }
  p.kb_per_split = (int)((num_kb + splits - 1) / splits);
  if (p.kb_per_split < 1) p.kb_per_split = 1;
  splits = (int)((num_kb + p.kb_per_split - 1) / p.kb_per_split);
  if (splits < 1) splits = 1;
  p.partial = splits > 1 ? static_cast<float*>(workspace) : nullptr;
  int rc = nc ? launch_nc(p, splits, st) : ((bn == 64) ? launch_gemm<64>(p, splits, st) : launch_gemm<128>(p, splits, st));
  if (rc != INCAGG_OK) return rc;
  if (splits > 1) {
    const int n_chunks = (int)((p.N + 31) / 32);
    const int64_t items = p.M * n_chunks * ((p.dual == DUAL_M || nc) ? 2 : (p.dual == DUAL_GROUP ? p.group_n : 1));
    const int blocks = (int)(items < (int64_t)sm_count() * 8 ? items : (int64_t)sm_count() * 8);
    launch(gemm_splitk_reduce_kernel, dim3(blocks), dim3(RED_WARPS * 32), (size_t)(0), st, p, splits, n_chunks);
    IA_LAUNCH_CHECK();
  }
  return INCAGG_OK;
}

}  // namespace incagg

using namespace incagg;

extern "C" size_t incagg_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  // split-K partials (rows padded to the 128-row tile, two output sets at most, <= 128 splits)
  const size_t mpad = (size_t)((M + G_BM - 1) / G_BM) * G_BM * 2;
  return sizeof(float) * mpad * (size_t)N * 128;
}

extern "C" int incagg_gemm_tf32x3(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A,
                                  int64_t lda, const float* B, int64_t ldb, float alpha, const float* Cin,
                                  int64_t ldcin, float beta, const float* bias, int relu, float* D,
                                  int64_t ldd, void* workspace, size_t workspace_bytes,
                                  incagg_stream_t stream) {
  IA_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "negative size");
  if (M == 0 || N == 0) return INCAGG_OK;
  IA_CHECK_ARG(D != nullptr, "D is NULL");
  IA_CHECK_ARG(K == 0 || (A != nullptr && B != nullptr), "NULL operand");
  IA_CHECK_ARG(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldd >= N, "leading dimension too small");
  IA_CHECK_ARG(Cin == nullptr || ldcin >= N, "ldcin too small");
  IA_CHECK_ARG((M + G_BM - 1) / G_BM <= 0x3fffffff, "M too large");
  GemmParams p{};
  p.A = A; p.lda = lda; p.transA = transA; p.B = B; p.ldb = ldb; p.transB = transB;
  p.Cin = Cin; p.ldcin = ldcin; p.bias = bias; p.D = D; p.ldd = ldd; p.M = M; p.N = N; p.K = K;
  IA_CHECK_ARG((relu & ~5) == 0, "flags: bit 0 ReLU, bit 2 Cin is a ReLU-backward gate");
  IA_CHECK_ARG(!(relu & 4) || (Cin != nullptr && bias == nullptr && !(relu & 1)), "gate needs Cin, no bias, no ReLU");
  p.alpha = alpha; p.beta = beta; p.relu = relu & 1; p.gate = (relu >> 2) & 1; p.scaleB = 1.f; p.scaleB2 = 1.f;
  p.dual = DUAL_NONE;
  return run_gemm(p, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int incagg_gemm_tf32x3_dual(int mode, int transA, int transB, int64_t M, int64_t N, int64_t K,
                                       int64_t K2, const float* A, int64_t lda, const float* A2, int64_t lda2,
                                       const float* B, int64_t ldb, const float* B2, int64_t ldb2,
                                       float alpha, float alpha2, float scaleB, float scaleB2,
                                       const float* Cin, int64_t ldcin, float beta, const float* Cin2,
                                       int64_t ldcin2, float beta2, int relu, float* D, int64_t ldd,
                                       float* D2, int64_t ldd2, void* workspace, size_t workspace_bytes,
                                       incagg_stream_t stream) {
  IA_CHECK_ARG(mode >= DUAL_K && mode <= DUAL_M, "mode must be 1 (K), 2 (N) or 3 (M)");
  IA_CHECK_ARG(M > 0 && N > 0 && K > 0, "empty problem");
  IA_CHECK_ARG(A && B && D, "NULL operand");
  IA_CHECK_ARG(mode != DUAL_K || (A2 && B2 && K2 > 0), "K-concatenation needs A2, B2, K2");
  IA_CHECK_ARG(mode != DUAL_N || (B2 && D2), "N-concatenation needs B2, D2");
  IA_CHECK_ARG(mode != DUAL_M || (A2 && D2), "M-concatenation needs A2, D2");
  GemmParams p{};
  p.A = A; p.lda = lda; p.transA = transA; p.B = B; p.ldb = ldb; p.transB = transB;
  p.A2 = A2; p.lda2 = lda2; p.B2 = B2; p.ldb2 = ldb2;
  p.Cin = Cin; p.ldcin = ldcin; p.beta = beta; p.Cin2 = Cin2; p.ldcin2 = ldcin2; p.beta2 = beta2;
  p.D = D; p.ldd = ldd; p.D2 = D2; p.ldd2 = ldd2; p.M = M; p.N = N; p.K = K; p.K2 = K2;
  IA_CHECK_ARG((relu & ~3) == 0 && (!(relu & 2) || mode == DUAL_N), "flags: bit 0 ReLU, bit 1 (mode 2 only) D2 accumulates");
  p.alpha = alpha; p.alpha2 = alpha2; p.scaleB = scaleB; p.scaleB2 = scaleB2; p.relu = relu & 1;
  p.acc2 = (relu >> 1) & 1; p.dual = mode;
  return run_gemm(p, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int incagg_gemm_tf32x3_group(int count, int transA, int transB, int64_t M, int64_t N, int64_t K,
                                        const float* const* A, const int64_t* lda, const float* const* B,
                                        const int64_t* ldb, const float* alpha, float beta, float* const* D,
                                        const int64_t* ldd, void* workspace, size_t workspace_bytes,
                                        incagg_stream_t stream) {
  IA_CHECK_ARG(count >= 1 && count <= G_MAX_GROUP, "1 <= count <= 16");
  IA_CHECK_ARG(M > 0 && N > 0 && K > 0, "empty problem");
  IA_CHECK_ARG(A && B && D && lda && ldb && ldd && alpha, "NULL argument");
  IA_CHECK_ARG(((M + G_BM - 1) / G_BM) * count <= 0x3fffffff, "too many tiles");
  GemmParams p{};
  for (int g = 0; g < count; ++g) {
    IA_CHECK_ARG(A[g] && B[g] && D[g], "NULL operand");
    IA_CHECK_ARG(lda[g] >= (transA ? M : K) && ldb[g] >= (transB ? K : N) && ldd[g] >= N, "leading dimension too small");
    p.grp.A[g] = A[g]; p.grp.B[g] = B[g]; p.grp.D[g] = D[g];
    p.grp.lda[g] = lda[g]; p.grp.ldb[g] = ldb[g]; p.grp.ldd[g] = ldd[g]; p.grp.alpha[g] = alpha[g];
  }
  p.group_n = count;
  // (set 0 stands in for the scalar fields; the kernels read the group table)
  p.A = A[0]; p.lda = lda[0]; p.transA = transA; p.B = B[0]; p.ldb = ldb[0]; p.transB = transB;
  p.D = D[0]; p.ldd = ldd[0]; p.Cin = beta != 0.f ? D[0] : nullptr; p.ldcin = ldd[0];
  p.M = M; p.N = N; p.K = K; p.alpha = alpha[0]; p.beta = beta; p.scaleB = 1.f; p.scaleB2 = 1.f;
  p.dual = DUAL_GROUP;
  return run_gemm(p, workspace, workspace_bytes, as_stream(stream));
}
