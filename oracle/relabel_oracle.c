/*
 * relabel_oracle.c — CPU restatement (plain C) of the reference's relabel ops.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing on the product path may link or call this file; it is the
 * checker that the CUDA kernels in incagg_gnn_b200/csrc/relabel.cu are compared against (tests/,
 * __graft_entry__.smoke(), bench.py's cpu_baseline leg).
 *
 * Follows (reference paths relative to the reference repo):
 *   oracle_relabel_one_hop               csrc/cpu/relabel_cpu.cpp:3-108
 *   oracle_relabel_one_hop_within_batch  csrc/cpu/relabel_cpu.cpp:111-214
 * Same sequential algorithm: one hash map global id -> local id, batch rows walked in order, the
 * columns of a row in CSR order, unseen columns appended in first-seen order.  The std::unordered_map
 * of the reference is an open-addressing table here (same lookups, same insert order).
 * Pinned against the reference itself: oracle/_ref/ref_relabel.so (compiled from the reference
 * sources by oracle/build_ref.sh) and the known answers of SURVEY.md §4 (tests/test_oracle.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int64_t* keys;
  int64_t* vals;
  uint64_t mask;
  int64_t size;
} map_t;

static uint64_t hash64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

static int map_init(map_t* m, int64_t expected) {
  uint64_t cap = 16;
  while (cap < (uint64_t)expected * 2 + 2) cap <<= 1;
  m->keys = (int64_t*)malloc(sizeof(int64_t) * cap);
  m->vals = (int64_t*)malloc(sizeof(int64_t) * cap);
  if (!m->keys || !m->vals) return -1;
  memset(m->keys, 0xff, sizeof(int64_t) * cap); /* -1 = empty */
  m->mask = cap - 1;
  m->size = 0;
  return 0;
}
static void map_free(map_t* m) { free(m->keys); free(m->vals); }

/* slot of key, or of the empty slot where it would be inserted */
static uint64_t map_slot(const map_t* m, int64_t key) {
  uint64_t s = hash64((uint64_t)key) & m->mask;
  while (m->keys[s] != -1 && m->keys[s] != key) s = (s + 1) & m->mask;
  return s;
}
/* n_id_map[key] = val (overwrites: the LAST duplicate of idx wins, relabel_cpu.cpp:31-36) */
static void map_set(map_t* m, int64_t key, int64_t val) {
  uint64_t s = map_slot(m, key);
  if (m->keys[s] == -1) { m->keys[s] = key; m->size++; }
  m->vals[s] = val;
}

/* sum of the degrees of the batch rows: the size of out_col / out_val */
int64_t oracle_relabel_degree_sum(const int64_t* rowptr, const int64_t* idx, int64_t B) {
  int64_t s = 0;
  for (int64_t i = 0; i < B; i++) s += rowptr[idx[i] + 1] - rowptr[idx[i]];
  return s;
}

/*
 * out_rowptr [B+1], out_col [degree_sum], out_val [degree_sum] or NULL, n_ids [capacity >=
 * min(degree_sum, num_nodes)] receives the halo ids (the caller forms n_id = cat(idx, n_ids)).
 * Returns the number of halo ids H, or -1 on allocation failure.
 * The non-bipartite rowptr padding (H copies of nnz, relabel_cpu.cpp:98-101) is the caller's concat.
 */
int64_t oracle_relabel_one_hop(const int64_t* rowptr, const int64_t* col, const float* val,
                               const int64_t* idx, int64_t B, int64_t num_nodes, int64_t* out_rowptr,
                               int64_t* out_col, float* out_val, int64_t* n_ids) {
  map_t m;
  int64_t nnz = oracle_relabel_degree_sum(rowptr, idx, B);
  /* distinct keys <= min(B + nnz, num_nodes) */
  if (map_init(&m, (B + nnz < num_nodes) ? B + nnz : num_nodes) != 0) return -1;
  int64_t offset = 0, H = 0;
  out_rowptr[0] = 0;
  for (int64_t i = 0; i < B; i++) { /* relabel_cpu.cpp:31-36 */
    int64_t v = idx[i];
    map_set(&m, v, i);
    offset += rowptr[v + 1] - rowptr[v];
    out_rowptr[i + 1] = offset;
  }
  offset = 0;
  for (int64_t i = 0; i < B; i++) { /* relabel_cpu.cpp:50-74 / 79-96 */
    int64_t v = idx[i];
    for (int64_t j = rowptr[v]; j < rowptr[v + 1]; j++) {
      int64_t w = col[j];
      uint64_t s = map_slot(&m, w);
      if (m.keys[s] == -1) { /* unseen: next halo id, first-seen order */
        int64_t c = B + H;
        m.keys[s] = w; m.vals[s] = c; m.size++;
        n_ids[H++] = w;
        out_col[offset] = c;
      } else {
        out_col[offset] = m.vals[s];
      }
      if (val) out_val[offset] = val[j];
      offset++;
    }
  }
  map_free(&m);
  return H;
}

/*
 * Keeps only edges whose column is in idx (relabel_cpu.cpp:111-214).  out_col / out_val have
 * capacity degree_sum.  Returns the number of kept edges; *distinct receives |n_id_map| (the number
 * of DISTINCT batch ids, which the reference uses for its non-bipartite padding, :208-211).
 */
int64_t oracle_relabel_one_hop_within_batch(const int64_t* rowptr, const int64_t* col,
                                            const float* val, const int64_t* idx, int64_t B,
                                            int64_t* out_rowptr, int64_t* out_col, float* out_val,
                                            int64_t* distinct) {
  map_t m;
  if (map_init(&m, B) != 0) return -1;
  for (int64_t i = 0; i < B; i++) map_set(&m, idx[i], i);
  int64_t offset = 0;
  out_rowptr[0] = 0;
  for (int64_t i = 0; i < B; i++) {
    int64_t v = idx[i];
    for (int64_t j = rowptr[v]; j < rowptr[v + 1]; j++) {
      uint64_t s = map_slot(&m, col[j]);
      if (m.keys[s] != -1) {
        out_col[offset] = m.vals[s];
        if (val) out_val[offset] = val[j];
        offset++;
      }
    }
    out_rowptr[i + 1] = offset;
  }
  if (distinct) *distinct = m.size;
  map_free(&m);
  return offset;
}

/* ---- fp32 CSR SpMM restatement (torch_sparse spmm_{sum,mean,min,max} semantics; upstream library,
 * not vendored in the reference -> parity UNPINNED, see oracle/README.md).  Sequential row order,
 * fp32 accumulation in CSR order; min/max: empty row -> 0, arg = first winning edge, -1 if empty. */
void oracle_spmm_csr(int reduce, const int64_t* rowptr, const int64_t* col, const float* val,
                     const float* X, int64_t ldx, float* out, int64_t ldo, int64_t* arg, int64_t rows,
                     int64_t F) {
  for (int64_t i = 0; i < rows; i++) {
    float* o = out + i * ldo;
    int64_t s = rowptr[i], e = rowptr[i + 1];
    for (int64_t f = 0; f < F; f++) {
      o[f] = 0.f;
      if (arg) arg[i * F + f] = -1;
    }
    for (int64_t j = s; j < e; j++) {
      const float* x = X + col[j] * ldx;
      float v = val ? val[j] : 1.f;
      for (int64_t f = 0; f < F; f++) {
        float t = v * x[f];
        if (reduce <= 1) o[f] += t;
        else if (j == s || (reduce == 2 ? t < o[f] : t > o[f])) {
          o[f] = t;
          if (arg) arg[i * F + f] = j;
        }
      }
    }
    if (reduce == 1 && e > s) {
      float d = (float)(e - s);
      for (int64_t f = 0; f < F; f++) o[f] /= d;
    }
  }
}
