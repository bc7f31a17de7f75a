/*
 * metis_shim.c — C-ABI wrapper of METIS k-way / recursive partitioning for incagg_gnn_b200.metis().
 *
 * The reference partitions with torch.ops.torch_sparse.partition (torch_geometric_autoscale/metis.py:31),
 * which calls METIS_PartGraphKway / METIS_PartGraphRecursive with default options.  torch_sparse is
 * not available here; the METIS library itself ships with the CUDA toolkit as libmetis_static.a
 * (idx_t = int64, probed), so this shim links it directly.  Host-side preprocessing, not on the
 * per-batch path.
 */
#include <stdint.h>
#include <stddef.h>

typedef int64_t idx_t;
int METIS_PartGraphKway(idx_t* nvtxs, idx_t* ncon, idx_t* xadj, idx_t* adjncy, idx_t* vwgt, idx_t* vsize,
                        idx_t* adjwgt, idx_t* nparts, float* tpwgts, float* ubvec, idx_t* options,
                        idx_t* edgecut, idx_t* part);
int METIS_PartGraphRecursive(idx_t* nvtxs, idx_t* ncon, idx_t* xadj, idx_t* adjncy, idx_t* vwgt, idx_t* vsize,
                             idx_t* adjwgt, idx_t* nparts, float* tpwgts, float* ubvec, idx_t* options,
                             idx_t* edgecut, idx_t* part);

/* rowptr [n+1], col [nnz] (int64, host); part_out [n] receives the cluster id of every node.
 * Returns the METIS status (1 = ok) and the edge cut in *edgecut. */
int incagg_metis_partition(int64_t n, int64_t* rowptr, int64_t* col, int64_t num_parts, int recursive,
                           int64_t* part_out, int64_t* edgecut) {
  idx_t nv = n, ncon = 1, np = num_parts, cut = -1;
  int rc;
  if (recursive)
    rc = METIS_PartGraphRecursive(&nv, &ncon, rowptr, col, NULL, NULL, NULL, &np, NULL, NULL, NULL, &cut, part_out);
  else
    rc = METIS_PartGraphKway(&nv, &ncon, rowptr, col, NULL, NULL, NULL, &np, NULL, NULL, NULL, &cut, part_out);
  if (edgecut) *edgecut = cut;
  return rc;
}
