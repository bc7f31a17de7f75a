"""Generates tests/golden/relabel_golden.npz from the REFERENCE's own compiled relabel op
(oracle/_ref/ref_relabel.so, built by oracle/build_ref.sh from /root/reference/csrc).  Run in the
build container (the reference is not present on the GPU box):

    python tests/golden/gen_relabel_golden.py

Cases: random CSR graphs with empty rows, duplicate edges, self loops, duplicate batch ids, halo
nodes seen from several rows, empty batch; with and without edge values; bipartite on/off.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def cases():
    rng = np.random.default_rng(20261018)
    out = []
    # the SURVEY §4 ring
    out.append(dict(rowptr=np.arange(0, 13, 2), col=np.array([1, 5, 0, 2, 1, 3, 2, 4, 3, 5, 4, 0]),
                    value=np.arange(12, dtype=np.float32), idx=np.array([1, 2])))
    for n, maxdeg, b, dup in [(50, 6, 10, False), (200, 12, 40, True), (1000, 30, 128, False),
                              (300, 4, 0, False), (64, 64, 64, False), (500, 9, 77, True)]:
        deg = rng.integers(0, maxdeg + 1, n)
        deg[rng.integers(0, n, n // 5)] = 0  # empty rows
        rowptr = np.concatenate([[0], np.cumsum(deg)])
        col = rng.integers(0, n, rowptr[-1])  # duplicates and self loops allowed
        value = rng.standard_normal(rowptr[-1]).astype(np.float32)
        idx = rng.integers(0, n, b) if dup else rng.permutation(n)[:b]
        out.append(dict(rowptr=rowptr, col=col, value=value, idx=idx))
    return out


def main():
    assert oracle.ref_available(), "build oracle/_ref first (bash oracle/build_ref.sh)"
    store = {}
    k = 0
    for c in cases():
        for fn in ("relabel_one_hop", "relabel_one_hop_within_batch"):
            for with_value in (True, False):
                for bipartite in (True, False):
                    v = c["value"] if with_value else None
                    r = oracle.ref_relabel(fn, c["rowptr"], c["col"], v, c["idx"], bipartite)
                    p = f"c{k}_"
                    store[p + "fn"] = np.array(fn)
                    store[p + "bipartite"] = np.array(bipartite)
                    store[p + "in_rowptr"] = c["rowptr"].astype(np.int64)
                    store[p + "in_col"] = c["col"].astype(np.int64)
                    store[p + "in_idx"] = c["idx"].astype(np.int64)
                    if with_value:
                        store[p + "in_value"] = c["value"]
                        store[p + "out_value"] = r[2]
                    store[p + "out_rowptr"], store[p + "out_col"], store[p + "out_n_id"] = r[0], r[1], r[3]
                    k += 1
    store["num_cases"] = np.array(k)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "relabel_golden.npz")
    np.savez_compressed(path, **store)
    print(f"wrote {k} cases to {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    main()
