"""Count ingest: the reference's `format_counts_*` (src/utils.jl:1-36) restated for
flat CSR buffers, which is what the C-ABI (include/mmsig.h) takes.

Reference layout: X[d][m] is an nnz x 2 Int matrix, column 1 = 1-based term index, column 2 =
count > 0 (src/utils.jl:1-7).  Here, per modality: rowptr int64[D+1], term int32[nnz]
(0-BASED), count int32[nnz]; rows keep the reference's order (ascending term).
"""
import numpy as np


def make_count_csr(dense):
    """dense: (V, D) integer matrix, one column per sample (the TSV layout).  Returns
    (rowptr, term, count) with zeros dropped, as make_count_matrix does (src/utils.jl:1-7)."""
    dense = np.asarray(dense)
    V, D = dense.shape
    mask = dense.T > 0                                  # (D, V), row-major scan = ascending term
    nnz_per = mask.sum(axis=1)
    rowptr = np.zeros(D + 1, dtype=np.int64)
    np.cumsum(nnz_per, out=rowptr[1:])
    dd, vv = np.nonzero(mask)
    term = vv.astype(np.int32)
    count = dense.T[dd, vv].astype(np.int32)
    return rowptr, term, count


def format_counts_device(dense, layout=0, handle=None, device=0):
    """format_counts_lda / one modality of format_counts_mmctm (src/utils.jl:1-36) ON THE GPU
    (mmsig_format_counts): dense is (V, D) for layout 0 (term-major, the TSV layout) or (D, V) for
    layout 1 (sample-major, Julia's column-major V x D); int32 or int64.  Returns (rowptr, term0, count)."""
    import ctypes as C
    from . import capi
    a = np.ascontiguousarray(dense)
    if a.dtype not in (np.int32, np.int64):
        a = a.astype(np.int64)
    V, D = (a.shape if layout == 0 else a.shape[::-1])
    h = handle or capi.Handle(device=device)
    try:
        rowptr = np.zeros(D + 1, dtype=np.int64)
        nnz = C.c_int64()
        h.check(h.lib.mmsig_format_counts(h.h, D, V, a.ctypes.data_as(C.c_void_p), a.dtype.itemsize, layout,
                                          rowptr.ctypes.data_as(capi.c_i64p), C.byref(nnz)))
        term, count = np.zeros(nnz.value, dtype=np.int32), np.zeros(nnz.value, dtype=np.int32)
        h.check(h.lib.mmsig_format_counts_fetch(h.h, term.ctypes.data_as(capi.c_i32p), count.ctypes.data_as(capi.c_i32p)))
    finally:
        if handle is None:
            h.close()
    return rowptr, term, count


def format_counts_mmctm(dense_list):
    """src/utils.jl:24-36: one CSR triple per modality."""
    return [make_count_csr(x) for x in dense_list]


def format_counts_ctm(dense):
    """src/utils.jl:20-22."""
    return format_counts_mmctm([dense])


def format_counts_lda(dense):
    """src/utils.jl:9-18."""
    return make_count_csr(dense)


def from_nested(X, M):
    """Reference nested form (list over d of list over m of (nnz, 2) arrays, 1-based terms)
    -> list over m of CSR triples."""
    D = len(X)
    out = []
    for m in range(M):
        rowptr = np.zeros(D + 1, dtype=np.int64)
        terms, cnts = [], []
        for d in range(D):
            a = np.asarray(X[d][m], dtype=np.int64).reshape(-1, 2)
            rowptr[d + 1] = rowptr[d] + a.shape[0]
            terms.append(a[:, 0] - 1)
            cnts.append(a[:, 1])
        out.append((rowptr,
                    np.concatenate(terms).astype(np.int32) if terms else np.zeros(0, np.int32),
                    np.concatenate(cnts).astype(np.int32) if cnts else np.zeros(0, np.int32)))
    return out


def infer_V(counts):
    """MMCTM(k, alpha, X) without V: max observed term (src/MMCTM.jl:94-108)."""
    return [int(t.max()) + 1 if t.size else 0 for _, t, _ in counts]


def read_tsv(path):
    """TSV with a `term` column and one column per sample (data/brca-eu_*.tsv)."""
    with open(path) as f:
        header = f.readline().rstrip("\n").split("\t")
        rows, terms = [], []
        for line in f:
            p = line.rstrip("\n").split("\t")
            terms.append(p[0])
            rows.append([int(x) for x in p[1:]])
    return terms, header[1:], np.asarray(rows, dtype=np.int64)


def shard_rows(rowptr_list, n_shards):
    """Contiguous sample ranges per rank, balanced by total nonzeros.
    Returns boundaries b[0..n_shards], shard r = samples b[r]:b[r+1]."""
    D = len(rowptr_list[0]) - 1
    w = np.zeros(D + 1, dtype=np.float64)
    for rp in rowptr_list:
        w += np.asarray(rp, dtype=np.float64)
    w += np.arange(D + 1) * 16.0          # per-sample fixed cost keeps empty rows from piling up
    tot = w[-1]
    b = [0]
    for r in range(1, n_shards):
        b.append(int(np.searchsorted(w, tot * r / n_shards)))
    b.append(D)
    b = np.maximum.accumulate(np.asarray(b, dtype=np.int64))
    return b


def slice_csr(csr, lo, hi):
    rowptr, term, count = csr
    a, z = int(rowptr[lo]), int(rowptr[hi])
    return (rowptr[lo:hi + 1] - rowptr[lo]).astype(np.int64), term[a:z], count[a:z]
