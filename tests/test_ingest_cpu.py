"""Count ingest (reference src/utils.jl:1-36), CPU side: the oracle's restatement of
make_count_matrix / format_counts_* against a hand-derived answer (the reference has no test of
src/utils.jl) and the committed brca-eu fixture, and the host mirror (mmsig.counts) against the oracle."""
import numpy as np

import orc
from mmsig.counts import make_count_csr, format_counts_mmctm, from_nested


def test_hand_derived_answer_format_counts():
    # the reference has no test of src/utils.jl; this answer is derived by hand from
    # make_count_matrix (src/utils.jl:1-7): entries > 0 as [index count] rows in ascending index
    dense = np.array([[1, 0], [0, 2], [3, 0]], dtype=np.int64)            # V=3 terms, D=2 samples
    rowptr, term, cnt = orc.make_count_csr(dense, 0)
    assert rowptr.tolist() == [0, 2, 3]
    assert term.tolist() == [0, 2, 1] and cnt.tolist() == [1, 3, 2]     # 0-based terms: Julia's [1 1; 3 3] and [2 2]
    r1 = orc.make_count_csr(np.ascontiguousarray(dense.T), 1)
    assert all(np.array_equal(a, b) for a, b in zip(r1, (rowptr, term, cnt)))


def test_oracle_reproduces_brca_fixture(brca):
    for rowptr, term, cnt in brca:                                        # the CSR committed from data/brca-eu_*.tsv
        D, V = len(rowptr) - 1, int(term.max()) + 1
        dense = np.zeros((V, D), dtype=np.int64)
        for d in range(D):
            dense[term[rowptr[d]:rowptr[d + 1]], d] = cnt[rowptr[d]:rowptr[d + 1]]
        got = orc.make_count_csr(dense, 0)
        assert np.array_equal(got[0], rowptr) and np.array_equal(got[1], term) and np.array_equal(got[2], cnt)


def test_host_mirror_equals_oracle_with_edge_cases():
    rng = np.random.default_rng(4)
    for V, D in [(1, 1), (7, 33), (96, 257), (83, 1000)]:
        dense = rng.poisson(0.8, (V, D)).astype(np.int64)
        dense[:, rng.integers(0, D)] = 0                                  # an empty sample
        dense[rng.integers(0, V), rng.integers(0, D)] = -3                # entries <= 0 are dropped (counts .> 0)
        dense[rng.integers(0, V), rng.integers(0, D)] = 2**31 - 1
        a, b = orc.make_count_csr(dense, 0), make_count_csr(dense)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    z = orc.make_count_csr(np.zeros((5, 9), np.int64), 0)
    assert z[0].tolist() == [0] * 10 and z[1].size == 0
    import pytest
    with pytest.raises(OverflowError):
        orc.make_count_csr(np.full((2, 2), 2**31, np.int64), 0)
    # nested reference layout (X[d][m], 1-based terms) -> CSR
    X = [[np.array([[1, 4], [3, 1]])], [np.zeros((0, 2), int)]]
    r, t, c = from_nested(X, 1)[0]
    assert r.tolist() == [0, 2, 2] and t.tolist() == [0, 2] and c.tolist() == [4, 1]
    assert len(format_counts_mmctm([np.ones((2, 3), int), np.ones((4, 3), int)])) == 2
