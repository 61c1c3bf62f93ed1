"""format_counts_* on the device (mmsig_format_counts, mmsig_*_set_data_dense) against the oracle's
restatement of src/utils.jl:1-36: integer work, bit-exact."""
import numpy as np
import pytest

import orc
import mmsig
from mmsig import capi
from mmsig.counts import format_counts_device, make_count_csr

pytestmark = pytest.mark.gpu


def _same(a, b):
    return all(np.array_equal(x, y) and x.dtype == y.dtype for x, y in zip(a, b))


@pytest.mark.parametrize("dtype", [np.int32, np.int64])
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("V,D", [(1, 1), (5, 31), (96, 257), (83, 5000), (32, 70001), (200, 3000), (1500, 300)])
def test_format_counts_equals_oracle(dtype, layout, V, D):
    rng = np.random.default_rng(V * 1000 + D)
    dense = rng.poisson(0.9, (V, D)).astype(np.int64)
    dense[:, rng.integers(0, D, size=max(1, D // 20))] = 0            # empty samples
    dense[rng.integers(0, V), rng.integers(0, D)] = -7                # dropped (counts .> 0)
    dense[rng.integers(0, V), rng.integers(0, D)] = 2**31 - 1         # the largest representable count
    ref = orc.make_count_csr(dense, 0)
    a = dense if layout == 0 else np.ascontiguousarray(dense.T)
    got = format_counts_device(a.astype(dtype), layout=layout)
    assert _same(got, ref)


def test_format_counts_all_zero_and_overflow():
    got = format_counts_device(np.zeros((9, 100), np.int32))
    assert got[0].tolist() == [0] * 101 and got[1].size == 0 and got[2].size == 0
    with pytest.raises(capi.MmsigError) as e:
        format_counts_device(np.full((3, 3), 2**31, np.int64))
    assert e.value.code == -6


def test_format_counts_brca_fixture(brca):
    for rowptr, term, cnt in brca:
        D, V = len(rowptr) - 1, int(term.max()) + 1
        dense = np.zeros((V, D), dtype=np.int64)
        for d in range(D):
            dense[term[rowptr[d]:rowptr[d + 1]], d] = cnt[rowptr[d]:rowptr[d + 1]]
        got = format_counts_device(dense)
        assert np.array_equal(got[0], rowptr) and np.array_equal(got[1], term) and np.array_equal(got[2], cnt)


def test_format_counts_large_properties():
    """BASELINE-sized modality (SNV: 96 terms) at D = 400k: size-independent properties -- totals and
    nonzero counts preserved, rows ascending, and a checksum of (d, term, count) triples equal to
    the dense matrix's."""
    V, D = 96, 400_000
    rng = np.random.default_rng(5)
    dense = rng.poisson(1.3, (V, D)).astype(np.int32)
    rowptr, term, cnt = format_counts_device(dense)
    assert rowptr[-1] == np.count_nonzero(dense) and int(cnt.sum()) == int(dense.sum())
    assert np.array_equal(np.diff(rowptr), (dense > 0).sum(axis=0))
    d_of = np.repeat(np.arange(D), np.diff(rowptr))
    inner = np.ones(term.size, bool)
    inner[rowptr[:-1][np.diff(rowptr) > 0]] = False                   # first record of each row
    assert np.all(np.diff(term)[inner[1:]] > 0)
    chk = lambda dd, vv, cc: int(((dd * 1315423911 + vv * 2654435761 + cc * 97) % (2**61 - 1)).sum() % (2**61 - 1))
    vv, dd = np.nonzero(dense)
    assert chk(d_of.astype(np.int64), term.astype(np.int64), cnt.astype(np.int64)) == \
        chk(dd.astype(np.int64), vv.astype(np.int64), dense[vv, dd].astype(np.int64))


@pytest.mark.parametrize("layout", [0, 1])
def test_set_data_dense_equals_csr_path(layout):
    K, V, D = [4, 3], [20, 9], 700
    rng = np.random.default_rng(2)
    dense = [rng.poisson(1.0, (v, D)).astype(np.int64) for v in V]
    dense[1][:, :40] = 0
    counts = [make_count_csr(x) for x in dense]
    g0 = mmsig.synth.init_gamma(K, V)
    a = mmsig.MMCTM(K, [0.1, 0.1], counts, V=V, gamma0=g0)
    dd = dense if layout == 0 else [np.ascontiguousarray(x.T) for x in dense]
    b = mmsig.MMCTM(K, [0.1, 0.1], None, V=V, gamma0=g0, dense=dd, dense_layout=layout)
    for _ in range(3):
        assert np.array_equal(a.iterate(), b.iterate())
    sa, sb = a.state(), b.state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    assert a.calculate_elbo()[0] == b.calculate_elbo()[0]
    a.close(); b.close()


def test_lda_set_data_dense_equals_csr_path():
    K, V, D = 5, 30, 900
    rng = np.random.default_rng(3)
    dense = rng.poisson(0.7, (V, D)).astype(np.int32)
    lam0 = mmsig.synth.init_lda_lambda(K, V)
    a = mmsig.LDA(K, 0.1, 0.1, make_count_csr(dense), V=V, lambda0=lam0)
    b = mmsig.LDA(K, 0.1, 0.1, None, V=V, lambda0=lam0, dense=dense)
    for _ in range(3):
        assert a.iterate() == b.iterate()
    a.close(); b.close()


def test_tsv_file_to_device_csr(tmp_path):
    """data/*.tsv-shaped files -> C++ reader -> dense term-major matrices -> CSR built on the GPU ->
    the same iterations as the host path (read_tsv + format_counts_mmctm + set_data)."""
    from mmsig import io
    from mmsig.counts import read_tsv, format_counts_mmctm
    K, V, D = [3, 2], [12, 7], 300
    rng = np.random.default_rng(8)
    paths, dense = [], []
    for m, v in enumerate(V):
        x = rng.poisson(1.2, (v, D))
        p = tmp_path / ("m%d.tsv" % m)
        io.write_counts_tsv(p, ["t%d" % i for i in range(v)], ["s%d" % i for i in range(D)], x)
        paths.append(p)
        dense.append(x)
    native = [io.read_counts_tsv_native(p)[2] for p in paths]
    assert all(np.array_equal(a, b) for a, b in zip(native, dense))
    g0 = mmsig.synth.init_gamma(K, V)
    a = mmsig.MMCTM(K, [0.1, 0.1], format_counts_mmctm([read_tsv(p)[2] for p in paths]), V=V, gamma0=g0)
    b = mmsig.MMCTM(K, [0.1, 0.1], None, V=V, gamma0=g0, dense=native)
    for _ in range(2):
        assert np.array_equal(a.iterate(), b.iterate())
    sa, sb = a.state(), b.state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    a.close(); b.close()
