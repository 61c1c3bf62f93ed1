"""On-disk formats: count TSV round trip and the result tables of
scripts/run_mmctm.jl:184-290 (host logic, no GPU)."""
import numpy as np

from mmsig import io
from mmsig.counts import make_count_csr, read_tsv


def test_julia_float_str_round_trips_and_style():
    assert io.julia_float_str(0.1) == "0.1" and io.julia_float_str(1.0) == "1.0"
    assert io.julia_float_str(1e-5) == "1.0e-5" and io.julia_float_str(0.0001) == "0.0001"
    assert io.julia_float_str(123456.7) == "123456.7" and io.julia_float_str(1234567.8) == "1.2345678e6"
    assert io.julia_float_str(-2.5e-7) == "-2.5e-7" and io.julia_float_str(0.0) == "0.0"
    assert io.julia_float_str(100000.0) == "100000.0" and io.julia_float_str(1e6) == "1.0e6"
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.standard_normal(2000) * 10.0 ** rng.integers(-12, 12, 2000), [5e-324, 1.7976931348623157e308]])
    for x in xs:
        assert float(io.julia_float_str(x)) == x


def test_counts_tsv_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    dense = rng.poisson(2.0, (6, 9))
    terms = ["A[C->A]%d" % i for i in range(6)]
    samples = ["DO%d" % i for i in range(9)]
    p = tmp_path / "c.tsv"
    io.write_counts_tsv(p, terms, samples, dense)
    t2, s2, d2 = read_tsv(p)
    assert t2 == terms and s2 == samples and np.array_equal(d2, dense)
    r, t, c = make_count_csr(d2)
    assert r[-1] == np.count_nonzero(dense) and c.sum() == dense.sum()


def test_result_tables(tmp_path):
    K, V = [2, 3], [4, 2]
    rng = np.random.default_rng(2)
    gamma = rng.integers(1, 50, sum(k * v for k, v in zip(K, V))).astype(float)
    lam = rng.standard_normal((5, sum(K)))
    terms = [["t%d" % i for i in range(4)], ["u0", "u1"]]
    rows = io.topic_table(K, V, gamma, terms, ["SNV", "SV"])
    assert len(rows) == 2 * 4 + 3 * 2 and rows[0][:4] == ("SNV", 1, 1, "t0") and rows[-1][:4] == ("SV", 3, 2, "u1")
    assert abs(sum(r[4] for r in rows[:4]) - 1.0) < 1e-15 and rows[0][4] == gamma[0] / gamma[:4].sum()
    labels, P = io.props_table(K, lam, ["SNV", "SV"])
    assert labels == ["SNV-1", "SNV-2", "SV-1", "SV-2", "SV-3"] and P.shape == (5, 5)
    assert np.allclose(P[:2].sum(axis=0), 1.0) and np.allclose(P[2:].sum(axis=0), 1.0)
    e = np.exp(lam[3, 2:5])
    assert np.array_equal(P[2:, 3], e / e.sum())
    io.write_sigs(tmp_path / "sigs.tsv", K, V, gamma, terms, ["SNV", "SV"])
    lines = (tmp_path / "sigs.tsv").read_text().splitlines()
    assert lines[0] == "modality\ttopic\tvalue\tterm\tprobability" and len(lines) == 15
    assert float(lines[1].split("\t")[4]) == rows[0][4]
    io.write_props(tmp_path / "props.tsv", K, lam, ["a", "b", "c", "d", "e"], ["SNV", "SV"])
    pl = (tmp_path / "props.tsv").read_text().splitlines()
    assert pl[0] == "topic\ta\tb\tc\td\te" and pl[1].startswith("SNV-1\t")
    assert np.array_equal(np.array([[float(x) for x in l.split("\t")[1:]] for l in pl[1:]]), P)
    S = np.array([[4.0, 2.0], [2.0, 9.0]])
    assert np.allclose(io.cov2cor(S), [[1, 1 / 3], [1 / 3, 1]])
    io.write_dlm(tmp_path / "cov.tsv", S)
    assert np.array_equal(io.read_dlm(tmp_path / "cov.tsv"), S)
    io.write_dlm(tmp_path / "mean.tsv", np.array([0.5, -1.25]))
    assert (tmp_path / "mean.tsv").read_text() == "0.5\n-1.25\n"


def test_native_tsv_reader_equals_python_reader(tmp_path):
    """mmsig_tsv_dims / mmsig_tsv_read (host-side C++ in libmmsig.so, no GPU needed) against read_tsv."""
    import pytest
    rng = np.random.default_rng(3)
    dense = rng.poisson(40.0, (96, 57))
    dense[5, 7] = 0
    dense[0, 0] = 2**31 - 1
    terms = ["A[C->%s]%d" % ("ATG"[i % 3], i) for i in range(96)]
    samples = ["DO%d" % i for i in range(57)]
    p = tmp_path / "c.tsv"
    io.write_counts_tsv(p, terms, samples, dense)
    t, s, d = io.read_counts_tsv_native(p)
    t2, s2, d2 = read_tsv(p)
    assert t == t2 == terms and s == s2 == samples
    assert d.dtype == np.int32 and np.array_equal(d, d2)
    # Windows line ends and a trailing blank line
    q = tmp_path / "crlf.tsv"
    q.write_bytes(p.read_bytes().replace(b"\n", b"\r\n") + b"\r\n")
    assert np.array_equal(io.read_counts_tsv_native(q)[2], d2)
    # ragged row, non-integer, overflow, missing file
    (tmp_path / "bad1.tsv").write_text("term\ta\tb\nx\t1\ny\t1\t2\n")
    (tmp_path / "bad2.tsv").write_text("term\ta\nx\t1.5\n")
    (tmp_path / "bad3.tsv").write_text("term\ta\nx\t2147483648\n")
    for name in ("bad1.tsv", "bad2.tsv", "bad3.tsv", "missing.tsv"):
        with pytest.raises(ValueError):
            io.read_counts_tsv_native(tmp_path / name)
