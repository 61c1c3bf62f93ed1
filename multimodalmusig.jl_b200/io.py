"""On-disk formats around the hot path: the count TSVs the reference reads
(data/*.tsv: a `term` column, one column per sample) and the result tables its driver script
writes (scripts/run_mmctm.jl:184-245,272-290: signatures, proportions, mean, covariance,
correlation).  Host-side text formats; the numbers come from the GPU state (model.state()).
"""
import numpy as np

from .counts import read_tsv  # noqa: F401  (re-exported: the reader lives with the CSR helpers)


def read_counts_tsv_native(path):
    """The integer body of a count TSV through the library's C++ reader (mmsig_tsv_dims / _read):
    a (V, D) int32 matrix, term-major, ready for MMCTM(..., dense=[...]) / format_counts_device.
    Names come from the first column / header as in read_tsv."""
    import ctypes as C
    from . import capi
    lib = capi.load()
    V, D = C.c_int64(), C.c_int64()
    b = str(path).encode()
    if lib.mmsig_tsv_dims(b, C.byref(V), C.byref(D)) != 0:
        raise ValueError((lib.mmsig_last_error(None) or b"").decode())
    dense = np.empty((V.value, D.value), dtype=np.int32)
    if lib.mmsig_tsv_read(b, V.value, D.value, dense.ctypes.data_as(capi.c_i32p)) != 0:
        raise ValueError((lib.mmsig_last_error(None) or b"").decode())
    with open(path) as f:
        samples = f.readline().rstrip("\r\n").split("\t")[1:]
        terms = [line.split("\t", 1)[0] for line in f if line.strip()]
    return terms, samples, dense


def julia_float_str(x):
    """Shortest round-trip decimal in the style Julia's `show(::Float64)` / CSV.write / writedlm
    use (digits as Python's repr; plain notation for 1e-4 <= |x| < 1e6, else d.ddde±x)
    [restated from memory of Base.Ryu.writeshortest; the digits are exact either way]."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Inf" if x > 0 else "-Inf"
    if x == 0.0:
        return "-0.0" if np.signbit(x) else "0.0"
    from decimal import Decimal
    sign, dig, exp = Decimal(repr(x)).as_tuple()
    digits = "".join(map(str, dig)).rstrip("0") or "0"
    e10 = len(dig) + exp - 1                       # decimal exponent of the leading digit
    sgn = "-" if sign else ""
    if -5 < e10 < 6:
        if e10 >= 0:
            ipart = digits[:e10 + 1].ljust(e10 + 1, "0")
            fpart = digits[e10 + 1:] or "0"
        else:
            ipart, fpart = "0", "0" * (-e10 - 1) + digits
        return "%s%s.%s" % (sgn, ipart, fpart)
    return "%s%s.%se%d" % (sgn, digits[0], digits[1:] or "0", e10)


def cov2cor(C):
    """scripts/run_mmctm.jl:184-187."""
    C = np.asarray(C, dtype=np.float64)
    s = np.sqrt(np.diag(C))
    return C / np.outer(s, s)


def topic_table(K, V, gamma, terms, modalities):
    """topicdf (scripts/run_mmctm.jl:189-209): rows (modality, topic, value, term, probability) with
    probability = γ[m][k] ./ sum(γ[m][k]); topic and value are 1-based as in the reference."""
    rows, off = [], 0
    for m, (k_m, v_m) in enumerate(zip(K, V)):
        for k in range(k_m):
            g = np.asarray(gamma[off:off + v_m], dtype=np.float64)
            p = g / g.sum()
            for v in range(v_m):
                rows.append((modalities[m], k + 1, v + 1, terms[m][v], float(p[v])))
            off += v_m
    return rows


def write_sigs(path, K, V, gamma, terms, modalities):
    """writesigs (scripts/run_mmctm.jl:211-214): tab-separated with a header line."""
    with open(path, "w") as f:
        f.write("modality\ttopic\tvalue\tterm\tprobability\n")
        for mo, k, v, t, p in topic_table(K, V, gamma, terms, modalities):
            f.write("%s\t%d\t%d\t%s\t%s\n" % (mo, k, v, t, julia_float_str(p)))


def props_table(K, lam, modalities):
    """propdf (scripts/run_mmctm.jl:216-241): per sample and modality softmax of the λ block (no
    max-subtraction, as the reference).  Returns (labels [ΣK], props [ΣK, D])."""
    lam = np.asarray(lam, dtype=np.float64)
    D = lam.shape[0]
    out = np.empty((sum(K), D))
    start = 0
    for k_m in K:
        e = np.exp(lam[:, start:start + k_m])
        out[start:start + k_m, :] = (e / e.sum(axis=1, keepdims=True)).T
        start += k_m
    labels = ["%s-%d" % (modalities[m], k + 1) for m in range(len(K)) for k in range(K[m])]
    return labels, out


def write_props(path, K, lam, samples, modalities):
    """writeprops (scripts/run_mmctm.jl:243-246)."""
    labels, P = props_table(K, lam, modalities)
    with open(path, "w") as f:
        f.write("topic\t" + "\t".join(samples) + "\n")
        for lab, row in zip(labels, P):
            f.write(lab + "\t" + "\t".join(julia_float_str(x) for x in row) + "\n")


def write_dlm(path, A):
    """writedlm(file, A) as scripts/run_mmctm.jl:276-284 uses it for μ, Σ and cor(Σ): tab-delimited
    rows, a vector as one value per line."""
    A = np.asarray(A, dtype=np.float64)
    if A.ndim == 1:
        A = A[:, None]
    with open(path, "w") as f:
        for row in A:
            f.write("\t".join(julia_float_str(x) for x in row) + "\n")


def read_dlm(path):
    return np.loadtxt(path, delimiter="\t", ndmin=2)


def write_counts_tsv(path, terms, samples, dense):
    """The layout of data/*.tsv (README.md:14-16): header `term<TAB>sample...`, one line per term."""
    dense = np.asarray(dense)
    with open(path, "w") as f:
        f.write("term\t" + "\t".join(samples) + "\n")
        for t, row in zip(terms, dense):
            f.write(t + "\t" + "\t".join(str(int(x)) for x in row) + "\n")


def write_model_outputs(model, prefix, terms, samples, modalities):
    """Everything main() of scripts/run_mmctm.jl writes after the fit (:272-290), from a fitted
    mmsig.MMCTM: <prefix>mean.tsv, cov.tsv, cor.tsv, sigs.tsv, props.tsv."""
    s = model.state(props=False)
    write_dlm(prefix + "mean.tsv", s["mu"])
    write_dlm(prefix + "cov.tsv", s["Sigma"])
    write_dlm(prefix + "cor.tsv", cov2cor(s["Sigma"]))
    write_sigs(prefix + "sigs.tsv", model.K, model.V, s["gamma"], terms, modalities)
    write_props(prefix + "props.tsv", model.K, s["lam"], samples, modalities)
