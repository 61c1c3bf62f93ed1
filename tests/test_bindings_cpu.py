"""The three statements of the drop-in boundary must agree argument for argument: the prototypes in
include/mmsig.h, the ctypes signatures in multimodalmusig.jl_b200/capi.py (what the tests and the bench
call through) and the `ccall`s of julia/MMSigB200.jl (what a MultiModalMuSig.jl maintainer loads; there
is no Julia in the build image, so this static check is what keeps the shim from drifting)."""
import ctypes as C
import os
import re

import mmsig
from conftest import ROOT


def _ckind(t):
    t = t.strip()
    if "*" in t or t.endswith("]"):
        return "ptr"
    t = re.sub(r"\bconst\b", "", t).split()
    base = t[0] if len(t) <= 2 else " ".join(t[:-1])
    return {"int32_t": "i32", "uint32_t": "u32", "int64_t": "i64", "uint64_t": "u64", "double": "f64",
            "int": "i32", "size_t": "u64", "void": "void"}[base]


def _header():
    src = open(os.path.join(ROOT, "include", "mmsig.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for ret, name, args in re.findall(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(mmsig_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src):
        args = " ".join(args.split())
        params = [] if args in ("", "void") else [_ckind(a) for a in args.split(",")]
        out[name] = (_ckind(ret) if "*" not in ret else "ptr", params)
    return out


_CT = {C.c_int32: "i32", C.c_uint32: "u32", C.c_int64: "i64", C.c_uint64: "u64", C.c_double: "f64",
       C.c_void_p: "ptr", C.c_char_p: "ptr", C.c_size_t: "u64", None: "void"}


def _ctkind(t):
    if t in _CT:
        return _CT[t]
    assert hasattr(t, "_type_") or hasattr(t, "contents"), t      # POINTER(...)
    return "ptr"


def test_ctypes_signatures_match_the_header():
    H = _header()
    assert len(H) >= 40
    for name, (res, args) in mmsig.capi._SIGS.items():
        hret, hargs = H[name]
        assert [_ctkind(a) for a in args] == hargs, name
        assert _ctkind(res) == hret, name


_JL = {"Int32": "i32", "UInt32": "u32", "Int64": "i64", "UInt64": "u64", "Float64": "f64", "Cint": "i32",
       "Cstring": "ptr", "Cvoid": "void"}


def _jlkind(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
        return "ptr"
    return _JL[t]


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out]


def _ccalls():
    src = open(os.path.join(ROOT, "julia", "MMSigB200.jl")).read()
    src = re.sub(r"#[^\n]*", "", src)
    calls = []
    for m in re.finditer(r"ccall\(\(:(mmsig_[a-z0-9_]+),\s*LIB\)\s*,", src):
        i, depth = m.end(), 1                       # walk to the matching ')' of ccall(
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        parts = _split_top(src[m.end():i - 1])
        ret, types, vals = parts[0], parts[1], parts[2:]
        assert types.startswith("(") and types.endswith(")"), (m.group(1), types)
        calls.append((m.group(1), ret, _split_top(types[1:-1]), vals))
    return calls


def test_julia_ccalls_match_the_header():
    H = _header()
    calls = _ccalls()
    assert len(calls) >= 30
    for name, ret, types, vals in calls:
        assert name in H, "%s is not declared in include/mmsig.h" % name
        hret, hargs = H[name]
        assert [_jlkind(t) for t in types] == hargs, (name, types, hargs)
        assert len(vals) == len(types), (name, "argument count differs from the type tuple")
        assert _jlkind(ret) == hret or (hret == "ptr" and ret.startswith("Ptr")), (name, ret)


def test_julia_shim_covers_the_model_families():
    src = open(os.path.join(ROOT, "julia", "MMSigB200.jl")).read()
    for sig in ("fit!(model::MMCTM", "fit!(model::IMMCTM", "fit!(model::LDA", "fit!(model::ILDA", "function fit_heldout",
                "function transform"):
        assert sig in src, sig


def test_julia_shim_blocks_and_brackets_balance():
    """No Julia in the image: at least every block opener has its `end` and every bracket closes."""
    src = open(os.path.join(ROOT, "julia", "MMSigB200.jl")).read()
    src = re.sub(r'"(?:\\.|[^"\\])*"', '""', src)
    src = re.sub(r"#[^\n]*", "", src)
    prev = None
    while prev != src:                      # bracketed content goes first (`end` as an index, comprehensions)
        prev = src
        src = re.sub(r"\[[^\[\]()]*\]", "_", src)
        src = re.sub(r"\([^()\[\]]*\)", "_", src)
    assert src.count("(") == src.count(")") and src.count("[") == src.count("]")
    depth = 0
    for t in re.findall(r"\b(function|for|if|while|try|let|begin|do|struct|module|quote|end)\b", src):
        depth += -1 if t == "end" else 1
        assert depth >= 0
    assert depth == 0


def test_julia_shim_touches_only_fields_the_reference_structs_have():
    """tests/golden/reference_struct_fields.json: the field names of the reference's model structs
    (generated by tests/golden/make_struct_fields.py in the build container)."""
    import json
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_struct_fields.json")))
    shim = open(os.path.join(ROOT, "julia", "MMSigB200.jl")).read()
    seen = 0
    for m in re.finditer(r"function (\w+!?)\((?:[^)]*?)(model|newmodel)::(\w+)(.*?)\nend\n", shim, re.S):
        typ, body = m.group(3), m.group(4)
        used = set(re.findall(r"\b(?:model|newmodel|heldout_model)\.([^\s\.\[\]\(\),;=+\-*/:]+)", body))
        assert used and used <= set(ref[typ]), (m.group(1), typ, used - set(ref[typ]))
        seen += 1
    assert seen >= 7


def test_config_struct_fields_agree():
    """mmsig_config (passed by pointer to mmsig_create / mmsig_group_create): the same fields in the same order and
    widths in include/mmsig.h, capi.Config and the Julia shim's MmsigConfig -- a drifted field would silently hand the
    library a wrong `precision` or `profile`."""
    src = open(os.path.join(ROOT, "include", "mmsig.h")).read()
    body = re.search(r"typedef struct \{(.*?)\}\s*mmsig_config;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    hfields = []
    for typ, name, arr in re.findall(r"(int32_t)\s+([a-z_]+)(?:\[(\d+)\])?;", body):
        hfields.append((name, int(arr) if arr else 1))
    assert [f[0] for f in hfields] == ["device", "stop_rule", "profile", "precision", "reserved"], hfields
    cfields = []
    for name, typ in mmsig.capi.Config._fields_:
        n = typ._length_ if hasattr(typ, "_length_") else 1
        base = typ._type_ if hasattr(typ, "_length_") else typ
        assert base is C.c_int32, name
        cfields.append((name, n))
    assert cfields == hfields
    assert C.sizeof(mmsig.capi.Config) == 4 * sum(n for _, n in hfields) == 32
    jl = open(os.path.join(ROOT, "julia", "MMSigB200.jl")).read()
    jbody = re.search(r"struct MmsigConfig\n(.*?)\nend", jl, flags=re.S).group(1)
    jfields = []
    for name, typ in re.findall(r"^\s*([a-z_]+)::(\S+)", jbody, flags=re.M):
        m = re.match(r"NTuple\{(\d+),Int32\}", typ)
        assert typ == "Int32" or m, (name, typ)
        jfields.append((name, int(m.group(1)) if m else 1))
    assert jfields == hfields
    # every constructor call of the shim passes as many values as the struct has fields
    for call in re.findall(r"MmsigConfig\((.*?)\)\)", jl):
        assert call.count("Int32(") + call.count("ids[1]") >= 4 and "ntuple" in call, call
    assert mmsig.capi.PRECISION_FP64 == 0 and mmsig.capi.PRECISION_FP32 == 1
    assert re.search(r"#define MMSIG_PRECISION_FP32\s+1", src) and re.search(r"#define MMSIG_PRECISION_FP64\s+0", src)
