"""CPU-side checks of the drop-in boundary: libmmsig.so loads, exports every symbol that
include/mmsig.h declares, and refuses to run without a GPU (no CPU fallback)."""
import os
import re

import pytest

import mmsig
from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "mmsig.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmsig_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = mmsig.capi.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libmmsig.so does not export %s" % n
    assert sorted(mmsig.capi.EXPORTS) == names, "capi.py and mmsig.h disagree"
    assert lib.mmsig_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mmsig.capi.MmsigError) as e:
        mmsig.capi.Handle()
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """The product path may mention the oracle in comments, never import, include, link or call it."""
    pkg = os.path.join(ROOT, "multimodalmusig.jl_b200")
    banned = ("import orc", "from orc", "liboracle", "mmsig_oracle", "orc_mmctm", "orc_lda", "orc_mma", "../oracle",
              "oracle/")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                for b in banned:
                    assert b not in txt, (f, b)
    for f in ("mmsig.py",):
        txt = open(os.path.join(ROOT, f)).read()
        assert "orc" not in txt


def test_limits_need_no_device():
    import ctypes as C
    import mmsig
    lib = mmsig.capi.load()
    v = [C.c_int32() for _ in range(5)]
    assert lib.mmsig_limits(*[C.byref(x) for x in v]) == 0
    assert [x.value for x in v] == [8, 64, 32, 1024, 65535]
    assert lib.mmsig_version() >= 111
