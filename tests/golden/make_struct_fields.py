"""Field names of the reference's model structs (src/MMCTM.jl, IMMCTM.jl, LDA.jl, ILDA.jl), written to
reference_struct_fields.json so that tests/test_bindings_cpu.py can check, without Julia and without
/root/reference, that julia/MMSigB200.jl only touches fields that exist.  Run in the build container:
    python tests/golden/make_struct_fields.py"""
import json
import os
import re

REF = "/root/reference/src"
out = {}
for name in ("MMCTM", "IMMCTM", "LDA", "ILDA"):
    s = open(os.path.join(REF, name + ".jl")).read()
    m = re.search(r"mutable struct %s\b(.*?)\n\s*function %s" % (name, name), s, re.S)
    out[name] = sorted(set(re.findall(r"^\s*([^\s:#]+)::", m.group(1), re.M)))
json.dump(out, open(os.path.join(os.path.dirname(__file__), "reference_struct_fields.json"), "w"), ensure_ascii=False,
          indent=1)
print({k: len(v) for k, v in out.items()})
