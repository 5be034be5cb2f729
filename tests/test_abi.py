"""The C-ABI library builds, loads, and exports every symbol declared in include/mgld.h (no compute, no GPU)."""
import ctypes
import os
import re

from common import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "mgld.h")).read()
    return sorted(set(re.findall(r"^(?:int|const char\*)\s+(mgld_[a-z0-9_]+)\s*\(", txt, flags=re.M)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from mgld_vsr_b200 import lib
    handle = lib.load()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in mgld.h but not exported by libmgld.so"
    assert handle.mgld_abi_version() == 2


def test_struct_mirrors_match_header_layout():
    from mgld_vsr_b200 import lib
    # field counts/order are checked against the header text so a header edit cannot silently desync ctypes
    txt = open(os.path.join(ROOT, "include", "mgld.h")).read()
    for name, cls in (("mgld_conv_gemm_desc", lib.ConvGemmDesc), ("mgld_attention_desc", lib.AttentionDesc)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), txt, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(",")
            first = names[0].split()[-1].lstrip("*")
            fields.append(first)
            fields += [n.strip().lstrip("*") for n in names[1:]]
        assert fields == [f[0] for f in cls._fields_], (name, fields)


def test_no_fallback_without_gpu():
    import pytest
    import torch
    from mgld_vsr_b200 import lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.MgldError):
        lib.lib()
