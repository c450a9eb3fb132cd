"""The C-ABI library loads and exports every symbol include/kaiju_b200.h declares; host-side entry points work without a
GPU; device entry points fail loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os, re
import numpy as np
import pytest
from conftest import ROOT


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "kaiju_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(kj_[a-z_0-9]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported(built):
    import kaiju_b200 as kb
    L = kb.lib(); syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), "missing export " + s


def test_loaders_and_views(built, golden):
    import kaiju_b200 as kb
    L = kb.lib()
    f = C.c_void_p(); assert L.kj_fmi_load(golden.fmi.encode(), C.byref(f)) == 0
    v = kb.KjIndexView(); L.kj_fmi_view(f, C.byref(v))
    assert v.alen == 21 and v.alphabet[:21] == b"*ACDEFGHIKLMNPQRSTVWY" and v.bwtlen > 0 and v.nseq > 800 and v.chpt_exp == 3
    t = C.c_void_p(); assert L.kj_nodes_load(golden.nodes.encode(), C.byref(t)) == 0
    tv = kb.KjTaxonomyView(); L.kj_nodes_view(t, C.byref(tv)); assert tv.n == 75
    L.kj_fmi_free(f); L.kj_nodes_free(t)
    # error paths return codes, never exit()
    assert L.kj_fmi_load(b"/nonexistent.fmi", C.byref(f)) == -2 and b"could not open" in L.kj_last_error()
    assert L.kj_fmi_load(golden.nodes.encode(), C.byref(f)) == -2


def test_no_cpu_fallback(built, golden):
    import torch
    import kaiju_b200 as kb
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(kb.KaijuError) as e:
        kb.Classifier(golden.fmi, golden.nodes)
    assert "-4" in str(e.value) or "no CUDA device" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """The product sources must not include/link/call anything under oracle/ or the emulator."""
    for d, _, files in os.walk(os.path.join(ROOT, "kaiju_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(d, f)).read()
                assert "oracle/" not in txt.replace("no oracle/", "") or f == "__init__.py" and "oracle/" not in txt, f
                assert "liboracle" not in txt and "kjemu_" not in txt and "ko_classify" not in txt, f
