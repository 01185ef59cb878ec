"""The C-ABI library loads and exports every symbol include/stfem_b200.h declares (no GPU needed)."""
import os
import re

import dealii_stfem_b200 as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "stfem_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(stfem_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported_and_bound():
    L = st.capi.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "symbol %s declared in include/stfem_b200.h is not exported" % n
        assert n in st.capi.SYMBOLS, "symbol %s has no ctypes binding" % n
    for n in st.capi.SYMBOLS:
        assert n in names, "binding %s is not declared in the header" % n


def test_version_and_error_paths_without_gpu():
    L = st.capi.lib()
    assert b"sm_100a" in L.stfem_version()
    # invalid arguments are reported through return codes + stfem_last_error, never by aborting
    assert L.stfem_quadrature_rule(0, 0, None, None) != 0
    assert b"stfem_quadrature_rule" in L.stfem_last_error()
    assert L.stfem_fe_time_weights(7, 1, 1.0, 1, None, None, None, None) != 0
