"""The C++ façade (include/stfem_b200.hpp) mirrors the reference's operator interface: it must compile
as plain host C++17 against the C header (CPU check) and, on the GPU, reproduce the oracle through the same calls a
deal.II-style driver makes (vmult, GMG handed to FGMRES)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "facade_demo.cpp")
LIBDIR = os.path.join(ROOT, "dealii-stfem_b200")


def _compile(out):
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", out, "-L", LIBDIR, "-lstfem_b200",
           "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_facade_compiles_and_links(tmp_path):
    import dealii_stfem_b200 as st
    st.capi.lib()                      # the library must exist (no fallback)
    _compile(str(tmp_path / "facade_demo"))


@pytest.mark.gpu
def test_facade_matches_oracle(tmp_path):
    from oracle import fe_time as ft, spatial as S
    exe = _compile(str(tmp_path / "facade_demo"))
    out = str(tmp_path / "y.bin")
    n, k = 4, 2
    r = subprocess.run([exe, out, str(n), str(k)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"N (\d+) nb (\d+) iterations (\d+) .* rel_residual_inf (\S+)", r.stdout)
    assert m, r.stdout
    N, nb, its, res = int(m.group(1)), int(m.group(2)), int(m.group(3)), float(m.group(4))
    mesh = S.Mesh(3, [n, n, n], 0)
    space = S.Space(mesh, k)
    assert space.n_dofs == N
    A, B, _, _ = ft.get_fe_time_weights("DG", 1, 0.025, 1)
    sysm = S.SystemMatrix(S.MatrixFreeOperator(space, 0.0, 1.0), S.MatrixFreeOperator(space, 1.0, 0.0), A, B)
    x = np.sin(0.1 * np.arange(N)[None, :] + np.arange(nb)[:, None])
    ref = sysm.vmult(x)
    y = np.fromfile(out, dtype=np.float64).reshape(nb, N)
    assert np.abs(y - ref).max() / np.abs(ref).max() < 1e-12
    assert 1 <= its <= 40 and res < 1e-9
    md = re.search(r"diagonal_sum (\S+)", r.stdout)
    assert md, r.stdout
    dsum = sysm.get_matrix_diagonal().sum()
    assert abs(float(md.group(1)) - dsum) <= 1e-11 * abs(dsum)
    assert "non-square vmult rejected: yes" in r.stdout
