"""kernel_variant 60 (EXPERIMENTAL fast-diagonalisation form of the Cartesian operator, csrc/st_vmult_cart_fd.cuh): the
host-side modes (stfem_cart_fd_modes) and the algebra the kernel implements, emulated in numpy cell by cell and compared
with the oracle's SystemMatrix::vmult.  CPU only; the CUDA kernel itself is exercised by tests/test_cart_fd_gpu.py."""
import ctypes as C

import numpy as np
import pytest

import dealii_stfem_b200 as st
from dealii_stfem_b200 import fe_time_host as fth
from oracle import quadrature as Q
from oracle import spatial as S


def modes(degree):
    n1 = degree + 1
    V, lam = np.zeros((n1, n1)), np.zeros(n1)
    st.capi.check(st.capi.lib().stfem_cart_fd_modes(degree, st.capi._dptr(V), st.capi._dptr(lam)))
    return V, lam


@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6])
def test_modes_diagonalise_the_reference_cell_pencil(degree):
    """Mh = V^T V, Kh = V^T diag(lam) V for Mh = S^T W S, Kh = D^T W D (QGauss(k+1), GLL nodes)."""
    n1 = degree + 1
    gll = Q.gauss_lobatto(n1)[0]
    xq, wq = Q.gauss(n1)
    Sm, Dm = Q.lagrange_eval(gll, xq).T, Q.lagrange_deriv(gll, xq).T         # [q, i]
    Mh, Kh = Sm.T @ (wq[:, None] * Sm), Dm.T @ (wq[:, None] * Dm)
    V, lam = modes(degree)
    assert np.abs(V.T @ V - Mh).max() < 1e-14
    assert np.abs(V.T @ (lam[:, None] * V) - Kh).max() < 1e-12 * np.abs(Kh).max()
    assert np.sum(np.abs(lam) < 1e-9) == 1 and np.all(lam > -1e-9)            # one constant mode, the rest positive


@pytest.mark.parametrize("degree,ttype,r,coef", [(2, "DG", 1, False), (4, "CGP", 2, False), (3, "DG", 2, True)])
def test_fast_diagonalisation_form_equals_the_operator(degree, ttype, r, coef):
    """A = sum_cells P_c^T vol (V^T x V^T x V^T) [Beta + c (lx/hx^2 + ly/hy^2 + lz/hz^2) Alpha] (V x V x V) P_c with
    constrained nodes read as 0 and not written: exactly what k_st_vmult_cart_fd computes."""
    n = [3, 2, 2]
    lo, up = [0.0, 0.0, 0.0], [1.5, 1.0, 0.8]
    mesh = S.Mesh(3, n, 0, lo, up)
    space = S.Space(mesh, degree)
    A, B, _, _ = fth.get_fe_time_weights(ttype, r, 0.05, 1)
    nb = A.shape[0]
    K, M = S.MatrixFreeOperator(space, 0.0, 1.0), S.MatrixFreeOperator(space, 1.0, 0.0)
    cc = np.ones(mesh.n_cells)
    if coef:
        cc = 1.0 + np.arange(mesh.n_cells) % 3
        K.laplace_coeff = np.repeat(cc[:, None], (degree + 1) ** 3, axis=1)
    src = np.stack([np.random.RandomState(3 + b).uniform(-1, 1, space.n_dofs) for b in range(nb)])
    ref = S.SystemMatrix(K, M, A, B).vmult(src)
    V, lam = modes(degree)
    h = [(up[d] - lo[d]) / n[d] for d in range(3)]
    vol = h[0] * h[1] * h[2]
    n1 = degree + 1
    free = ~space.constrained
    out = np.zeros_like(src)
    lsum = (lam[None, None, :] / h[0] ** 2 + lam[None, :, None] / h[1] ** 2 + lam[:, None, None] / h[2] ** 2)    # [mz, my, mx]
    for c in range(mesh.n_cells):
        dofs = space.cell_dofs[c]
        u = np.where(free[dofs], src[:, dofs], 0.0).reshape(nb, n1, n1, n1)                       # [b, z, y, x]
        m = np.einsum("az,by,cx,szyx->sabc", V, V, V, u)
        mode = B[:, :, None, None, None] + cc[c] * lsum[None, None] * A[:, :, None, None, None]
        m2 = np.einsum("rsabc,sabc->rabc", mode, m)
        loc = vol * np.einsum("az,by,cx,rabc->rzyx", V, V, V, m2).reshape(nb, -1)
        np.add.at(out, (slice(None), dofs), np.where(free[dofs], loc, 0.0))
    assert np.abs(out - ref).max() <= 1e-13 * np.abs(ref).max()


# ---- the kernel source itself, compiled for the host (tests/cpp/cart_fd_host_emulation.cpp: one std::thread per CUDA
# thread, std::barrier for __syncthreads, a global array for the shared memory) and run on small meshes: checks the thread
# mapping, the shared-memory layout, the ragged last CTA and the Dirichlet masking of k_st_vmult_cart_fd against the oracle.
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cartfd") / "cart_fd_emul")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-Wno-unknown-pragmas", "-o", exe,
                        os.path.join(ROOT, "tests", "cpp", "cart_fd_host_emulation.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_emulator_uses_the_exchange_layout_of_the_cuda_header():
    """The harness restates ExchLayout<3..5>; the constants must be the ones of st_vmult_cart.cuh."""
    cu = open(os.path.join(ROOT, "dealii-stfem_b200", "csrc", "st_vmult_cart.cuh")).read()
    em = open(os.path.join(ROOT, "tests", "cpp", "cart_fd_host_emulation.cpp")).read()
    pat = r"struct ExchLayout<(\d)> \{ static constexpr int LS = (\d+), CBS = (\d+);"
    ref = {m[0]: m[1:] for m in re.findall(pat, cu)}
    got = {m[0]: m[1:] for m in re.findall(pat, em)}
    assert got and all(ref[k] == v for k, v in got.items())


@pytest.mark.parametrize("degree,ttype,r,cells,upper,mask,coef", [
    (4, "CGP", 2, [3, 2, 2], [1.2, 0.8, 1.0], 0x3f, False),     # configs[1] family: 3 cells per CTA, 4 CTAs
    (4, "DG", 2, [4, 1, 2], [1.0, 1.0, 1.0], 0x00, True),       # nb = 3, no constraints, ragged last CTA
    (3, "DG", 1, [3, 3, 2], [1.0, 1.5, 0.5], 0x15, True),       # Dirichlet on the lower faces only
    (2, "CGP", 2, [2, 3, 2], [1.0, 2.0, 0.7], 0x2a, False),     # upper faces
    (3, "DG", 2, [2, 2, 2], [1.0, 1.0, 1.0], 0x3f, False),
])
def test_emulated_kernel_matches_oracle(emulator, tmp_path, degree, ttype, r, cells, upper, mask, coef):
    mesh = S.Mesh(3, cells, 0, [0.0, 0.0, 0.0], upper)
    space = S.Space(mesh, degree, dirichlet_faces=mask)
    A, B, _, _ = fth.get_fe_time_weights(ttype, r, 0.05, 1)
    nb = A.shape[0]
    K, M = S.MatrixFreeOperator(space, 0.0, 1.0), S.MatrixFreeOperator(space, 1.0, 0.0)
    cc = np.ones(mesh.n_cells)
    if coef:
        cc = 1.0 + (np.arange(mesh.n_cells) % 4) * 0.75
        K.laplace_coeff = np.repeat(cc[:, None], (degree + 1) ** 3, axis=1)
    src = np.stack([np.random.RandomState(42 + b).uniform(-1, 1, space.n_dofs) for b in range(nb)])
    ref = S.SystemMatrix(K, M, A, B).vmult(src)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    np.concatenate([A.reshape(-1), B.reshape(-1), cc, src.reshape(-1)]).astype(np.float64).tofile(fin)
    h = [upper[d] / cells[d] for d in range(3)]
    cmd = [emulator, str(degree), str(nb)] + [str(c) for c in cells] + ["%.17g" % v for v in h] + [hex(mask), fin, fout]
    run = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    out = np.fromfile(fout, dtype=np.float64).reshape(nb, space.n_dofs)
    assert np.abs(out - ref).max() <= 1e-13 * np.abs(ref).max(), np.abs(out - ref).max() / np.abs(ref).max()
    assert np.all(out[:, space.constrained] == 0)
    # transpose: the caller hands the kernel Alpha^T, Beta^T (capi_op.cu, d_alphaT / d_betaT)
    np.concatenate([A.T.reshape(-1), B.T.reshape(-1), cc, src.reshape(-1)]).astype(np.float64).tofile(fin)
    run = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0
    out_t = np.fromfile(fout, dtype=np.float64).reshape(nb, space.n_dofs)
    ref_t = S.SystemMatrix(K, M, A, B).Tvmult(src)
    assert np.abs(out_t - ref_t).max() <= 1e-13 * np.abs(ref_t).max()


def test_emulated_kernel_in_float_at_production_mesh_width(emulator, tmp_path):
    """FP32 (the precision of the multigrid levels) at h = 1/96, tau = 2^-6 — the mesh width of configs[1], where the
    mode eigenvalues reach 380 / h^2 = 3.5e6: the fast-diagonalisation form is as accurate as the direct form in float
    (3e-7) and far inside the 1e-5 bar of north_star."""
    degree, cells, h = 4, [4, 3, 2], 1.0 / 96
    mesh = S.Mesh(3, cells, 0, [0.0, 0.0, 0.0], [c * h for c in cells])
    space = S.Space(mesh, degree)
    A, B, _, _ = fth.get_fe_time_weights("CGP", 2, 2.0 ** -6, 1)
    src = np.stack([np.sin(3.0 * np.arange(space.n_dofs) / space.n_dofs + b) for b in range(2)])
    src[:, space.constrained] = 0
    ref = S.SystemMatrix(S.MatrixFreeOperator(space, 0.0, 1.0), S.MatrixFreeOperator(space, 1.0, 0.0), A, B).vmult(src)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    np.concatenate([A.reshape(-1), B.reshape(-1), np.ones(mesh.n_cells), src.reshape(-1)]).tofile(fin)
    err = {}
    for prec in ("f64", "f32"):
        cmd = [emulator, str(degree), "2"] + [str(c) for c in cells] + ["%.17g" % h] * 3 + ["0x3f", fin, fout] + ([prec] if prec == "f32" else [])
        run = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert run.returncode == 0, run.stdout + run.stderr
        out = np.fromfile(fout, dtype=np.float64).reshape(2, space.n_dofs)
        err[prec] = np.abs(out - ref).max() / np.abs(ref).max()
    assert err["f64"] < 1e-13 and err["f32"] < 2e-6, err
