"""kernel_variant 60 (EXPERIMENTAL fast-diagonalisation form of the Cartesian operator, csrc/st_vmult_cart_fd.cuh): the
host-side modes (stfem_cart_fd_modes) and the algebra the kernel implements, emulated in numpy cell by cell and compared
with the oracle's SystemMatrix::vmult.  CPU only; the CUDA kernel itself is exercised by tests/test_next_round_gpu.py."""
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
