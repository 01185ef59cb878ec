"""The C++/OpenMP CPU restatement (oracle/cpu_ref.cpp, used as cpu_baseline) against the numpy oracle,
and the numpy operator against an assembled matrix (the mf == sparse check of the reference's
tests/tp_05dgp_support.cc:140-149).  CPU only."""
import numpy as np
import pytest

from oracle import cpu_ref, fe_time as ft, spatial as S


@pytest.mark.parametrize("dim,k,ref,dist", [(2, 1, 3, 0.0), (2, 2, 3, 0.0), (2, 4, 1, 0.2), (3, 1, 2, 0.0),
                                            (3, 2, 1, 0.2), (3, 3, 1, 0.1), (3, 4, 1, 0.0)])
def test_cpu_ref_matches_numpy_oracle(dim, k, ref, dist):
    mesh = S.Mesh(dim, [1] * dim, ref, distort=dist)
    sp = S.Space(mesh, k)
    A, B, _, _ = ft.get_fe_time_weights("DG", 1, 0.05, 2)
    sysm = S.SystemMatrix(S.MatrixFreeOperator(sp, 0, 1), S.MatrixFreeOperator(sp, 1, 0), A, B)
    src = np.random.RandomState(1).uniform(-1, 1, (4, sp.n_dofs))
    r, c = sysm.vmult(src), cpu_ref.system_vmult(sp, A, B, src)
    assert np.abs(r - c).max() <= 1e-13 * np.abs(r).max()
    rt, ct = sysm.Tvmult(src), cpu_ref.system_vmult(sp, A, B, src, transpose=True)
    assert np.abs(rt - ct).max() <= 1e-13 * np.abs(rt).max()


@pytest.mark.parametrize("dim,k,dist", [(2, 2, 0.0), (2, 3, 0.15), (3, 2, 0.0), (3, 2, 0.2)])
def test_matrix_free_equals_assembled(dim, k, dist):
    mesh = S.Mesh(dim, [1] * dim, 2 if dim == 2 else 1, distort=dist)
    sp = S.Space(mesh, k)
    u = np.random.RandomState(0).uniform(-1, 1, sp.n_dofs)
    um = np.where(sp.constrained, 0, u)
    for op in (S.MatrixFreeOperator(sp, 0.0, 1.0), S.MatrixFreeOperator(sp, 1.0, 0.0)):
        Aa = op.compute_system_matrix()
        r1, r2 = op.vmult(u), Aa @ um
        r2[sp.constrained] = 0
        assert np.abs(r1 - r2).max() <= 1e-13 * np.abs(r1).max()
        assert abs(Aa - Aa.T).max() <= 1e-13 * abs(Aa).max()
        # constrained rows: positive diagonal only
        d = Aa.diagonal()
        assert np.all(d[sp.constrained] > 0)


def test_coefficient_table_uses_mt19937_default_seed():
    """boost::mt19937(default_seed = 5489), one 32-bit draw per value (operators.h:915-921)."""
    c = S.Coefficient(2, [2, 2], [0, 0], [1, 1], distort_coeff=0.5)
    first = 0.5 + 1.0 * (3499211612 / 4294967296.0)          # first genrand_int32 of MT19937(5489)
    assert abs(c.table.reshape(-1)[0] - first) < 1e-15
    pts = np.array([[0.1, 0.1], [0.1, 0.5], [0.7, 0.5]])
    base = np.array([1.0, 9.0, 16.0])
    idx = [(0, 0), (0, 1), (1, 1)]
    assert np.allclose(c(pts), base * np.array([c.table[i] for i in idx]))
