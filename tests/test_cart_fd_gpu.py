"""Parity of kernel_variant 60 (per-cell Cartesian operator in fast-diagonalisation form, csrc/st_vmult_cart_fd.cuh) with the
oracle.  Written at the end of round 1, first run on a B200 in round 2 (12 cases green, 1.04 ms against 1.13 ms of the round-1
default on configs[1]; profiles/r02_variant60.md).  It serves operators with a per-cell coefficient on Cartesian meshes when
selected; the constant-coefficient default is the brick kernel.  The algebra is also verified on the CPU in
tests/test_cart_fd_modes.py."""
import numpy as np
import pytest

from oracle import spatial as S

pytestmark = pytest.mark.gpu

TOL = {0: 1e-12, 1: 1e-5}
CASES = [
    # k, cells, upper, ttype, r, nts, dirichlet mask, per-cell coefficient
    (2, [2, 3, 1], [1.0, 2.0, 0.7], "CGP", 2, 1, 0x3f, False),
    (3, [2, 2, 3], [1.0, 1.0, 1.0], "DG", 2, 1, 0x15, False),      # nb = 3, Dirichlet on the lower faces only
    (3, [4, 3, 2], [1.0, 1.0, 1.0], "DG", 1, 1, 0x3f, True),
    (4, [3, 2, 2], [1.2, 0.8, 1.0], "CGP", 2, 1, 0x3f, False),     # configs[1] family
    (4, [5, 1, 1], [1.0, 1.0, 1.0], "DG", 2, 1, 0x00, False),      # no constraints, nb = 3
    (4, [7, 5, 3], [1.0, 1.0, 1.0], "CGP", 2, 1, 0x3f, True),      # several CTAs, ragged last CTA
]


@pytest.mark.parametrize("number_type", [0, 1])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "k%d_%s_%s%d_m%x_c%d" % (c[0], "x".join(map(str, c[1])), c[3], c[4], c[6], c[7]))
def test_cartesian_fd_kernel(ctx, case, number_type):
    import dealii_stfem_b200 as st
    from dealii_stfem_b200 import fe_time_host as fth
    k, cells, upper, ttype, r, nts, mask, coef = case
    mesh = S.Mesh(3, cells, 0, lower=[0, 0, 0], upper=upper)
    space = S.Space(mesh, k, dirichlet_faces=mask)
    A, B, _, _ = fth.get_fe_time_weights(ttype, r, 0.05, nts)
    nb = A.shape[0]
    K, M = S.MatrixFreeOperator(space, 0.0, 1.0), S.MatrixFreeOperator(space, 1.0, 0.0)
    kw = {}
    if coef:
        cc = 1.0 + (np.arange(mesh.n_cells) % 4) * 0.75
        K.laplace_coeff = np.repeat(cc[:, None], (k + 1) ** 3, axis=1)
        kw = {"laplace_coeff_cell": cc}
    sysm = S.SystemMatrix(K, M, A, B)
    dt = np.float64 if number_type == 0 else np.float32
    src = np.stack([np.random.RandomState(42 + b).uniform(-1, 1, space.n_dofs) for b in range(nb)]).astype(dt)
    ref, ref_t = sysm.vmult(src.astype(np.float64)), sysm.Tvmult(src.astype(np.float64))
    gm = st.Mesh(ctx, mesh.n, lower=mesh.lower, upper=mesh.upper, dirichlet_faces=mask)
    outs = {}
    for variant in (0, 60):
        op = st.Operator(gm, k, A, B, number_type=number_type, variant=variant, **kw)
        d_src, d_dst = op.new_vector().upload(src), op.new_vector()
        op.vmult(d_dst, d_src)
        out = d_dst.download().astype(np.float64)
        assert np.abs(out - ref).max() <= TOL[number_type] * np.abs(ref).max(), "variant %d" % variant
        assert np.all(out[:, space.constrained] == 0)
        op.Tvmult(d_dst, d_src)
        assert np.abs(d_dst.download().astype(np.float64) - ref_t).max() <= TOL[number_type] * np.abs(ref_t).max(), "variant %d (T)" % variant
        outs[variant] = out
        d_src.free(); d_dst.free(); op.close()
    assert np.abs(outs[60] - outs[0]).max() <= TOL[number_type] * np.abs(ref).max()
    gm.close()
