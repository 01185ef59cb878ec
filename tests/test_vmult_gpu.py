"""Parity of the CUDA space-time operator (through the C ABI) with the CPU oracle.

Tolerances (BASELINE.json north_star): relative 1e-12 in FP64, 1e-5 in FP32 (multigrid levels).
The oracle applies the reference's UNFUSED algorithm (operators.h:536-559); the kernel is fused.
"""
import numpy as np
import pytest

from oracle import fe_time as ft
from oracle import spatial as S

pytestmark = pytest.mark.gpu

TOL = {0: 1e-12, 1: 1e-5}


def _setup(dim, k, ref, distort, subdivisions=None, lower=None, upper=None):
    sub = subdivisions or [1] * dim
    mesh = S.Mesh(dim, sub, ref, lower=lower, upper=upper, distort=distort)
    space = S.Space(mesh, k)
    return mesh, space


def _time_matrices(ttype, r, nts, tau=0.05):
    w = ft.get_fe_time_weights(ttype, r, tau, nts)
    return w[0], w[1], w[2], w[3]


def _rand_block(nb, n, seed=42):
    return np.stack([np.random.RandomState(seed + b).uniform(-1, 1, n) for b in range(nb)])


def _gpu_op(ctx, mesh, k, A, B, number_type=0, coeff=None, coeff_q=None, variant=0):
    import dealii_stfem_b200 as st
    gm = st.Mesh(ctx, mesh.n, lower=mesh.lower, upper=mesh.upper,
                 vertices=None if mesh.cartesian else mesh.vertices.reshape(-1, mesh.dim))
    return gm, st.Operator(gm, k, A, B, number_type=number_type, laplace_coeff_cell=coeff,
                           laplace_coeff_q=coeff_q, variant=variant)


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


CASES = [
    # dim, k, ref, distort, ttype, r, nts
    (2, 1, 2, 0.0, "DG", 0, 1),
    (2, 2, 3, 0.0, "DG", 1, 1),      # config 1 family (Q2 x DG(1))
    (2, 2, 2, 0.0, "DG", 1, 2),
    (2, 3, 2, 0.15, "CGP", 2, 1),
    (2, 5, 1, 0.1, "CGP", 4, 2),
    (2, 6, 1, 0.0, "DG", 2, 1),
    (3, 1, 2, 0.0, "DG", 1, 1),
    (3, 2, 1, 0.2, "DG", 1, 1),
    (3, 3, 1, 0.15, "DG", 2, 1),     # config 4 family (Q3 x DG(2), perturbed)
    (3, 4, 1, 0.0, "CGP", 2, 1),     # config 2 family (Q4 x cG(2))
    (3, 4, 1, 0.1, "DG", 1, 1),      # config 5 family (Q4 x DG(1))
    (3, 3, 1, 0.0, "DG", 3, 4),      # 16 blocks (tf05-style, 4 steps at once)
    (3, 5, 0, 0.1, "DG", 0, 1),
]


@pytest.mark.parametrize("number_type", [0, 1])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "d%d_k%d_r%d_p%g_%s%d_x%d" % c)
def test_vmult_matches_oracle(ctx, case, number_type):
    dim, k, ref, distort, ttype, r, nts = case
    mesh, space = _setup(dim, k, ref, distort)
    A, B, _, _ = _time_matrices(ttype, r, nts)
    nb = A.shape[0]
    dt = np.float64 if number_type == 0 else np.float32
    K = S.MatrixFreeOperator(space, 0.0, 1.0)
    M = S.MatrixFreeOperator(space, 1.0, 0.0)
    sysm = S.SystemMatrix(K, M, A, B)
    src = _rand_block(nb, space.n_dofs).astype(dt)
    ref_dst = sysm.vmult(src.astype(np.float64))
    gm, op = _gpu_op(ctx, mesh, k, A, B, number_type)
    d_src, d_dst = op.new_vector().upload(src), op.new_vector()
    op.vmult(d_dst, d_src)
    out = d_dst.download()
    assert out.dtype == dt
    assert _rel(out.astype(np.float64), ref_dst) < TOL[number_type]
    # constrained rows are exactly zero
    assert np.all(out[:, space.constrained] == 0)
    # Tvmult
    ref_t = sysm.Tvmult(src.astype(np.float64))
    op.Tvmult(d_dst, d_src)
    assert _rel(d_dst.download().astype(np.float64), ref_t) < TOL[number_type]
    for v in (d_src, d_dst):
        v.free()
    op.close()
    gm.close()


@pytest.mark.parametrize("dim,k", [(2, 2), (3, 3)])
def test_vmult_slice_add(ctx, dim, k):
    mesh, space = _setup(dim, k, 1, 0.1)
    _, _, G, Z = _time_matrices("CGP", 3, 1)
    K = S.MatrixFreeOperator(space, 0.0, 1.0)
    M = S.MatrixFreeOperator(space, 1.0, 0.0)
    sysm = S.SystemMatrix(K, M, G, Z)
    nb = G.shape[0]
    src0 = _rand_block(1, space.n_dofs, seed=7)
    dst0 = _rand_block(nb, space.n_dofs, seed=9)
    dst0[:, space.constrained] = 0
    ref_dst = sysm.vmult_slice_add(dst0.copy(), src0[0])
    gm, op = _gpu_op(ctx, mesh, k, G, Z)
    d_src = op.new_vector(1).upload(src0)
    d_dst = op.new_vector(nb).upload(dst0)
    op.vmult_slice_add(d_dst, d_src)
    assert _rel(d_dst.download(), ref_dst) < 1e-12
    d_src.free(); d_dst.free(); op.close(); gm.close()


def test_heterogeneous_coefficient_perturbed_mesh(ctx):
    """Config 4: [-1,1]^3, subdivisions 5, Coefficient<dim> on K only (tp_01.cc:118-119)."""
    dim, k = 3, 3
    mesh, space = _setup(dim, k, 0, 0.15, subdivisions=[5, 5, 5], lower=[-1, -1, -1], upper=[1, 1, 1])
    coef = S.Coefficient(dim, [5, 5, 5], [-1, -1, -1], [1, 1, 1], distort_coeff=0.6)
    A, B, _, _ = _time_matrices("DG", 2, 1)
    K = S.MatrixFreeOperator(space, 0.0, 1.0)
    K.evaluate_coefficient(coef)
    M = S.MatrixFreeOperator(space, 1.0, 0.0)
    cc = K.laplace_coeff        # Table [cell][q] like operators.h:1185-1186
    sysm = S.SystemMatrix(K, M, A, B)
    src = _rand_block(3, space.n_dofs)
    ref_dst = sysm.vmult(src)
    gm, op = _gpu_op(ctx, mesh, k, A, B, coeff_q=cc)
    d_src, d_dst = op.new_vector().upload(src), op.new_vector()
    op.vmult(d_dst, d_src)
    assert _rel(d_dst.download(), ref_dst) < 1e-12
    d_src.free(); d_dst.free(); op.close(); gm.close()


def test_linearity_and_symmetry_large(ctx):
    """Size-independent properties at a size the oracle does not reach: A(ax+by) = aAx + bAy and
    <Ax, y> = <x, A^T y> on a 48^3 Q4 mesh (1.4e7 space-time DoFs)."""
    import dealii_stfem_b200 as st
    A, B, _, _ = _time_matrices("CGP", 2, 1)
    gm = st.Mesh(ctx, [48, 48, 48])
    op = st.Operator(gm, 4, A, B)
    n, nb = op.n, 2
    x, y = _rand_block(nb, n, 1), _rand_block(nb, n, 5)
    dx, dy, dz, dr = (op.new_vector() for _ in range(4))
    dx.upload(x); dy.upload(y); dz.upload(2.0 * x - 3.0 * y)
    op.vmult(dr, dx); Ax = dr.download()
    op.vmult(dr, dy); Ay = dr.download()
    op.vmult(dr, dz); Az = dr.download()
    assert _rel(Az, 2.0 * Ax - 3.0 * Ay) < 1e-12
    op.Tvmult(dr, dy); ATy = dr.download()
    lhs, rhs = np.vdot(Ax, y), np.vdot(x, ATy)   # y, x include boundary entries: rows/cols are zero there
    assert abs(lhs - rhs) <= 1e-11 * max(abs(lhs), np.linalg.norm(Ax) * np.linalg.norm(y) * 1e-3)
    for v in (dx, dy, dz, dr):
        v.free()
    op.close(); gm.close()


def test_per_cell_coefficient_cartesian(ctx):
    """Coefficient<dim> is constant per cell on the unperturbed practical mesh: the per-cell path."""
    dim, k = 3, 2
    mesh, space = _setup(dim, k, 1, 0.0, subdivisions=[5, 5, 5], lower=[-1, -1, -1], upper=[1, 1, 1])
    coef = S.Coefficient(dim, [5, 5, 5], [-1, -1, -1], [1, 1, 1], distort_coeff=0.5)
    A, B, _, _ = _time_matrices("DG", 1, 2)
    K = S.MatrixFreeOperator(space, 0.0, 1.0)
    K.evaluate_coefficient(coef)
    M = S.MatrixFreeOperator(space, 1.0, 0.0)
    cc = K.laplace_coeff
    assert np.all(cc == cc[:, :1])
    sysm = S.SystemMatrix(K, M, A, B)
    src = _rand_block(4, space.n_dofs)
    ref_dst = sysm.vmult(src)
    for kw in (dict(coeff=cc[:, 0]), dict(coeff_q=cc)):
        gm, op = _gpu_op(ctx, mesh, k, A, B, **kw)
        d_src, d_dst = op.new_vector().upload(src), op.new_vector()
        op.vmult(d_dst, d_src)
        assert _rel(d_dst.download(), ref_dst) < 1e-12
        d_src.free(); d_dst.free(); op.close(); gm.close()


# ----------------------------------------------------------------------------------------------------------------
# Cartesian 3D fast path (csrc/st_vmult_cart.cuh): nodal 1D-matrix form of the same operator.  Checked against the
# oracle AND against the generic q-point kernel (variant 1) on anisotropic boxes, partial Dirichlet masks, all
# supported degrees, rectangular time matrices and per-cell coefficients.
CART_CASES = [
    # k, subdivisions, upper, ttype, r, nts, dirichlet mask
    (1, [3, 2, 2], [1.5, 1.0, 0.5], "DG", 1, 1, 0x3f),
    (2, [2, 3, 1], [1.0, 2.0, 0.7], "CGP", 2, 1, 0x3f),
    (3, [2, 2, 3], [1.0, 1.0, 1.0], "DG", 2, 1, 0x15),       # Dirichlet on the three lower faces only
    (4, [3, 2, 2], [1.2, 0.8, 1.0], "CGP", 2, 1, 0x3f),      # config 2 family
    (4, [2, 2, 2], [1.0, 1.0, 1.0], "DG", 1, 2, 0x2a),       # 4 blocks, upper faces
    (4, [5, 1, 1], [1.0, 1.0, 1.0], "DG", 2, 1, 0x00),       # no constraints, nb = 3
    (5, [2, 1, 2], [1.0, 0.5, 1.0], "DG", 0, 1, 0x3f),
]


@pytest.mark.parametrize("number_type", [0, 1])
@pytest.mark.parametrize("case", CART_CASES, ids=lambda c: "k%d_%s_%s%d_x%d_m%x" % (c[0], "x".join(map(str, c[1])), c[3], c[4], c[5], c[6]))
def test_cartesian_kernel(ctx, case, number_type):
    import dealii_stfem_b200 as st
    k, sub, upper, ttype, r, nts, mask = case
    mesh = S.Mesh(3, sub, 0, lower=[0, 0, 0], upper=upper)
    space = S.Space(mesh, k, dirichlet_faces=mask)
    A, B, _, _ = _time_matrices(ttype, r, nts)
    nb = A.shape[0]
    dt = np.float64 if number_type == 0 else np.float32
    sysm = S.SystemMatrix(S.MatrixFreeOperator(space, 0.0, 1.0), S.MatrixFreeOperator(space, 1.0, 0.0), A, B)
    src = _rand_block(nb, space.n_dofs).astype(dt)
    ref_dst = sysm.vmult(src.astype(np.float64))
    ref_t = sysm.Tvmult(src.astype(np.float64))
    gm = st.Mesh(ctx, mesh.n, lower=mesh.lower, upper=mesh.upper, dirichlet_faces=mask)
    outs = {}
    for variant in (0, 1):
        op = st.Operator(gm, k, A, B, number_type=number_type, variant=variant)
        d_src, d_dst = op.new_vector().upload(src), op.new_vector()
        op.vmult(d_dst, d_src)
        out = d_dst.download()
        assert _rel(out.astype(np.float64), ref_dst) < TOL[number_type], "variant %d" % variant
        assert np.all(out[:, space.constrained] == 0)
        op.Tvmult(d_dst, d_src)
        assert _rel(d_dst.download().astype(np.float64), ref_t) < TOL[number_type], "variant %d (T)" % variant
        outs[variant] = out
        d_src.free(); d_dst.free(); op.close()
    assert _rel(outs[0].astype(np.float64), outs[1].astype(np.float64)) < TOL[number_type]
    gm.close()


def test_cartesian_kernel_slice_and_cell_coefficient(ctx):
    import dealii_stfem_b200 as st
    k = 4
    mesh = S.Mesh(3, [5, 5, 5], 0, lower=[-1, -1, -1], upper=[1, 1, 1])
    space = S.Space(mesh, k)
    coef = S.Coefficient(3, [5, 5, 5], [-1, -1, -1], [1, 1, 1], distort_coeff=0.5)
    K = S.MatrixFreeOperator(space, 0.0, 1.0)
    K.evaluate_coefficient(coef)
    M = S.MatrixFreeOperator(space, 1.0, 0.0)
    cc = K.laplace_coeff[:, 0].copy()
    gm = st.Mesh(ctx, mesh.n, lower=mesh.lower, upper=mesh.upper)
    # square system with a per-cell coefficient
    A, B, G, Z = _time_matrices("CGP", 3, 1)
    sysm = S.SystemMatrix(K, M, A, B)
    src = _rand_block(A.shape[0], space.n_dofs)
    op = st.Operator(gm, k, A, B, laplace_coeff_cell=cc)
    d_src, d_dst = op.new_vector().upload(src), op.new_vector()
    op.vmult(d_dst, d_src)
    assert _rel(d_dst.download(), sysm.vmult(src)) < 1e-12
    d_src.free(); d_dst.free(); op.close()
    # nb x 1 slice operator accumulating into dst (operators.h:586-611)
    sl = S.SystemMatrix(K, M, G, Z)
    nb = G.shape[0]
    src0 = _rand_block(1, space.n_dofs, seed=7)
    dst0 = _rand_block(nb, space.n_dofs, seed=9)
    dst0[:, space.constrained] = 0
    ref_dst = sl.vmult_slice_add(dst0.copy(), src0[0])
    op = st.Operator(gm, k, G, Z, laplace_coeff_cell=cc)
    d_src = op.new_vector(1).upload(src0)
    d_dst = op.new_vector(nb).upload(dst0)
    op.vmult_slice_add(d_dst, d_src)
    assert _rel(d_dst.download(), ref_dst) < 1e-12
    d_src.free(); d_dst.free(); op.close(); gm.close()


@pytest.mark.parametrize("cells,k", [([6, 5, 10], 3), ([4, 4, 37], 2), ([3, 3, 3], 4)])
def test_host_buffer_entry_point_pipeline(ctx, cells, k):
    """stfem_op_vmult_host (the e2e path of bench.py): upload / cell kernel / download pipelined over z slabs must give
    the device-resident result, incl. slab counts that do not divide the mesh and the unpipelined small-mesh path."""
    import dealii_stfem_b200 as st
    mesh = S.Mesh(3, cells, 0)
    space = S.Space(mesh, k)
    A, B, _, _ = _time_matrices("CGP", 2, 1)
    sysm = S.SystemMatrix(S.MatrixFreeOperator(space, 0.0, 1.0), S.MatrixFreeOperator(space, 1.0, 0.0), A, B)
    src = _rand_block(2, space.n_dofs)
    gm = st.Mesh(ctx, mesh.n)
    op = st.Operator(gm, k, A, B)
    out = np.full_like(src, np.nan)
    for transpose, ref in ((False, sysm.vmult(src)), (True, sysm.Tvmult(src))):
        op.vmult_host(out, src, transpose=transpose)
        assert _rel(out, ref) < 1e-12
        assert np.all(out[:, space.constrained] == 0)
    op.close(); gm.close()


@pytest.mark.parametrize("number_type", [0, 1])
@pytest.mark.parametrize("k,sub,ttype,r,nts,mask", [(1, [3, 2, 2], "DG", 1, 1, 0x3f), (2, [2, 3, 2], "CGP", 2, 1, 0x3f),
                                                    (3, [3, 3, 2], "DG", 2, 1, 0x15), (4, [2, 2, 3], "CGP", 2, 1, 0x3f),
                                                    (4, [3, 1, 2], "DG", 1, 2, 0x00)])
def test_general_geometry_plane_kernel(ctx, k, sub, ttype, r, nts, mask, number_type):
    """csrc/st_vmult_plane.cuh (perturbed MappingQ1 cells + per-q coefficient) against the oracle and against the
    generic q-point kernel (variant 1), incl. Tvmult, partial Dirichlet masks and rectangular slice operators."""
    import dealii_stfem_b200 as st
    mesh = S.Mesh(3, sub, 0, lower=[0, 0, 0], upper=[1.0, 1.2, 0.8], distort=0.15)
    space = S.Space(mesh, k, dirichlet_faces=mask)
    A, B, G, Z = _time_matrices(ttype, r, nts)
    nb = A.shape[0]
    dt = np.float64 if number_type == 0 else np.float32
    Kop = S.MatrixFreeOperator(space, 0.0, 1.0)
    rng = np.random.RandomState(11)
    cq = rng.uniform(0.5, 2.0, (mesh.n_cells, (k + 1) ** 3))
    Kop.laplace_coeff = cq
    Mop = S.MatrixFreeOperator(space, 1.0, 0.0)
    sysm = S.SystemMatrix(Kop, Mop, A, B)
    src = _rand_block(nb, space.n_dofs).astype(dt)
    ref_dst, ref_t = sysm.vmult(src.astype(np.float64)), sysm.Tvmult(src.astype(np.float64))
    gm = st.Mesh(ctx, mesh.n, lower=mesh.lower, upper=mesh.upper, vertices=mesh.vertices.reshape(-1, 3), dirichlet_faces=mask)
    outs = {}
    for variant in (0, 1):
        op = st.Operator(gm, k, A, B, number_type=number_type, laplace_coeff_q=cq, variant=variant)
        d_src, d_dst = op.new_vector().upload(src), op.new_vector()
        op.vmult(d_dst, d_src)
        outs[variant] = d_dst.download()
        assert _rel(outs[variant].astype(np.float64), ref_dst) < TOL[number_type], "variant %d" % variant
        assert np.all(outs[variant][:, space.constrained] == 0)
        op.Tvmult(d_dst, d_src)
        assert _rel(d_dst.download().astype(np.float64), ref_t) < TOL[number_type], "variant %d (T)" % variant
        d_src.free(); d_dst.free(); op.close()
    assert _rel(outs[0].astype(np.float64), outs[1].astype(np.float64)) < TOL[number_type]
    if number_type == 0:
        sl = S.SystemMatrix(Kop, Mop, G, Z)
        src0 = _rand_block(1, space.n_dofs, seed=7)
        dst0 = _rand_block(G.shape[0], space.n_dofs, seed=9)
        dst0[:, space.constrained] = 0
        ref_sl = sl.vmult_slice_add(dst0.copy(), src0[0])
        op = st.Operator(gm, k, G, Z, laplace_coeff_q=cq)
        d_src, d_dst = op.new_vector(1).upload(src0), op.new_vector(G.shape[0]).upload(dst0)
        op.vmult_slice_add(d_dst, d_src)
        assert _rel(d_dst.download(), ref_sl) < 1e-12
        d_src.free(); d_dst.free(); op.close()
    gm.close()


@pytest.mark.parametrize("number_type", [0, 1])
@pytest.mark.parametrize("k,sub,ttype,r,mask,coef", [(1, [3, 2, 2], "DG", 1, 0x3f, False), (2, [2, 3, 2], "CGP", 2, 0x3f, True),
                                                     (3, [3, 3, 2], "DG", 2, 0x15, True), (4, [2, 2, 3], "CGP", 2, 0x3f, False),
                                                     (4, [4, 3, 2], "DG", 1, 0x00, True)])
def test_general_geometry_on_the_fly(ctx, k, sub, ttype, r, mask, coef, number_type):
    """Perturbed MappingQ1 cells with the geometry computed on the fly from the cell vertices (kernel_variant 6, and the
    default when the stored metric would not fit; csrc/st_vmult_plane.cuh, OTF) against the oracle (which applies the reference's stored J^-1 / JxW
    algorithm, include/operators.h:1135-1173) and against the precomputed-metric kernel (variant 5) and the generic q-point
    kernel (variant 1); optional per-cell coefficient (Coefficient<dim> constant per cell)."""
    import dealii_stfem_b200 as st
    mesh = S.Mesh(3, sub, 0, lower=[0, 0, 0], upper=[1.0, 1.2, 0.8], distort=0.2)
    space = S.Space(mesh, k, dirichlet_faces=mask)
    A, B, _, _ = _time_matrices(ttype, r, 1)
    nb = A.shape[0]
    dt = np.float64 if number_type == 0 else np.float32
    Kop = S.MatrixFreeOperator(space, 0.0, 1.0)
    cc = None
    if coef:
        cc = 0.5 + (np.arange(mesh.n_cells) % 5) * 0.4
        Kop.laplace_coeff = np.repeat(cc[:, None], (k + 1) ** 3, axis=1)
    Mop = S.MatrixFreeOperator(space, 1.0, 0.0)
    sysm = S.SystemMatrix(Kop, Mop, A, B)
    src = _rand_block(nb, space.n_dofs).astype(dt)
    ref_dst, ref_t = sysm.vmult(src.astype(np.float64)), sysm.Tvmult(src.astype(np.float64))
    gm = st.Mesh(ctx, mesh.n, lower=mesh.lower, upper=mesh.upper, vertices=mesh.vertices.reshape(-1, 3), dirichlet_faces=mask)
    outs = {}
    for variant in (6, 5, 1):
        op = st.Operator(gm, k, A, B, number_type=number_type, laplace_coeff_cell=cc, variant=variant)
        d_src, d_dst = op.new_vector().upload(src), op.new_vector()
        op.vmult(d_dst, d_src)
        outs[variant] = d_dst.download()
        assert _rel(outs[variant].astype(np.float64), ref_dst) < TOL[number_type], "variant %d" % variant
        assert np.all(outs[variant][:, space.constrained] == 0)
        op.Tvmult(d_dst, d_src)
        assert _rel(d_dst.download().astype(np.float64), ref_t) < TOL[number_type], "variant %d (T)" % variant
        d_src.free(); d_dst.free(); op.close()
    assert _rel(outs[6].astype(np.float64), outs[5].astype(np.float64)) < TOL[number_type]
    gm.close()


@pytest.mark.parametrize("number_type", [0, 1])
@pytest.mark.parametrize("cells,variant", [([12, 2, 3], 40), ([24, 2, 2], 41), ([8, 3, 2], 42), ([12, 3, 2], 20)])
def test_cartesian_kernel_tma_and_pipelined_variants(ctx, cells, variant, number_type):
    """Tuning variants of the Cartesian Q4 kernel (cp.async.bulk + mbarrier gather: 40-42; register-prefetch
    persistent kernel: 20) must reproduce the oracle like the default one (odd row lengths: every second row of the
    block vectors is only 8-byte aligned, which the bulk-copy path has to handle)."""
    import dealii_stfem_b200 as st
    k = 4
    mesh = S.Mesh(3, cells, 0, lower=[0, 0, 0], upper=[1.0, 0.5, 0.75])
    space = S.Space(mesh, k, dirichlet_faces=0x1b)
    A, B, _, _ = _time_matrices("CGP", 2, 1)
    dt = np.float64 if number_type == 0 else np.float32
    sysm = S.SystemMatrix(S.MatrixFreeOperator(space, 0.0, 1.0), S.MatrixFreeOperator(space, 1.0, 0.0), A, B)
    src = _rand_block(2, space.n_dofs).astype(dt)
    ref_dst = sysm.vmult(src.astype(np.float64))
    gm = st.Mesh(ctx, mesh.n, lower=mesh.lower, upper=mesh.upper, dirichlet_faces=0x1b)
    op = st.Operator(gm, k, A, B, number_type=number_type, variant=variant)
    d_src, d_dst = op.new_vector().upload(src), op.new_vector()
    op.vmult(d_dst, d_src)
    out = d_dst.download()
    assert _rel(out.astype(np.float64), ref_dst) < TOL[number_type]
    assert np.all(out[:, space.constrained] == 0)
    d_src.free(); d_dst.free(); op.close(); gm.close()
