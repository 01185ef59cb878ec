"""Host-side problem data of the practical runs (csrc/capi_problem.cu) against the oracle's restatement of
Coefficient<dim> (reference include/operators.h:870-965, 1060-1087) and of the cut-off initial value
(tests/tp_01.cc:374-381).  CPU only: these entry points do no GPU work."""
import math

import numpy as np
import pytest
from scipy.integrate import quad

from dealii_stfem_b200 import problem_host as ph
from oracle import spatial as S
from oracle import tp_01


def test_distortion_table_is_mt19937_default_seed_one_draw_per_value():
    """boost::mt19937(default_seed = 5489) through boost::uniform_real_distribution: u / 2^32 * (b - a) + a."""
    t = ph.coefficient_distortion([2, 3], 0.5)
    assert t.shape == (2, 3)
    # the first outputs of MT19937 seeded with 5489 (the generator's published known answers)
    first = [3499211612, 581869302, 3890346734, 3586334585, 545404204, 4161255391]
    assert np.array_equal(t.reshape(-1), np.array([u / 4294967296.0 * 1.0 + 0.5 for u in first]))
    c = S.Coefficient(3, [5, 5, 5], [-1, -1, -1], [1, 1, 1], distort_coeff=0.6)
    assert np.array_equal(ph.coefficient_distortion([5, 5, 5], 0.6), c.table)


@pytest.mark.parametrize("dim,sub,ref,degree,distort_grid,dc", [
    (3, [5, 5, 5], 1, 3, 0.15, 0.6),      # BASELINE configs[3]: perturbed mesh, distorted coefficient
    (3, [5, 5, 5], 0, 2, 0.0, 0.5),       # tests/json/practical01.json (Cartesian)
    (2, [5, 5], 2, 2, 0.1, 0.5),
    (2, [3, 2], 1, 4, 0.0, 0.0),          # no distortion: the three-valued coefficient alone
])
def test_coefficient_at_qpoints_matches_oracle(dim, sub, ref, degree, distort_grid, dc):
    lo, up = [-1.0] * dim, [1.0] * dim
    mesh = S.Mesh(dim, sub, ref, lo, up, distort=distort_grid)
    space = S.Space(mesh, degree)
    coef = S.Coefficient(dim, sub, lo, up, distort_coeff=dc)
    _, _, pts = space.geometry()
    want = coef(pts)
    got = ph.coefficient_at_qpoints(mesh.n, lo, up, degree, sub, lo, up, dc,
                                    vertices=mesh.vertices.reshape(-1, dim) if distort_grid else None)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    if dc == 0.0:
        assert set(np.unique(got)) <= {1.0, 9.0, 16.0} and len(np.unique(got)) == 3


def test_cutoff_unit_integrals_rederived_by_quadrature():
    """The three constants (deal.II integral_Cinfty) are the unit-ball integrals of e exp(-1/(1-r^2))."""
    f = lambda r: math.exp(1 - 1 / (1 - r * r)) if r < 1 else 0.0
    i1 = 2 * quad(f, 0, 1, epsabs=1e-14)[0]
    i2 = 2 * math.pi * quad(lambda r: r * f(r), 0, 1, epsabs=1e-14)[0]
    i3 = 4 * math.pi * quad(lambda r: r * r * f(r), 0, 1, epsabs=1e-14)[0]
    assert np.allclose([i1, i2, i3], tp_01.INTEGRAL_CINFTY, rtol=1e-12)


@pytest.mark.parametrize("dim,sub,ref,degree,distort_grid,radius", [
    (3, [5, 5, 5], 1, 3, 0.15, 1.0e-2),
    (3, [5, 5, 5], 1, 3, 0.0, 1.0e-2),
    (2, [5, 5], 1, 2, 0.1, 0.3),          # a radius that covers several cells
    (2, [2, 2], 2, 4, 0.0, 0.4),
])
def test_cutoff_interpolation_matches_oracle(dim, sub, ref, degree, distort_grid, radius):
    lo, up = [-1.0] * dim, [1.0] * dim
    mesh = S.Mesh(dim, sub, ref, lo, up, distort=distort_grid)
    if radius > 0.1:
        center = [0.05] * dim
    else:
        # a bump of radius 1e-2 only sees the mesh vertex it sits on (the reference's sourcePoint 0,0,0 is a vertex of
        # the undistorted practical mesh): centre it on the (possibly displaced) vertex next to the origin
        V = mesh.vertices.reshape(-1, dim)
        center = list(V[np.argmin(np.sum(V * V, axis=1))])
    space = S.Space(mesh, degree)
    want = tp_01.interpolate(space, lambda pts: tp_01.cutoff_cinfty(pts, center, radius))
    got = ph.cutoff_cinfty_interpolate(mesh.n, lo, up, degree, center, radius=radius,
                                       vertices=mesh.vertices.reshape(-1, dim) if distort_grid else None)
    assert got.shape == want.shape
    assert np.count_nonzero(want) >= 1
    assert np.allclose(got, want, rtol=1e-13, atol=0.0)
    if radius > 0.1 and distort_grid == 0.0:
        # integrate_to_one: the nodal interpolant integrates to about 1 (weights = integrals of the basis functions)
        w = tp_01.integrate_rhs(space, lambda pts: np.ones(pts.shape[:-1]))
        assert abs(w @ got - 1.0) < 0.05


def test_driver_level_coefficient_selects_per_cell_or_per_q_path():
    """HeatWaveProblem._laplace_coefficient (host logic of the product driver): on the practical Cartesian mesh the
    coefficient is constant per cell on every level (jumps at x, y = 0.2 are cell boundaries of the 5-subdivision box);
    on a perturbed mesh the per-q table is handed over.  Values = the oracle's evaluate_coefficient."""
    import types

    import dealii_stfem_b200 as st
    p = st.parse_parameters({"spaceTimeConvergenceTest": "false", "hyperRectLowerLeft": "-1,-1,-1", "hyperRectUpperRight": "1,1,1",
                             "subdivisions": "5,5,5", "distortCoeff": "0.5"}, 3)
    lo, up = p["hyperRectLowerLeft"], p["hyperRectUpperRight"]
    coef = S.Coefficient(3, p["subdivisions"], lo, up, distort_coeff=0.5)
    for rf, degree in ((0, 2), (1, 2), (1, 1)):
        mesh = S.Mesh(3, p["subdivisions"], rf, lo, up)
        fake = types.SimpleNamespace(p=p, _coeff_cache={}, mesh_desc={rf: (mesh.n, lo, up, None)})
        kw = st.HeatWaveProblem._laplace_coefficient(fake, rf, degree)
        assert list(kw) == ["laplace_coeff_cell"]
        _, _, pts = S.Space(mesh, degree).geometry()
        assert np.array_equal(kw["laplace_coeff_cell"], coef(pts)[:, 0])
        assert st.HeatWaveProblem._laplace_coefficient(fake, rf, degree) is kw          # cached
    mesh = S.Mesh(3, p["subdivisions"], 1, lo, up, distort=0.15)
    fake = types.SimpleNamespace(p=p, _coeff_cache={}, mesh_desc={1: (mesh.n, lo, up, mesh.vertices.reshape(-1, 3))})
    kw = st.HeatWaveProblem._laplace_coefficient(fake, 1, 2)
    assert list(kw) == ["laplace_coeff_q"]
    _, _, pts = S.Space(mesh, 2).geometry()
    assert np.array_equal(kw["laplace_coeff_q"], coef(pts))
    # convergence tests carry no coefficient
    p2 = st.parse_parameters({}, 3)
    assert st.HeatWaveProblem._laplace_coefficient(types.SimpleNamespace(p=p2), 0, 2) == {}


@pytest.mark.parametrize("dim,distort", [(2, 0.0), (2, 0.15), (3, 0.0), (3, 0.15)])
def test_oracle_point_evaluation_reproduces_linear_functions(dim, distort):
    """The oracle's FEPointEvaluation restatement (cell search + Newton inversion of the MappingQ1 map): a linear function
    of x lies in the mapped FE_Q(k) space, so its point values are exact on perturbed meshes too."""
    lo, up = [-1.0] * dim, [1.0] * dim
    mesh = S.Mesh(dim, [5] * dim, 1, lo, up, distort=distort)
    space = S.Space(mesh, 2)
    coef = np.arange(1, dim + 1, dtype=float)
    u = tp_01.interpolate(space, lambda pts: 0.3 + pts @ coef)[None, :]
    pts = [[0.75, 0.0], [0.013, -0.48]] if dim == 2 else [[0.75, 0.0, 0.0], [0.0, 0.0, 0.75], [0.75, 0.1, 0.75], [0.31, -0.77, 0.05]]
    got = tp_01.point_evaluate(space, pts, u)[0]
    assert np.allclose(got, 0.3 + np.asarray(pts) @ coef, rtol=0, atol=1e-13)


@pytest.mark.parametrize("ttype,r", [("CGP", 1), ("CGP", 3), ("DG", 0), ("DG", 2)])
def test_time_evaluation_matrix_matches_oracle(ttype, r):
    from dealii_stfem_b200 import fe_time_host as fth
    samples = (r + 1) ** 2 if r > 0 else 2
    M = fth.get_time_evaluation_matrix(ttype, r, samples)
    assert M.shape == (samples, r + 1)
    assert np.allclose(M, tp_01.time_evaluation_matrix(ttype, r, samples), rtol=0, atol=1e-13)
    assert np.allclose(M.sum(axis=1), 1.0)          # partition of unity


@pytest.mark.parametrize("grid", [[2, 1, 1], [2, 2, 1], [2, 2, 2]])
def test_coefficient_and_initial_value_on_partition_bricks(grid):
    """Multi-GPU runs (one brick of the box partition per rank, stfem_partition_brick): evaluating the coefficient table and
    the cut-off initial value on a rank's brick (local box, GLOBAL coefficient grid / source point) gives exactly the
    slice of the global arrays that the brick covers — what HeatWaveProblem does for partition != None."""
    from dealii_stfem_b200 import dist
    dim, degree, sub, ref, dc = 3, 2, [5, 5, 5], 1, 0.6
    lo, up = [-1.0] * 3, [1.0] * 3
    n = [s << ref for s in sub]
    nq = (degree + 1) ** 3
    glob = ph.coefficient_at_qpoints(n, lo, up, degree, sub, lo, up, dc).reshape(n[2], n[1], n[0], nq)
    center = [0.0, 0.0, 0.0]
    u0 = ph.cutoff_cinfty_interpolate(n, lo, up, degree, center, radius=0.3).reshape([degree * m + 1 for m in n][::-1])
    covered = np.zeros(n[::-1], bool)
    for rank in range(int(np.prod(grid))):
        coords = dist.coords_of(rank, grid)
        n_loc, off, llo, lup, _ = dist.partition_brick(n, lo, up, grid, coords)
        loc = ph.coefficient_at_qpoints(n_loc, llo, lup, degree, sub, lo, up, dc).reshape(n_loc[2], n_loc[1], n_loc[0], nq)
        sl = tuple(slice(off[a], off[a] + n_loc[a]) for a in (2, 1, 0))
        assert np.array_equal(loc, glob[sl])
        covered[sl] = True
        u_loc = ph.cutoff_cinfty_interpolate(n_loc, llo, lup, degree, center, radius=0.3).reshape([degree * m + 1 for m in n_loc][::-1])
        sln = tuple(slice(degree * off[a], degree * (off[a] + n_loc[a]) + 1) for a in (2, 1, 0))
        assert np.allclose(u_loc, u0[sln], rtol=1e-12, atol=1e-14)          # brick coordinates are lower + h*(c + xi)
    assert covered.all()
