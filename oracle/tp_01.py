"""ORACLE (test infrastructure only — never imported by the product path).

numpy restatement of the reference driver tests/tp_01.cc (heat + wave, `convergence_test` lambda
:56-725) with the time integrators of include/time_integrators.h (assemble_force :73-110, extrapolate
:180-190, TimeIntegratorFO::solve :300-321, TimeIntegratorWave::solve :400-447), the parameter defaults
of include/parameters.h:12-176, the analytic functions of include/exact_solution.h:27-197 and the
space-time error functional ErrorCalculator::evaluate_error (exact_solution.h:533-633).

Pinned by the reference's tests/tp_01.output (tests/golden/tp_01.json): the printed L-inf/L2/H1
space-time errors are solver independent and must match to the 6 printed digits.
"""
import math

import numpy as np

from . import fe_time as ft
from . import quadrature as Q
from . import spatial as S
from . import stmg

PI = math.pi


def default_parameters(dim=2):
    """include/parameters.h:12-80 defaults, JSON key names of parse() :92-144."""
    return {
        "spaceTimeMg": True, "mgTimeBeforeSpace": False, "timeType": "CGP", "problemType": "wave",
        "coarseningType": "space_or_time", "spaceTimeLevelFirst": True, "usePMg": False, "pMgType": "bisect",
        "nTimestepsAtOnce": 1, "nTimestepsAtOnceMin": -1, "feDegree": 1, "feDegreeMin": -1,
        "feDegreeMinSpace": -1, "nDegCycles": 1, "nRefCycles": 1, "frequency": 1.0, "refinement": 2,
        "spaceTimeConvergenceTest": True, "extrapolate": True, "hyperRectLowerLeft": [0.0] * dim,
        "hyperRectUpperRight": [1.0] * dim, "subdivisions": [1] * dim, "distortGrid": 0.0, "distortCoeff": 0.0,
        "endTime": 1.0, "smoother": "relaxation", "smoothingSteps": 1, "smoothingRange": 1.0,
        "relaxation": 0.0, "coarseGridSmootherType": "Smoother", "restrictIsTransposeProlongate": True,
        "variable": True, "smoothingEigCgNIterations": 20,
        "innerPreconditioner": "vanka",   # not a reference key: "jacobi" = stmg.PointJacobi inside Relaxation / Chebyshev
        "sourcePoint": None,         # parameters.h:79: midpoint of the DEFAULT box (member initialiser order)
    }


def parse_parameters(json_dict, dim=2):
    """Parameters<dim>::parse (parameters.h:85-176): strings -> typed values + derived defaults."""
    p = default_parameters(dim)
    p["sourcePoint"] = [0.5] * dim
    for k, v in json_dict.items():
        if k not in p:
            continue
        d = p[k]
        if isinstance(d, bool):
            p[k] = str(v).lower() == "true"
        elif isinstance(d, int):
            p[k] = int(v)
        elif isinstance(d, float):
            p[k] = float(v)
        elif isinstance(d, list):
            p[k] = [type(d[0])(float(x)) for x in str(v).split(",")]
        else:
            p[k] = str(v)
    nts = p["nTimestepsAtOnce"]
    if p["nTimestepsAtOnceMin"] == -1:
        p["nTimestepsAtOnceMin"] = nts // 2
    p["nTimestepsAtOnceMin"] = min(max(p["nTimestepsAtOnceMin"], 1), nts)
    lowest = 0 if p["timeType"] == "DG" else 1
    if p["feDegreeMin"] == -1:
        p["feDegreeMin"] = p["feDegree"] - 1
    p["feDegreeMin"] = min(max(p["feDegreeMin"], lowest), p["feDegree"])
    if p["feDegreeMinSpace"] == -1:
        p["feDegreeMinSpace"] = p["feDegreeMin"]
    return p


# ----------------------------------------------------------------------------- analytic functions
def exact_solution(pts, t, f=1.0):
    """ExactSolution::value, exact_solution.h:36-43."""
    return math.sin(2 * PI * f * t) * np.prod(np.sin(2 * PI * f * pts), axis=-1)


def exact_gradient(pts, t, f=1.0):
    """ExactSolution::gradient, exact_solution.h:44-57."""
    d = pts.shape[-1]
    s, c = np.sin(2 * PI * f * pts), np.cos(2 * PI * f * pts)
    g = np.empty_like(pts)
    for i in range(d):
        v = 2 * PI * f * math.sin(2 * PI * f * t) * c[..., i]
        for j in range(d):
            if j != i:
                v = v * s[..., j]
        g[..., i] = v
    return g


def exact_velocity(pts, t, f=1.0):
    """wave::ExactSolutionV, exact_solution.h:161-168."""
    return 2 * PI * f * math.cos(2 * PI * f * t) * np.prod(np.sin(2 * PI * f * pts), axis=-1)


def rhs_heat(pts, t, f=1.0):
    """RHSFunction::value, exact_solution.h:72-81."""
    d = pts.shape[-1]
    v = d * 4 * PI * PI * f * f * math.sin(2 * PI * f * t) + 2 * PI * f * math.cos(2 * PI * f * t)
    return v * np.prod(np.sin(2 * PI * f * pts), axis=-1)


def rhs_wave(pts, t, f=1.0):
    """wave::RHSFunction::value, exact_solution.h:183-191."""
    d = pts.shape[-1]
    v = (2.0 ** d) * (PI * f) ** 2 * math.sin(2 * PI * f * t)
    return v * np.prod(np.sin(2 * PI * f * pts), axis=-1)


# Integral of e * exp(-1/(1-|x|^2)) over the unit ball in 1, 2, 3 dimensions (deal.II function_lib_cutoff.cc,
# integral_Cinfty; re-derived here by quadrature to all printed digits, see tests/test_problem_host.py)
INTEGRAL_CINFTY = (1.20690032243787617533623799633, 1.26811216112759608094632335664, 1.1990039070192139033798473858)


def cutoff_cinfty(pts, center, radius=1.0e-2, integrate_to_one=True):
    """Functions::CutOffFunctionCinfty<dim>(radius, center, 1, invalid, integrate_to_one)::value (deal.II 9.6
    base/function_lib_cutoff.cc; constructed at tests/tp_01.cc:376-378): rescaling * e * exp(-r^2 / (r^2 - d^2)) for
    d < r (0 where the exponent is below -50), rescaling = 1 / (integral_Cinfty[dim-1] * r^dim).
    Third-party algorithm (deal.II is not vendored): parity unpinned by the reference's own outputs."""
    dim = pts.shape[-1]
    d = np.sqrt(np.sum((pts - np.asarray(center, float)) ** 2, axis=-1))
    r = radius
    resc = 1.0 / (INTEGRAL_CINFTY[dim - 1] * r ** dim) if integrate_to_one else 1.0
    out = np.zeros(d.shape)
    inside = d < r
    e = -r * r / (r * r - d[inside] ** 2)
    out[inside] = np.where(e < -50, 0.0, resc * math.e * np.exp(e))
    return out


# ----------------------------------------------------------------------------- spatial helpers
def integrate_rhs(space, fun):
    """VectorTools::create_right_hand_side(mapping, dof_handler, quad, f, rhs, constraints)
    (tp_01.cc:382-392): rhs_i = sum_q f(x_q) phi_i(x_q) JxW, constrained rows 0."""
    G, JxW, pts = space.geometry()
    fq = fun(pts) * JxW                                            # [C, nq]
    d, n1 = space.dim, space.n1
    C = fq.shape[0]
    t = fq.reshape((C,) + (n1,) * d)
    loc = S._interp(space.S.T.copy(), t, d).reshape(C, -1)
    rhs = np.zeros(space.n_dofs)
    np.add.at(rhs, space.cell_dofs.reshape(-1), loc.reshape(-1))
    rhs[space.constrained] = 0.0
    return rhs


def dof_points(space):
    """support points of the lexicographic DoF grid (MappingQ1 of the GLL nodes)."""
    d = space.dim
    X = space.mesh.cell_vertices()
    g = space.gll
    N1 = np.stack([1.0 - g, g], axis=0)
    if d == 2:
        Nv = np.stack([np.einsum("y,x->yx", N1[vy], N1[vx]).reshape(-1) for vy in range(2) for vx in range(2)])
    else:
        Nv = np.stack([np.einsum("z,y,x->zyx", N1[vz], N1[vy], N1[vx]).reshape(-1)
                       for vz in range(2) for vy in range(2) for vx in range(2)])
    pts_cell = np.einsum("cva,vq->cqa", X, Nv)
    pts = np.zeros((space.n_dofs, d))
    pts[space.cell_dofs.reshape(-1)] = pts_cell.reshape(-1, d)
    return pts


def interpolate(space, fun):
    """VectorTools::interpolate (tp_01.cc:393-400)."""
    return fun(dof_points(space))


def locate_point(space, x):
    """(cell, reference coordinates) of a real point on the MappingQ1 mesh: Newton inversion of the d-linear map on
    every cell at once, first cell (lexicographic) that contains the point (what RemotePointEvaluation returns first,
    tp_01.cc:461-462, 566-574)."""
    d = space.dim
    X = space.mesh.cell_vertices()                                   # [C, 2^d, d]
    C = X.shape[0]
    xi = np.full((C, d), 0.5)
    for _ in range(30):
        N = np.ones((C, 1 << d))
        dN = np.ones((C, 1 << d, d))
        for v in range(1 << d):
            for a in range(d):
                bit = (v >> a) & 1
                na = xi[:, a] if bit else 1.0 - xi[:, a]
                N[:, v] *= na
                for b in range(d):
                    dN[:, v, b] *= (1.0 if bit else -1.0) if a == b else na
        f = np.einsum("cv,cva->ca", N, X)
        J = np.einsum("cva,cvb->cab", X, dN)
        xi = xi + np.linalg.solve(J, (x[None, :] - f)[..., None])[..., 0]
    inside = np.all((xi >= -1e-10) & (xi <= 1 + 1e-10), axis=1)
    c = int(np.argmax(inside))
    assert inside[c], "point outside the mesh"
    return c, xi[c]


def point_evaluate(space, points, u):
    """FEPointEvaluation::evaluate at real points (tp_01.cc:463-481): u_h(x_p) for u[nb, N] -> [nb, n_points]."""
    d = space.dim
    out = np.zeros((u.shape[0], len(points)))
    for p, x in enumerate(np.asarray(points, float)):
        c, xi = locate_point(space, x)
        L = [Q.lagrange_eval(space.gll, np.array([xi[a]]))[:, 0] for a in range(d)]
        w = np.einsum("y,x->yx", L[1], L[0]).reshape(-1) if d == 2 else np.einsum("z,y,x->zyx", L[2], L[1], L[0]).reshape(-1)
        out[:, p] = u[:, space.cell_dofs[c]] @ w
    return out


def time_evaluation_matrix(ttype, r, samples):
    """get_time_evaluation_matrix (fe_time.h:307-326) for get_time_basis(type, r) (fe_time.cc:162-179)."""
    return Q.lagrange_eval(ft.time_nodes(ttype, r), np.arange(samples) / max(samples - 1.0, 1.0)).T.copy()


class ErrorCalculator:
    """exact_solution.h:503-649, called with (type, fe_degree, fe_degree) (tp_01.cc:492-498):
    QGauss<dim>(fe_degree+1) per cell in space, QGauss<1>(fe_degree+1) in time."""

    def __init__(self, ttype, time_degree, space_quad_degree, space, frequency):
        self.ttype, self.r, self.space, self.f = ttype, time_degree, space, frequency
        self.nt_dofs = time_degree + 1 if ttype == ft.DG else time_degree
        self.tq, self.tw = Q.gauss(time_degree + 1)
        xq, wq = Q.gauss(space_quad_degree + 1)
        d = space.dim
        self.Sq = Q.lagrange_eval(space.gll, xq).T.copy()
        self.Dq = Q.lagrange_deriv(space.gll, xq).T.copy()
        # geometry at these points
        X = space.mesh.cell_vertices()
        N1 = np.stack([1.0 - xq, xq], axis=0)
        dN1 = np.stack([-np.ones_like(xq), np.ones_like(xq)], axis=0)
        nq = len(xq)
        nv = 1 << d
        Nv = np.zeros((nv, nq ** d))
        dNv = np.zeros((nv, d, nq ** d))
        for v in range(nv):
            bits = [(v >> a) & 1 for a in range(d)]
            f = [N1[bits[a]] for a in range(d)]
            df = [dN1[bits[a]] for a in range(d)]
            if d == 2:
                Nv[v] = np.einsum("y,x->yx", f[1], f[0]).reshape(-1)
                dNv[v, 0] = np.einsum("y,x->yx", f[1], df[0]).reshape(-1)
                dNv[v, 1] = np.einsum("y,x->yx", df[1], f[0]).reshape(-1)
                w = np.einsum("y,x->yx", wq, wq).reshape(-1)
            else:
                Nv[v] = np.einsum("z,y,x->zyx", f[2], f[1], f[0]).reshape(-1)
                dNv[v, 0] = np.einsum("z,y,x->zyx", f[2], f[1], df[0]).reshape(-1)
                dNv[v, 1] = np.einsum("z,y,x->zyx", f[2], df[1], f[0]).reshape(-1)
                dNv[v, 2] = np.einsum("z,y,x->zyx", df[2], f[1], f[0]).reshape(-1)
                w = np.einsum("z,y,x->zyx", wq, wq, wq).reshape(-1)
        J = np.einsum("cva,vbq->cqab", X, dNv)
        self.Jinv = np.linalg.inv(J)
        self.JxW = np.linalg.det(J) * w[None, :]
        self.pts = np.einsum("cva,vq->cqa", X, Nv)
        self.basis_nodes = ft.time_nodes(ttype, time_degree)

    def numerical_solution(self, tq, x, prev_x, block_offset):
        """evaluate_numerical_solution (tp_01.cc:409-432)."""
        vals = Q.lagrange_eval(self.basis_nodes, [tq])[:, 0]
        out = np.zeros(self.space.n_dofs)
        for i, v in enumerate(vals):
            if v == 0.0:
                continue
            if self.ttype == ft.DG:
                out += v * x[block_offset + i]
            else:
                out += v * (prev_x if block_offset + i == 0 else x[block_offset + i - 1])
        out[self.space.constrained] = 0.0                      # constraints.distribute
        return out

    def evaluate_error(self, time, tau, x, prev_x, n_timesteps_at_once):
        s = self.space
        d, n1 = s.dim, s.n1
        l2 = h1 = 0.0
        l8 = -1.0
        for it in range(n_timesteps_at_once):
            for q in range(len(self.tq)):
                t = time + tau * it + self.tq[q] * tau
                cur_prev = prev_x if it == 0 else x[self.nt_dofs * it - 1]
                num = self.numerical_solution(self.tq[q], x, cur_prev, self.nt_dofs * it)
                u = num[s.cell_dofs].reshape((s.cell_dofs.shape[0],) + (n1,) * d)
                uq = S._interp(self.Sq, u, d).reshape(u.shape[0], -1)
                diff = uq - exact_solution(self.pts, t, self.f)
                l2sq = float((diff ** 2 * self.JxW).sum())
                l8 = max(l8, float(np.abs(diff).max()))
                # gradient
                if d == 2:
                    gx = np.einsum("ay,bx,cyx->cab", self.Sq, self.Dq, u).reshape(u.shape[0], -1)
                    gy = np.einsum("ay,bx,cyx->cab", self.Dq, self.Sq, u).reshape(u.shape[0], -1)
                    gref = np.stack([gx, gy], axis=-1)
                else:
                    gx = np.einsum("az,by,ex,czyx->cabe", self.Sq, self.Sq, self.Dq, u).reshape(u.shape[0], -1)
                    gy = np.einsum("az,by,ex,czyx->cabe", self.Sq, self.Dq, self.Sq, u).reshape(u.shape[0], -1)
                    gz = np.einsum("az,by,ex,czyx->cabe", self.Dq, self.Sq, self.Sq, u).reshape(u.shape[0], -1)
                    gref = np.stack([gx, gy, gz], axis=-1)
                greal = np.einsum("cqba,cqb->cqa", self.Jinv, gref)
                gd = greal - exact_gradient(self.pts, t, self.f)
                h1sq = float(((gd ** 2).sum(-1) * self.JxW).sum())
                l2 += tau * self.tw[q] * l2sq
                h1 += tau * self.tw[q] * h1sq
        return {"L2": l2, "Linf": l8, "H1": h1}


# ----------------------------------------------------------------------------- the driver
def build_levels(p, dim, refinement, fe_degree, tau, mg_dtype, coeff=None, mesh=None):
    """tests/tp_01.cc:171-321: level sequence, per-level FE degree, meshes, time weights, operators."""
    ttype = p["timeType"]
    nts = p["nTimestepsAtOnce"]
    space_time_mg = p["spaceTimeMg"]
    fe_degree_min = p["feDegreeMin"] if space_time_mg else fe_degree
    nts_min = max(p["nTimestepsAtOnceMin"], 1) if space_time_mg else nts
    poly_time = ft.get_poly_mg_sequence(fe_degree, fe_degree_min, p["pMgType"])
    poly_space = ft.get_poly_mg_sequence(fe_degree, p["feDegreeMinSpace"], p["pMgType"])
    if mesh is None:
        mesh = S.Mesh(dim, p["subdivisions"], refinement, p["hyperRectLowerLeft"], p["hyperRectUpperRight"],
                      distort=p["distortGrid"])
    meshes = [mesh]
    while all(v % 2 == 0 for v in meshes[-1].n) and len(meshes) < refinement + 1:
        meshes.append(meshes[-1].coarsen())
    meshes = meshes[::-1]                                     # coarse -> fine
    mg_type_level = ft.get_mg_sequence(len(meshes), poly_time, poly_space, nts, nts_min, "t", p["coarseningType"],
                                       p["mgTimeBeforeSpace"], p["usePMg"], p["spaceTimeLevelFirst"])
    n_levels = len(mg_type_level) + 1
    # get_space_time_triangulation (fe_time.h:130-155)
    level_mesh = [None] * n_levels
    mi = len(meshes) - 1
    level_mesh[-1] = meshes[mi]
    for ii in range(len(mg_type_level) - 1, -1, -1):
        if mg_type_level[ii] == "h":
            mi -= 1
        level_mesh[ii] = meshes[mi]
    # fe_pmg (fe_time.h:68-107 with strides = 1): space degree = poly_space + 1, stepping at 'p' levels
    space_degrees = [q + 1 for q in poly_space]
    fi = 0 if p["usePMg"] else len(space_degrees) - 1
    level_degree = []
    for l in range(n_levels):
        level_degree.append(space_degrees[fi])
        if p["usePMg"] and l < len(mg_type_level) and mg_type_level[l] == "p":
            fi += 1
    if p["problemType"] == "heat":
        fetw = ft.get_fe_time_weights_levels(ttype, tau, nts, mg_type_level, poly_time)
    else:
        fetw = ft.get_fe_time_weights_wave_levels(ttype, tau, nts, mg_type_level, poly_time)
    spaces = [S.Space(level_mesh[l], level_degree[l]) for l in range(n_levels)]
    ops = [stmg.LevelOperator(spaces[l], fetw[l][0], fetw[l][1], mg_dtype, coeff) for l in range(n_levels)]
    smoother = {"relaxation": 1, "chebyshev": 2, "identity": 0}[p["smoother"].lower()]
    ptypes = ft.get_precondition_stmg_types(mg_type_level, p["coarseningType"], p["mgTimeBeforeSpace"],
                                            p["spaceTimeLevelFirst"], smoother)
    return dict(mg_type_level=mg_type_level, poly_time=poly_time, spaces=spaces, ops=ops, ptypes=ptypes,
                level_degree=level_degree, fetw=fetw)


def convergence_test(p, dim, refinement, fe_degree, mg_dtype=np.float32, use_mg=True, max_steps=None,
                     solver="fgmres", reduce=1e-12, abstol=1e-12, tau=None, return_state=False):
    """One (refinement, degree) run of tests/tp_01.cc:56-725.  Returns the table row.
    tau: overrides the time step size of tp_01.cc:106-109 (tests/transfer_01.cc:429 uses 2^-(i+1) on a fixed mesh)."""
    ttype = p["timeType"]
    is_cgp = ttype == ft.CGP
    nts = p["nTimestepsAtOnce"]
    nt_dofs = fe_degree if is_cgp else fe_degree + 1
    nb = nt_dofs * nts
    f = p["frequency"]
    mesh = S.Mesh(dim, p["subdivisions"], refinement, p["hyperRectLowerLeft"], p["hyperRectUpperRight"],
                  distort=p["distortGrid"])
    space = S.Space(mesh, fe_degree + 1)
    spc_step = min((mesh.upper[a] - mesh.lower[a]) / mesh.subdivisions[a] for a in range(dim)) * 1.0
    # minimal_cell_diameter / sqrt(dim) of the unrefined Cartesian mesh = min edge for cubes
    diam = math.sqrt(sum(((mesh.upper[a] - mesh.lower[a]) / mesh.subdivisions[a]) ** 2 for a in range(dim)))
    spc_step = diam / math.sqrt(dim)
    time, end_time = 0.0, p["endTime"]
    n_steps = int((end_time - time) / spc_step)
    if tau is None:
        tau = (end_time - time) * 2.0 ** (-(refinement + 1)) / n_steps
    wave = p["problemType"] == "wave"
    coeff = None
    if not p["spaceTimeConvergenceTest"]:
        coeff = S.Coefficient(dim, p["subdivisions"], p["hyperRectLowerLeft"], p["hyperRectUpperRight"], p["distortCoeff"])

    K = S.MatrixFreeOperator(space, 0.0, 1.0)
    M = S.MatrixFreeOperator(space, 1.0, 0.0)
    if coeff is not None:
        K.evaluate_coefficient(coeff)
    A1, B1, G1, Z1 = ft.get_fe_time_weights(ttype, fe_degree, tau, 1)
    A, B, G, Z = ft.get_fe_time_weights(ttype, fe_degree, tau, nts)
    zero = np.zeros_like(G)
    if wave:
        lhs_uK, lhs_uM, rhs_uK, rhs_uM, rhs_vM = ft.get_fe_time_weights_wave(ttype, A1, B1, G1, Z1, nts)
        rhs_matrix_v = S.SystemMatrix(K, M, zero, rhs_vM)
    else:
        lhs_uK, lhs_uM = A, B
        rhs_uK = G if is_cgp else zero
        rhs_uM = Z if is_cgp else G
    # assembled fine operator for the oracle's Krylov loop (identical to the matrix-free one)
    fine = stmg.LevelOperator(space, lhs_uK, lhs_uM, np.float64, coeff)
    rhs_matrix = S.SystemMatrix(K, M, rhs_uK, rhs_uM)

    gmg = None
    lv = None
    if use_mg:
        lv = build_levels(p, dim, refinement, fe_degree, tau, mg_dtype, coeff, mesh)
        inner = None
        if p["innerPreconditioner"] == "jacobi":
            inner = [stmg.PointJacobi(o, mg_dtype) if t != 0 else None for o, t in zip(lv["ops"], lv["ptypes"])]
        gmg = stmg.GMG(ttype, lv["ops"], lv["spaces"], lv["mg_type_level"], lv["poly_time"], nts, lv["ptypes"],
                       mg_dtype, vanka=inner, smoothing_steps=p["smoothingSteps"], relaxation=p["relaxation"],
                       smoothing_range=p["smoothingRange"], eig_n_iterations=p["smoothingEigCgNIterations"],
                       variable=p["variable"], restrict_is_transpose_prolongate=p["restrictIsTransposeProlongate"])
    rhs_fun = (lambda pts, t: rhs_wave(pts, t, f)) if wave else (lambda pts, t: rhs_heat(pts, t, f))
    if not p["spaceTimeConvergenceTest"]:
        rhs_fun = None
    quad_time = ft.time_nodes(ttype, fe_degree)
    errc = ErrorCalculator(ttype, fe_degree, fe_degree, space, f)

    x = np.zeros((nb, space.n_dofs))
    v = np.zeros((nb, space.n_dofs))
    if p["spaceTimeConvergenceTest"]:
        x[-1] = interpolate(space, lambda pts: exact_solution(pts, 0.0, f))
    else:
        # tp_01.cc:374-381, 551-553: initial value = C-infinity bump of radius 1e-2 around sourcePoint, v(0) = 0, f = 0
        x[-1] = interpolate(space, lambda pts: cutoff_cinfty(pts, p["sourcePoint"]))
    if wave:
        if p["spaceTimeConvergenceTest"]:
            v[-1] = interpolate(space, lambda pts: exact_velocity(pts, 0.0, f))
        Ainv = np.linalg.inv(A1)
        AixB, AixG, AixZ = Ainv @ B1, Ainv @ G1, Ainv @ Z1
        if ttype == ft.DG:
            AixG = -AixG
        else:
            AixZ = -AixZ
    # point evaluation functionals of the practical runs (tp_01.cc:455-459, 559-635)
    real_points = [[0.75, 0.0]] if dim == 2 else [[0.75, 0.0, 0.0], [0.0, 0.0, 0.75], [0.75, 0.1, 0.75]]
    functional_rows = []
    if not p["spaceTimeConvergenceTest"]:
        prev_pt = point_evaluate(space, real_points, x[-1:])[0]
        samples = (fe_degree + 1) ** 2
        TE = time_evaluation_matrix(ttype, fe_degree, samples)
    l2 = h1 = 0.0
    l8 = -1.0
    total_it = 0
    n_solves = 0
    its_per_solve = []
    Aop = fine.vmult
    Mop = gmg.vmult if gmg is not None else (lambda r: r.copy())
    while time < end_time:
        n_solves += 1
        prev_x = x[-1].copy()
        rhs = rhs_matrix.vmult_slice(prev_x, nb)
        if wave:
            prev_v = v[-1].copy()
            rhs = rhs_matrix_v.vmult_slice_add(rhs, prev_v)
        # assemble_force (time_integrators.h:73-110)
        if rhs_fun is not None:
            for it in range(nts):
                for j, c in enumerate(quad_time):
                    t_ = time + tau * it + tau * c
                    fv = integrate_rhs(space, lambda pts: rhs_fun(pts, t_))
                    if ttype == ft.DG:
                        rhs[it * nt_dofs + j] += A1[j, j] * fv
                    elif j == 0:
                        for i in range(nt_dofs):
                            rhs[it * nt_dofs + i] += -G1[i, 0] * fv
                    else:
                        rhs[it * nt_dofs + j - 1] += A1[j - 1, j - 1] * fv
        # extrapolate (time_integrators.h:180-190)
        x0 = np.tile(prev_x, (nb, 1)) if p["extrapolate"] else np.zeros_like(x)
        x, its, _ = stmg.fgmres(Aop, x0, rhs, Mop, abstol=abstol, reduce=reduce)
        total_it += its
        its_per_solve.append(its)
        if wave:
            v = np.zeros_like(x)
            for it in range(nts):
                pu = prev_x if it == 0 else x[it * nt_dofs - 1]
                sl = slice(it * nt_dofs, (it + 1) * nt_dofs)
                v[sl] += AixB @ x[sl]
                if ttype == ft.DG:
                    v[sl] += AixG @ pu[None, :]
                else:
                    pv = prev_v if it == 0 else v[it * nt_dofs - 1]
                    v[sl] += AixG @ pv[None, :]
                    v[sl] += AixZ @ pu[None, :]
        x[:, space.constrained] = 0.0
        if p["spaceTimeConvergenceTest"]:
            e = errc.evaluate_error(time, tau, x, prev_x, nts)
            l2 += e["L2"]
            h1 += e["H1"]
            l8 = max(l8, e["Linf"])
        else:
            vals = point_evaluate(space, real_points, x)
            cg = 1 if is_cgp else 0
            for it in range(nts):
                pt = np.zeros((fe_degree + 1, len(real_points)))
                if cg:
                    pt[0] = prev_pt
                pt[cg:] = vals[it * nt_dofs:(it + 1) * nt_dofs]
                res = TE @ pt
                for row in range(samples):
                    functional_rows.append((time + tau * (it + row / max(samples - 1.0, 1.0)),) + tuple(res[row]))
                prev_pt = vals[(it + 1) * nt_dofs - 1]
        time += nts * tau
        if max_steps is not None and n_solves >= max_steps:
            break
    state = dict(x=x, v=v if wave else None, iterations_per_solve=its_per_solve,
                 functional_rows=functional_rows) if return_state else {}
    return dict(state, cells=mesh.n_cells, s_dofs=space.n_dofs, t_dofs=nb, iterations=total_it, timesteps=n_solves,
                linf=l8, l2=math.sqrt(l2), h1=math.sqrt(h1), tau=tau,
                levels="".join(lv["mg_type_level"]) if lv else "", n_levels=(len(lv["mg_type_level"]) + 1) if lv else 0)
