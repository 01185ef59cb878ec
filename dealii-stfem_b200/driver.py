"""tp_01-style driver on top of the C ABI (reference tests/tp_01.cc:56-725): parameters with the reference's
JSON keys (include/parameters.h:92-176), level hierarchy (tp_01.cc:171-321), time loop (:646-702) and the
convergence-table row (:712-723).  Everything numerical happens in libstfem_b200.so; this file is host glue
(the reference's driver is host C++ doing the same bookkeeping).  No oracle import — product path."""
import ctypes as C
import math

import numpy as np

from . import capi
from . import fe_time_host as ft
from . import problem_host

_vp, _vpp, _dp = C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_double)


class TiDesc(C.Structure):
    _fields_ = [("time_type", C.c_int), ("time_degree", C.c_int), ("n_timesteps_at_once", C.c_int), ("problem", C.c_int),
                ("Alpha_1", _dp), ("Beta_1", _dp), ("Gamma_1", _dp), ("Zeta_1", _dp), ("matrix", _vp), ("preconditioner", _vp),
                ("rhs_matrix", _vp), ("rhs_matrix_v", _vp), ("rhs_function_id", C.c_int), ("frequency", C.c_double),
                ("extrapolate", C.c_int), ("gmres_tolerance", C.c_double), ("abs_tol", C.c_double), ("max_iterations", C.c_int),
                ("max_basis_size", C.c_int)]


capi.SYMBOLS.update({
    "stfem_dev_copy": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "stfem_ti_create": (C.c_int, [C.POINTER(TiDesc), _vpp]),
    "stfem_ti_destroy": (C.c_int, [_vp]),
    "stfem_ti_solve_heat": (C.c_int, [_vp, _vpp, _vp, _vpp, C.c_double, C.c_double, C.POINTER(C.c_int)]),
    "stfem_ti_solve_wave": (C.c_int, [_vp, _vpp, _vpp, _vpp, _vp, _vp, C.c_double, C.c_double, C.POINTER(C.c_int)]),
    "stfem_ti_last_residuals": (C.c_int, [_vp, _dp, _dp]),
    "stfem_interpolate": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_double, _vp]),
    "stfem_integrate_rhs": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _vp]),
    "stfem_point_evaluate": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_int, _vpp, _dp]),
    "stfem_evaluate_error": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vpp, _vp, C.c_double, C.c_double, C.c_double,
                                       C.c_int, _dp]),
})

F_ZERO, F_EXACT, F_RHS_HEAT, F_EXACT_V, F_RHS_WAVE = 0, 1, 2, 3, 4


def default_parameters(dim=2):
    """include/parameters.h:12-80."""
    return {
        "spaceTimeMg": True, "mgTimeBeforeSpace": False, "timeType": "CGP", "problemType": "wave",
        "coarseningType": "space_or_time", "spaceTimeLevelFirst": True, "usePMg": False, "pMgType": "bisect",
        "nTimestepsAtOnce": 1, "nTimestepsAtOnceMin": -1, "feDegree": 1, "feDegreeMin": -1, "feDegreeMinSpace": -1,
        "nDegCycles": 1, "nRefCycles": 1, "frequency": 1.0, "refinement": 2, "spaceTimeConvergenceTest": True,
        "extrapolate": True, "hyperRectLowerLeft": [0.0] * dim, "hyperRectUpperRight": [1.0] * dim,
        "subdivisions": [1] * dim, "distortGrid": 0.0, "distortCoeff": 0.0, "endTime": 1.0, "smoother": "relaxation",
        "smoothingSteps": 1, "smoothingRange": 1.0, "relaxation": 0.0, "coarseGridSmootherType": "Smoother",
        "restrictIsTransposeProlongate": True, "variable": True, "smoothingEigCgNIterations": 20,
        "sourcePoint": [0.5] * dim,  # parameters.h:79: midpoint of the DEFAULT box (member initialiser order)
        "functionalFile": "functionals.txt", "doOutput": False, "printTiming": False,
        # accepted and ignored like in tests/tp_01.cc (unused there): relativeTolerance, timeRefineOffset, deltaTime
        "relativeTolerance": 1.0e-12, "timeRefineOffset": 1, "deltaTime": 0.0,
        "agglomerateBelow": 16,     # multi-GPU only (not a reference key): see HeatWaveProblem
        "innerPreconditioner": "vanka",   # not a reference key: "jacobi" = point-Jacobi inside Relaxation / Chebyshev
        "vankaStorage": "level",          # not a reference key: "half" = FP16 storage of the dense patch inverses
        "levelKernelVariant": 0,          # not a reference key: stfem_op_desc::kernel_variant of the multigrid level operators
    }


def parse_parameters(json_dict, dim=2):
    """Parameters<dim>::parse (include/parameters.h:85-176)."""
    p = default_parameters(dim)
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


def level_time_weights(ttype, tau, nts, mg_type_level, poly_time, wave):
    """get_fe_time_weights / get_fe_time_weights_wave per level (fe_time.h:411-474)."""
    out = [None] * (len(mg_type_level) + 1)
    pi, n, t = len(poly_time) - 1, nts, tau
    out[-1] = ft.get_fe_time_weights(ttype, poly_time[pi], t, n)
    idx = len(out) - 2
    for mgt in reversed(mg_type_level):
        if mgt == "k":
            pi -= 1
        elif mgt == "t":
            n //= 2
            t *= 2
        out[idx] = ft.get_fe_time_weights(ttype, poly_time[pi], t, n)
        idx -= 1
    if wave:
        out = [ft.get_fe_time_weights_wave(ttype, w[0], w[1], w[2], w[3]) for w in out]
    return out


def format_functional_row(t, values):
    """One line of the functional file (tp_01.cc:620-630): `setw(16) << scientific << t`, then for every point
    `setw(16) << scientific << " " << value` — the width applies to the blank, so a value follows 16 blanks."""
    return "%16s" % ("%.6e" % t) + "".join(" " * 16 + "%.6e" % v for v in values) + "\n"


class HeatWaveProblem:
    """One (refinement, degree) run of the reference's convergence_test lambda."""

    def __init__(self, ctx, params, dim, refinement, fe_degree, vertices_fn=None, mg_number_type=capi.F32, space_degree=None,
                 partition=None, gmres_tolerance=1e-12, abs_tol=1e-12):
        """partition = (proc_grid, coords): this rank owns one brick of the box partition (multi-GPU runs; the context
        must hold a communicator, dist.init_comm).  params describe the GLOBAL mesh."""
        self.ctx, self.p, self.dim = ctx, params, dim
        self.partition = partition
        p = params
        self.ttype = p["timeType"]
        self.is_cgp = self.ttype == "CGP"
        self.r = fe_degree
        self.k = fe_degree + 1 if space_degree is None else space_degree          # tp_01.cc:77
        self.nts = p["nTimestepsAtOnce"]
        self.nd = self.r if self.is_cgp else self.r + 1
        self.nb = self.nd * self.nts
        self.wave = p["problemType"] == "wave"
        sub = p["subdivisions"]
        lo, up = p["hyperRectLowerLeft"], p["hyperRectUpperRight"]
        self.n_cells = [s * (1 << refinement) for s in sub]
        diam = math.sqrt(sum(((up[a] - lo[a]) / sub[a]) ** 2 for a in range(dim)))
        spc_step = diam / math.sqrt(dim)                                           # tp_01.cc:87
        n_steps = int((p["endTime"] - 0.0) / spc_step)
        self.tau = p["endTime"] * 2.0 ** (-(refinement + 1)) / n_steps            # tp_01.cc:106-109
        self.freq = p["frequency"]
        # ---- level hierarchy (tp_01.cc:171-214)
        stmg = p["spaceTimeMg"]
        fe_degree_min = p["feDegreeMin"] if stmg else fe_degree
        nts_min = max(p["nTimestepsAtOnceMin"], 1) if stmg else self.nts
        self.poly_time = ft.get_poly_mg_sequence(fe_degree, fe_degree_min, p["pMgType"])
        poly_space = ft.get_poly_mg_sequence(fe_degree, p["feDegreeMinSpace"], p["pMgType"])
        n_sp_lvl = refinement + 1
        self.mg_type_level = ft.get_mg_sequence(n_sp_lvl, self.poly_time, poly_space, self.nts, nts_min, "t", p["coarseningType"],
                                                p["mgTimeBeforeSpace"], p["usePMg"], p["spaceTimeLevelFirst"])
        nl = len(self.mg_type_level) + 1
        level_ref = [None] * nl
        rr = refinement
        level_ref[-1] = rr
        for ii in range(nl - 2, -1, -1):
            if self.mg_type_level[ii] == "h":
                rr -= 1
            level_ref[ii] = rr
        shift = self.k - (fe_degree + 1)
        space_degrees = [q + 1 + shift for q in poly_space]
        fi = 0 if p["usePMg"] else len(space_degrees) - 1
        level_degree = []
        for l in range(nl):
            level_degree.append(space_degrees[fi])
            if p["usePMg"] and l < nl - 1 and self.mg_type_level[l] == "p":
                fi += 1
        smoother = {"relaxation": 1, "chebyshev": 2, "identity": 0}[p["smoother"].lower()]
        self.ptypes = ft.get_precondition_stmg_types(self.mg_type_level, p["coarseningType"], p["mgTimeBeforeSpace"],
                                                     p["spaceTimeLevelFirst"], smoother)
        fetw = level_time_weights(self.ttype, self.tau, self.nts, self.mg_type_level, self.poly_time, self.wave)
        # ---- meshes per refinement (geometric coarsening sequence: every second vertex)
        self.meshes = {}
        self.ghost_desc = {}        # refinement -> the same for the ghost-extended brick of a partitioned level
        mesh_desc = {}              # refinement -> (n_cells, lower, upper, vertices) of the (local) mesh, for host set-up helpers
        fine_vertices = vertices_fn(self.n_cells) if vertices_fn is not None else None
        for rf in sorted(set(level_ref)):
            n = [s * (1 << rf) for s in sub]
            v = None
            if fine_vertices is not None:
                step = 1 << (refinement - rf)
                v = np.ascontiguousarray(fine_vertices[tuple([slice(None, None, step)] * dim)]).reshape(-1, dim)
            # coarse-level agglomeration: levels whose local brick would have fewer than `agglomerate_below` cells per
            # direction are kept as the GLOBAL mesh on every rank (solved redundantly, one all-reduce + one broadcast per
            # V-cycle at the switch) instead of exchanging latency-bound halos on tiny bricks
            agglomerate = partition is not None and rf != refinement and \
                min(nn // g for nn, g in zip(n, partition[0])) < self.p.get("agglomerateBelow", 16)
            if partition is None or agglomerate:
                self.meshes[rf] = capi.Mesh(ctx, n, lower=lo, upper=up, vertices=v)
                mesh_desc[rf] = (n, lo, up, v)
            else:
                from . import dist
                n_loc, off, llo, lup, mask = dist.partition_brick(n, lo, up, partition[0], partition[1])
                # ghost layer (deal.II's ghost cells): the brick extended by one cell layer across every face shared with
                # another rank - what the dense cell-patch smoother assembles its interface patches on
                glo, ghi = dist.ghost_layers(partition[0], partition[1])
                hh = [(up[a] - lo[a]) / n[a] for a in range(dim)]
                n_ext = [n_loc[a] + glo[a] + ghi[a] for a in range(dim)]
                lo_ext = [llo[a] - glo[a] * hh[a] for a in range(dim)]
                up_ext = [lup[a] + ghi[a] * hh[a] for a in range(dim)]
                v_loc = v_ext = None
                if v is not None:
                    v_loc = dist.brick_vertices(v, n, off, n_loc)
                    v_ext = dist.brick_vertices(v, n, off, n_loc, glo, ghi)
                self.meshes[rf] = capi.Mesh(ctx, n_loc, lower=llo, upper=lup, vertices=v_loc, dirichlet_faces=mask)
                dist.set_partition(self.meshes[rf], partition[0], partition[1])
                if v_ext is not None:
                    dist.set_ghost_vertices(self.meshes[rf], v_ext)
                mesh_desc[rf] = (list(n_loc), list(llo), list(lup), v_loc)
                self.ghost_desc[rf] = (n_ext, lo_ext, up_ext, v_ext)
        self.mesh_desc = mesh_desc
        self._coeff_cache = {}
        self.level_ops = [capi.Operator(self.meshes[level_ref[l]], level_degree[l], fetw[l][0], fetw[l][1], number_type=mg_number_type,
                                        variant=int(p.get("levelKernelVariant", 0)),
                                        **self._laplace_coefficient(level_ref[l], level_degree[l])) for l in range(nl)]
        for l in range(nl):
            self._set_ghost_coefficient(self.level_ops[l], level_ref[l], level_degree[l])
        self.mg = capi.Multigrid(ctx, self.level_ops, self.mg_type_level, self.ptypes, self.ttype, self.nts, self.poly_time,
                                 smoothing_steps=p["smoothingSteps"], relaxation=p["relaxation"], smoothing_range=p["smoothingRange"],
                                 eig_n_iterations=p["smoothingEigCgNIterations"], variable=p["variable"],
                                 restrict_is_transpose_prolongate=p["restrictIsTransposeProlongate"],
                                 inner_preconditioner=p.get("innerPreconditioner", "vanka"),
                                 vanka_storage=p.get("vankaStorage", "level"),
                                 coarse_grid_maxiter=0 if str(p.get("coarseGridSmootherType", "Smoother")) == "Smoother"
                                 else int(p.get("coarseGridMaxiter", 10)),
                                 coarse_grid_abstol=float(p.get("coarseGridAbstol", 1e-20))) if p.get("useMg", True) else None
        # ---- fine operators (tp_01.cc:121-168)
        fmesh = self.meshes[refinement]
        self.fmesh = fmesh
        A1, B1, G1, Z1 = ft.get_fe_time_weights(self.ttype, fe_degree, self.tau, 1)
        A, B, G, Z = ft.get_fe_time_weights(self.ttype, fe_degree, self.tau, self.nts)
        zero = np.zeros_like(G)
        self.rhs_matrix_v = None
        if self.wave:
            lhs_uK, lhs_uM, rhs_uK, rhs_uM, rhs_vM = ft.get_fe_time_weights_wave(self.ttype, A1, B1, G1, Z1, self.nts)
            self.rhs_matrix_v = capi.Operator(fmesh, self.k, zero, rhs_vM, **self._laplace_coefficient(refinement, self.k))
        else:
            lhs_uK, lhs_uM = A, B
            rhs_uK = G if self.is_cgp else zero
            rhs_uM = Z if self.is_cgp else G
        self.matrix = capi.Operator(fmesh, self.k, lhs_uK, lhs_uM, **self._laplace_coefficient(refinement, self.k))
        self.rhs_matrix = capi.Operator(fmesh, self.k, rhs_uK, rhs_uM, **self._laplace_coefficient(refinement, self.k))
        self._w = [np.ascontiguousarray(m, np.float64) for m in (A1, B1, G1, Z1)]
        d = TiDesc()
        d.time_type = 1 if self.is_cgp else 2
        d.time_degree, d.n_timesteps_at_once, d.problem = fe_degree, self.nts, 2 if self.wave else 1
        d.Alpha_1, d.Beta_1, d.Gamma_1, d.Zeta_1 = (capi._dptr(m) for m in self._w)
        d.matrix, d.rhs_matrix = self.matrix.h, self.rhs_matrix.h
        d.preconditioner = self.mg.h if self.mg is not None else None
        d.rhs_matrix_v = self.rhs_matrix_v.h if self.rhs_matrix_v is not None else None
        conv = p["spaceTimeConvergenceTest"]
        d.rhs_function_id = (F_RHS_WAVE if self.wave else F_RHS_HEAT) if conv else F_ZERO
        d.frequency, d.extrapolate = self.freq, int(p["extrapolate"])
        d.gmres_tolerance, d.abs_tol, d.max_iterations, d.max_basis_size = gmres_tolerance, abs_tol, 200, 100   # time_integrators.h:56-59
        self.ti = C.c_void_p()
        capi.check(capi.lib().stfem_ti_create(C.byref(d), C.byref(self.ti)))
        self.n = self.matrix.n
        nbv = self.nb
        self.x = capi.DeviceBlockVector(ctx, nbv, self.n)
        self.rhs = capi.DeviceBlockVector(ctx, nbv, self.n)
        self.v = capi.DeviceBlockVector(ctx, nbv, self.n) if self.wave else None
        self.prev_x = capi.DeviceBlockVector(ctx, 1, self.n)
        self.prev_v = capi.DeviceBlockVector(ctx, 1, self.n) if self.wave else None
        self.x.zero()
        L = capi.lib()
        if self.wave:
            self.v.zero()
        if conv:
            capi.check(L.stfem_interpolate(fmesh.h, self.k, F_EXACT, self.freq, 0.0, self.x.ptrs[nbv - 1]))      # tp_01.cc:551
            if self.wave:
                capi.check(L.stfem_interpolate(fmesh.h, self.k, F_EXACT_V, self.freq, 0.0, self.v.ptrs[nbv - 1]))
        else:
            # tp_01.cc:374-381: u(0) = C-infinity bump of radius 1e-2 around sourcePoint, v(0) = 0, no source term
            n_, lo_, up_, v_ = mesh_desc[refinement]
            u0 = problem_host.cutoff_cinfty_interpolate(n_, lo_, up_, self.k, p["sourcePoint"], vertices=v_)
            capi.check(L.stfem_dev_upload(ctx.h, self.x.ptrs[nbv - 1], u0.ctypes.data, u0.nbytes))
            ctx.synchronize()
        self.time = 0.0
        self.total_iterations = 0
        self.n_solves = 0
        self.err = np.array([0.0, -1.0, 0.0])
        # point-evaluation functionals of the practical runs (tp_01.cc:455-459, 559-583)
        self.real_points = np.array([[0.75, 0.0]] if dim == 2 else [[0.75, 0.0, 0.0], [0.0, 0.0, 0.75], [0.75, 0.1, 0.75]])
        self.functional_rows = []           # (t, value at every point), in the order the reference writes them
        self.functional_file = None         # set to a path to append the reference's text format
        self._prev_pt = None
        if not conv:
            self._prev_pt = self.point_evaluate([self.x.ptrs[nbv - 1]])[0]

    def _laplace_coefficient(self, rf, degree):
        """K_mf.evaluate_coefficient(coeff) when !spaceTimeConvergenceTest (tp_01.cc:118-119, 279-280): keyword arguments
        for capi.Operator.  On Cartesian meshes whose cells do not straddle a coefficient jump the table is constant per
        cell and the per-cell path (Cartesian kernel) is taken."""
        p = self.p
        if p["spaceTimeConvergenceTest"]:
            return {}
        key = (rf, degree)
        if key not in self._coeff_cache:
            n, lo, up, v = self.mesh_desc[rf]
            cq = problem_host.coefficient_at_qpoints(n, lo, up, degree, p["subdivisions"], p["hyperRectLowerLeft"],
                                                     p["hyperRectUpperRight"], p["distortCoeff"], vertices=v)
            if v is None and np.all(cq == cq[:, :1]):
                self._coeff_cache[key] = {"laplace_coeff_cell": np.ascontiguousarray(cq[:, 0])}
            else:
                self._coeff_cache[key] = {"laplace_coeff_q": cq}
        return self._coeff_cache[key]

    def _set_ghost_coefficient(self, op, rf, degree):
        """Partitioned levels of practical runs: the Laplace coefficient of the ghost cells (dense Vanka set-up)."""
        if rf not in self.ghost_desc or self.p["spaceTimeConvergenceTest"]:
            return
        from . import dist
        p = self.p
        n, lo, up, v = self.ghost_desc[rf]
        cq = problem_host.coefficient_at_qpoints(n, lo, up, degree, p["subdivisions"], p["hyperRectLowerLeft"],
                                                 p["hyperRectUpperRight"], p["distortCoeff"], vertices=v)
        if "laplace_coeff_cell" in self._coeff_cache[(rf, degree)]:
            dist.set_ghost_coefficients(op, np.ascontiguousarray(cq[:, 0]), None)
        else:
            dist.set_ghost_coefficients(op, None, cq)

    def point_evaluate(self, block_ptrs):
        """u_b(real_points[p]) for the given device vectors: [len(block_ptrs), n_points]."""
        nbk, npt = len(block_ptrs), len(self.real_points)
        out = np.zeros((nbk, npt))
        ptrs = (C.c_void_p * nbk)(*block_ptrs)
        pts = np.ascontiguousarray(self.real_points, np.float64)
        capi.check(capi.lib().stfem_point_evaluate(self.fmesh.h, self.k, npt, capi._dptr(pts), nbk, ptrs, capi._dptr(out)))
        return out

    def _do_point_evaluation(self):
        """do_point_evaluation (tp_01.cc:584-635): every time DoF at the points, interpolated in time at (r+1)^2
        equidistant samples per time step."""
        samples = (self.r + 1) * (self.r + 1)
        TE = ft.get_time_evaluation_matrix(self.ttype, self.r, samples)
        vals = self.point_evaluate([self.x.ptrs[b] for b in range(self.nb)])
        cg = 1 if self.is_cgp else 0
        lines = []
        for it in range(self.nts):
            pt = np.zeros((self.r + 1, len(self.real_points)))
            if cg:
                pt[0] = self._prev_pt
            pt[cg:] = vals[it * self.nd:(it + 1) * self.nd]
            res = TE @ pt
            for row in range(samples):
                t_ = self.time + self.tau * (it + row / max(samples - 1.0, 1.0))
                self.functional_rows.append((t_,) + tuple(res[row]))
                lines.append(format_functional_row(t_, res[row]))
            lines.append("\n")
            self._prev_pt = vals[(it + 1) * self.nd - 1]
        if self.functional_file:
            with open(self.functional_file, "a") as f:
                f.writelines(lines)

    def _copy(self, dst_ptr, src_ptr):
        # device-to-device copy of one spatial vector on the library stream
        capi.check(capi.lib().stfem_dev_copy(self.ctx.h, dst_ptr, src_ptr, self.n * 8))

    def step(self, evaluate_error=True):
        """One pass of the time loop body (tp_01.cc:646-685)."""
        L = capi.lib()
        nb = self.nb
        self._copy(self.prev_x.ptrs[0], self.x.ptrs[nb - 1])
        it = C.c_int()
        if not self.wave:
            capi.check(L.stfem_ti_solve_heat(self.ti, self.x.ptrs, self.prev_x.ptrs[0], self.rhs.ptrs, self.time, self.tau, C.byref(it)))
        else:
            self._copy(self.prev_v.ptrs[0], self.v.ptrs[nb - 1])
            capi.check(L.stfem_ti_solve_wave(self.ti, self.x.ptrs, self.v.ptrs, self.rhs.ptrs, self.prev_x.ptrs[0],
                                             self.prev_v.ptrs[0], self.time, self.tau, C.byref(it)))
        self.total_iterations += it.value
        self.n_solves += 1
        if evaluate_error and self.p["spaceTimeConvergenceTest"]:
            capi.check(L.stfem_evaluate_error(self.fmesh.h, self.k, 1 if self.is_cgp else 2, self.r, self.nts, self.x.ptrs,
                                              self.prev_x.ptrs[0], self.time, self.tau, self.freq, self.r + 1, capi._dptr(self.err)))
        elif self._prev_pt is not None:
            self._do_point_evaluation()
        self.time += self.nts * self.tau
        return it.value

    def run(self, max_steps=None):
        while self.time < self.p["endTime"]:
            self.step()
            if max_steps is not None and self.n_solves >= max_steps:
                break
        return self.row()

    def row(self):
        """One line of the convergence table (tests/tp_01.cc:703-715).  Partitioned runs: the error integrals of the local
        cells are summed (L2, H1) / maximised (Linf) over the ranks and the DoF count is the global one, as the reference
        does with Utilities::MPI::sum / max (include/exact_solution.h:612-631)."""
        err, s_dofs = self.err, int(self.n)
        if self.partition is not None:
            from . import dist
            sums = dist.allreduce(self.ctx, [self.err[0], self.err[2]], "sum")
            err = np.array([sums[0], dist.allreduce(self.ctx, [self.err[1]], "max")[0], sums[1]])
            s_dofs = int(np.prod([self.k * n + 1 for n in self.n_cells]))
        return dict(cells=int(np.prod(self.n_cells)), s_dofs=s_dofs, t_dofs=self.nb, iterations=self.total_iterations,
                    timesteps=self.n_solves, linf=float(err[1]), l2=math.sqrt(err[0]), h1=math.sqrt(err[2]),
                    levels="".join(self.mg_type_level), tau=self.tau)

    def close(self):
        L = capi.lib()
        if self.ti:
            L.stfem_ti_destroy(self.ti)
            self.ti = C.c_void_p()
        for v in (self.x, self.rhs, self.v, self.prev_x, self.prev_v):
            if v is not None:
                v.free()
        if self.mg is not None:
            self.mg.close()
        for o in [self.matrix, self.rhs_matrix, self.rhs_matrix_v] + self.level_ops:
            if o is not None:
                o.close()
        for m in self.meshes.values():
            m.close()
