"""ctypes binding of the C ABI (include/stfem_b200.h) — the stub a maintainer of a Python
harness would add.  Thin by design: numpy in, numpy out, no torch.

The product path fails loudly when libstfem_b200.so is missing: there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstfem_b200.so")

F64, F32 = 0, 1
MAX_BLOCKS = 16
_NP = {F64: np.float64, F32: np.float32}


class StfemError(RuntimeError):
    pass


class OpDesc(C.Structure):
    _fields_ = [("degree", C.c_int), ("number_type", C.c_int), ("nb_rows", C.c_int), ("nb_cols", C.c_int),
                ("Alpha", C.POINTER(C.c_double)), ("Beta", C.POINTER(C.c_double)),
                ("laplace_coeff_cell", C.POINTER(C.c_double)), ("laplace_coeff_q", C.POINTER(C.c_double)),
                ("kernel_variant", C.c_int)]


_lib = None

# every symbol include/stfem_b200.h declares: (restype, argtypes)
_vp, _vpp, _dp = C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_double)
SYMBOLS = {
    "stfem_last_error": (C.c_char_p, []),
    "stfem_version": (C.c_char_p, []),
    "stfem_ctx_create": (C.c_int, [C.c_int, _vpp]),
    "stfem_ctx_destroy": (C.c_int, [_vp]),
    "stfem_ctx_synchronize": (C.c_int, [_vp]),
    "stfem_ctx_stream": (_vp, [_vp]),
    "stfem_ctx_launch_count": (C.c_longlong, [_vp]),
    "stfem_ctx_timer_start": (C.c_int, [_vp]),
    "stfem_ctx_timer_stop": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "stfem_dev_alloc": (C.c_int, [_vp, C.c_size_t, _vpp]),
    "stfem_dev_free": (C.c_int, [_vp, _vp]),
    "stfem_dev_upload": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "stfem_dev_download": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "stfem_dev_memset": (C.c_int, [_vp, _vp, C.c_int, C.c_size_t]),
    "stfem_host_alloc_pinned": (C.c_int, [C.c_size_t, _vpp]),
    "stfem_host_free_pinned": (C.c_int, [_vp]),
    "stfem_mesh_create": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.c_uint, _vpp]),
    "stfem_mesh_destroy": (C.c_int, [_vp]),
    "stfem_op_create": (C.c_int, [_vp, C.POINTER(OpDesc), _vpp]),
    "stfem_op_destroy": (C.c_int, [_vp]),
    "stfem_op_n_dofs_per_block": (C.c_longlong, [_vp]),
    "stfem_op_n_blocks": (C.c_int, [_vp]),
    "stfem_op_vmult": (C.c_int, [_vp, _vpp, _vpp, C.c_int]),
    "stfem_op_vmult_slice_add": (C.c_int, [_vp, _vpp, _vp]),
    "stfem_op_diagonal": (C.c_int, [_vp, _vpp]),
    "stfem_op_vmult_host": (C.c_int, [_vp, _vpp, _vpp, C.c_int]),
    "stfem_op_host_copy_floor": (C.c_int, [_vp, _vpp, _vpp, C.c_int, C.POINTER(C.c_double)]),
    "stfem_op_set_timing": (C.c_int, [_vp, C.c_int]),
    "stfem_op_last_kernel_ms": (C.c_float, [_vp]),
    "stfem_mg_create": (C.c_int, [_vp, _vp, _vpp]),
    "stfem_mg_destroy": (C.c_int, [_vp]),
    "stfem_mg_n_levels": (C.c_int, [_vp]),
    "stfem_mg_vmult": (C.c_int, [_vp, _vpp, _vpp]),
    "stfem_mg_level_apply": (C.c_int, [_vp, C.c_int, C.c_int, _vpp, _vpp]),
    "stfem_mg_level_info": (C.c_int, [_vp, C.c_int, _dp]),
    "stfem_solver_create": (C.c_int, [_vpp]),
    "stfem_solver_destroy": (C.c_int, [_vp]),
    "stfem_fgmres_solve": (C.c_int, [_vp, _vp, _vp, _vpp, _vpp, C.c_int, C.c_int, C.c_double, C.c_double,
                                     C.POINTER(C.c_int), _dp, _dp]),
    "stfem_fe_time_n_blocks": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "stfem_fe_time_weights": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_int] + [_dp] * 4),
    "stfem_fe_time_weights_wave": (C.c_int, [C.c_int, C.c_int] + [_dp] * 4 + [C.c_int] + [_dp] * 5),
    "stfem_time_transfer_matrix": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int,
                                             C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "stfem_poly_mg_sequence": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]),
    "stfem_mg_sequence": (C.c_int, [C.c_int] * 5 + [C.c_char] + [C.c_int] * 4 + [C.c_char_p, C.c_int]),
    "stfem_precondition_stmg_types": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "stfem_quadrature_rule": (C.c_int, [C.c_int, C.c_int, _dp, _dp]),
    "stfem_cart_fd_modes": (C.c_int, [C.c_int, _dp, _dp]),
    "stfem_coefficient_distortion": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.c_double, _dp]),
    "stfem_coefficient_at_qpoints": (C.c_int, [C.c_int, C.POINTER(C.c_int), _dp, _dp, _dp, C.c_int, C.POINTER(C.c_int), _dp, _dp,
                                               C.c_double, _dp, _dp]),
    "stfem_cutoff_cinfty_interpolate": (C.c_int, [C.c_int, C.POINTER(C.c_int), _dp, _dp, _dp, C.c_int, C.c_double, _dp, C.c_int,
                                                  _dp]),
}


def lib():
    """Load the shared library (once).  Raises if it has not been built — no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise StfemError("%s not found: run `python dealii-stfem_b200/build.py` (no CPU fallback exists)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise StfemError("stfem error %d: %s" % (rc, lib().stfem_last_error().decode()))


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Context:
    def __init__(self, device=0):
        self.h = C.c_void_p()
        check(lib().stfem_ctx_create(device, C.byref(self.h)))

    def synchronize(self):
        check(lib().stfem_ctx_synchronize(self.h))

    @property
    def stream(self):
        return lib().stfem_ctx_stream(self.h)

    @property
    def launches(self):
        return lib().stfem_ctx_launch_count(self.h)

    def timer_start(self):
        check(lib().stfem_ctx_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        check(lib().stfem_ctx_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def close(self):
        if self.h:
            lib().stfem_ctx_destroy(self.h)
            self.h = C.c_void_p()


def pinned_array(shape, dtype):
    """numpy array backed by cudaMallocHost memory (kept alive by the returned array's base)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    check(lib().stfem_host_alloc_pinned(n, C.byref(p)))
    buf = (C.c_byte * n).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr, p


def free_pinned(p):
    lib().stfem_host_free_pinned(p)


class DeviceBlockVector:
    """nb separate device arrays of N numbers (reference BlockVectorT, include/types.h:20-23)."""

    def __init__(self, ctx, nb, n, number_type=F64):
        self.ctx, self.nb, self.n, self.number_type = ctx, nb, n, number_type
        self.dtype = _NP[number_type]
        self.itemsize = np.dtype(self.dtype).itemsize
        self.ptrs = (C.c_void_p * nb)()
        for b in range(nb):
            p = C.c_void_p()
            check(lib().stfem_dev_alloc(ctx.h, n * self.itemsize, C.byref(p)))
            self.ptrs[b] = p.value

    def upload(self, arr):
        arr = np.ascontiguousarray(arr, dtype=self.dtype).reshape(self.nb, self.n)
        for b in range(self.nb):
            check(lib().stfem_dev_upload(self.ctx.h, self.ptrs[b], arr[b].ctypes.data, self.n * self.itemsize))
        return self

    def download(self):
        out = np.empty((self.nb, self.n), dtype=self.dtype)
        for b in range(self.nb):
            check(lib().stfem_dev_download(self.ctx.h, out[b].ctypes.data, self.ptrs[b], self.n * self.itemsize))
        return out

    def zero(self):
        for b in range(self.nb):
            check(lib().stfem_dev_memset(self.ctx.h, self.ptrs[b], 0, self.n * self.itemsize))

    def free(self):
        for b in range(self.nb):
            if self.ptrs[b]:
                lib().stfem_dev_free(self.ctx.h, self.ptrs[b])
                self.ptrs[b] = None


class Mesh:
    def __init__(self, ctx, n_cells, lower=None, upper=None, vertices=None, dirichlet_faces=None):
        self.ctx = ctx
        self.dim = len(n_cells)
        self.n_cells = [int(v) for v in n_cells]
        n = (C.c_int * self.dim)(*self.n_cells)
        lo = np.zeros(self.dim) if lower is None else np.asarray(lower, np.float64)
        up = np.ones(self.dim) if upper is None else np.asarray(upper, np.float64)
        vp = None
        if vertices is not None:
            self._v = np.ascontiguousarray(vertices, np.float64)
            vp = _dptr(self._v)
        mask = (0x3f if self.dim == 3 else 0xf) if dirichlet_faces is None else dirichlet_faces
        self.h = C.c_void_p()
        check(lib().stfem_mesh_create(ctx.h, self.dim, n, _dptr(lo), _dptr(up), vp, mask, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().stfem_mesh_destroy(self.h)
            self.h = C.c_void_p()


class Operator:
    """SystemMatrix<dim, Number, MatrixFreeOperatorScalar> of the reference (operators.h:516-663)."""

    def __init__(self, mesh, degree, Alpha, Beta, number_type=F64, laplace_coeff_cell=None, laplace_coeff_q=None,
                 variant=0):
        self.mesh, self.ctx, self.number_type = mesh, mesh.ctx, number_type
        A = np.ascontiguousarray(np.atleast_2d(Alpha), np.float64)
        B = np.ascontiguousarray(np.atleast_2d(Beta), np.float64)
        assert A.shape == B.shape
        d = OpDesc()
        d.degree, d.number_type, d.nb_rows, d.nb_cols = degree, number_type, A.shape[0], A.shape[1]
        d.Alpha, d.Beta = _dptr(A), _dptr(B)
        if laplace_coeff_cell is not None:
            self._coef = np.ascontiguousarray(laplace_coeff_cell, np.float64)
            d.laplace_coeff_cell = _dptr(self._coef)
        if laplace_coeff_q is not None:
            self._coefq = np.ascontiguousarray(laplace_coeff_q, np.float64)
            d.laplace_coeff_q = _dptr(self._coefq)
        d.kernel_variant = variant
        self.h = C.c_void_p()
        check(lib().stfem_op_create(mesh.h, C.byref(d), C.byref(self.h)))
        self.nb_rows, self.nb_cols = A.shape
        self.n = lib().stfem_op_n_dofs_per_block(self.h)

    def new_vector(self, nb=None):
        return DeviceBlockVector(self.ctx, self.nb_rows if nb is None else nb, self.n, self.number_type)

    def vmult(self, dst, src, transpose=False):
        check(lib().stfem_op_vmult(self.h, dst.ptrs, src.ptrs, 1 if transpose else 0))

    def Tvmult(self, dst, src):
        self.vmult(dst, src, True)

    def vmult_slice_add(self, dst, src0):
        check(lib().stfem_op_vmult_slice_add(self.h, dst.ptrs, src0.ptrs[0]))

    def diagonal(self, dst):
        """get_matrix_diagonal (operators.h:613-625) into a block vector of the operator's number type."""
        check(lib().stfem_op_diagonal(self.h, dst.ptrs))

    def host_copy_floor(self, dst, src, reps=3):
        """Milliseconds per round of uploading src and downloading dst concurrently without a kernel (stfem_op_host_copy_floor)."""
        nb = self.nb_rows
        dp = (C.c_void_p * nb)(*[dst[b].ctypes.data for b in range(nb)])
        spp = (C.c_void_p * nb)(*[src[b].ctypes.data for b in range(nb)])
        ms = C.c_double()
        check(lib().stfem_op_host_copy_floor(self.h, dp, spp, int(reps), C.byref(ms)))
        return ms.value

    def vmult_host(self, dst, src, transpose=False):
        """dst, src: numpy [nb, N] (C-contiguous rows; pinned or pageable)."""
        nb = self.nb_rows
        dp = (C.c_void_p * nb)(*[dst[b].ctypes.data for b in range(nb)])
        spp = (C.c_void_p * nb)(*[src[b].ctypes.data for b in range(nb)])
        check(lib().stfem_op_vmult_host(self.h, dp, spp, 1 if transpose else 0))

    def set_timing(self, on=True):
        check(lib().stfem_op_set_timing(self.h, 1 if on else 0))

    def last_kernel_ms(self):
        return lib().stfem_op_last_kernel_ms(self.h)

    def close(self):
        if self.h:
            lib().stfem_op_destroy(self.h)
            self.h = C.c_void_p()


class MgDesc(C.Structure):
    _fields_ = [("n_levels", C.c_int), ("level_ops", C.POINTER(C.c_void_p)), ("mg_type_level", C.c_char_p),
                ("smoother_types", C.POINTER(C.c_int)), ("time_type", C.c_int), ("n_timesteps_at_once", C.c_int),
                ("poly_time_sequence", C.POINTER(C.c_int)), ("n_poly_time", C.c_int), ("smoothing_steps", C.c_int),
                ("relaxation", C.c_double), ("smoothing_range", C.c_double), ("eig_n_iterations", C.c_int),
                ("variable", C.c_int), ("restrict_is_transpose_prolongate", C.c_int), ("inner_preconditioner", C.c_int),
                ("coarse_grid_maxiter", C.c_int), ("coarse_grid_abstol", C.c_double), ("vanka_storage", C.c_int)]


class Multigrid:
    """GMG of the reference (include/stmg.h:1047-1344) on a list of level Operators (coarse -> fine)."""

    def __init__(self, ctx, level_ops, mg_type_level, smoother_types, time_type, n_timesteps_at_once, poly_time_sequence,
                 smoothing_steps=1, relaxation=0.0, smoothing_range=1.0, eig_n_iterations=20, variable=True,
                 restrict_is_transpose_prolongate=True, inner_preconditioner="vanka", vanka_storage="level",
                 coarse_grid_maxiter=0, coarse_grid_abstol=1e-20):
        """inner_preconditioner: "vanka" (PreconditionVanka, the reference) or "jacobi" (point-Jacobi, diagonal inverse).
        vanka_storage: "level" (patch inverses in the level precision) or "half" (FP16, dense patches only).
        coarse_grid_maxiter > 0: GMRES on the coarsest level (coarseGridSmootherType != "Smoother", stmg.h:1240-1302)."""
        self.ctx, self.ops = ctx, list(level_ops)
        nl = len(self.ops)
        d = MgDesc()
        d.n_levels = nl
        self._ops = (C.c_void_p * nl)(*[o.h.value for o in self.ops])
        d.level_ops = self._ops
        self._types = "".join(mg_type_level).encode()
        d.mg_type_level = self._types
        self._sm = (C.c_int * nl)(*smoother_types)
        d.smoother_types = self._sm
        d.time_type = {"CGP": 1, "DG": 2}[time_type]
        d.n_timesteps_at_once = n_timesteps_at_once
        self._poly = (C.c_int * len(poly_time_sequence))(*poly_time_sequence)
        d.poly_time_sequence, d.n_poly_time = self._poly, len(poly_time_sequence)
        d.smoothing_steps, d.relaxation, d.smoothing_range = smoothing_steps, relaxation, smoothing_range
        d.eig_n_iterations, d.variable = eig_n_iterations, int(variable)
        d.restrict_is_transpose_prolongate = int(restrict_is_transpose_prolongate)
        d.inner_preconditioner = {"vanka": 0, "jacobi": 1}[inner_preconditioner]
        d.vanka_storage = {"level": 0, "half": 1}[vanka_storage]
        d.coarse_grid_maxiter, d.coarse_grid_abstol = int(coarse_grid_maxiter), float(coarse_grid_abstol)
        self.h = C.c_void_p()
        check(lib().stfem_mg_create(ctx.h, C.byref(d), C.byref(self.h)))

    @property
    def n_levels(self):
        return lib().stfem_mg_n_levels(self.h)

    def vmult(self, dst, src):
        check(lib().stfem_mg_vmult(self.h, dst.ptrs, src.ptrs))

    def level_apply(self, level, what, dst, src):
        check(lib().stfem_mg_level_apply(self.h, level, what, dst.ptrs, src.ptrs))

    def level_info(self, level):
        out = np.zeros(10)
        check(lib().stfem_mg_level_info(self.h, level, _dptr(out)))
        keys = ["smoother", "lambda", "omega", "theta", "delta", "steps", "N", "blocks", "patch_matrices", "patch_bytes"]
        return dict(zip(keys, out))

    def close(self):
        if self.h:
            lib().stfem_mg_destroy(self.h)
            self.h = C.c_void_p()


class Fgmres:
    """SolverFGMRES + ReductionControl as configured by the reference (time_integrators.h:56-59)."""

    def __init__(self):
        self.h = C.c_void_p()
        check(lib().stfem_solver_create(C.byref(self.h)))

    def solve(self, A, x, b, M=None, max_basis=100, max_iter=200, abstol=1e-12, reduce=1e-12):
        it, r0, r1 = C.c_int(), C.c_double(), C.c_double()
        rc = lib().stfem_fgmres_solve(self.h, A.h, M.h if M is not None else None, x.ptrs, b.ptrs, max_basis, max_iter,
                                      abstol, reduce, C.byref(it), C.byref(r0), C.byref(r1))
        self.iterations, self.initial_residual, self.final_residual = it.value, r0.value, r1.value
        check(rc)
        return it.value

    def close(self):
        if self.h:
            lib().stfem_solver_destroy(self.h)
            self.h = C.c_void_p()
