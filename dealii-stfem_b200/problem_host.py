"""Host-side problem data of the reference's practical runs (spaceTimeConvergenceTest = false): thin numpy wrappers of
the C ABI's set-up helpers (csrc/capi_problem.cu).  No GPU work, no oracle."""
import ctypes as C

import numpy as np

from . import capi


def _ints(v):
    return (C.c_int * len(v))(*[int(x) for x in v])


def _dbl(v):
    return None if v is None else np.ascontiguousarray(v, np.float64)


def _ptr(a):
    return None if a is None else capi._dptr(a)


def coefficient_distortion(subdivisions, distort_coeff):
    """Coefficient<dim>::distortion (include/operators.h:905-921), shape = subdivisions."""
    out = np.empty(int(np.prod(subdivisions)), np.float64)
    capi.check(capi.lib().stfem_coefficient_distortion(len(subdivisions), _ints(subdivisions), float(distort_coeff), _ptr(out)))
    return out.reshape([int(s) for s in subdivisions])


def coefficient_at_qpoints(n_cells, lower, upper, degree, subdivisions, coeff_lower, coeff_upper, distort_coeff, vertices=None,
                           c123=None):
    """evaluate_coefficient(Coefficient<dim>) (operators.h:1060-1087): [n_cells_total, (degree+1)^dim]."""
    dim = len(n_cells)
    lo, up, v = _dbl(lower), _dbl(upper), _dbl(vertices)
    clo, cup, c = _dbl(coeff_lower), _dbl(coeff_upper), _dbl(c123)
    out = np.empty((int(np.prod(n_cells)), (degree + 1) ** dim), np.float64)
    capi.check(capi.lib().stfem_coefficient_at_qpoints(dim, _ints(n_cells), _ptr(lo), _ptr(up), _ptr(v), degree, _ints(subdivisions),
                                                       _ptr(clo), _ptr(cup), float(distort_coeff), _ptr(c), _ptr(out)))
    return out


def cutoff_cinfty_interpolate(n_cells, lower, upper, degree, center, radius=1.0e-2, integrate_to_one=True, vertices=None):
    """Nodal values of Functions::CutOffFunctionCinfty (tests/tp_01.cc:376-378, 551), lexicographic."""
    dim = len(n_cells)
    lo, up, v, ce = _dbl(lower), _dbl(upper), _dbl(vertices), _dbl(center)
    out = np.empty(int(np.prod([degree * n + 1 for n in n_cells])), np.float64)
    capi.check(capi.lib().stfem_cutoff_cinfty_interpolate(dim, _ints(n_cells), _ptr(lo), _ptr(up), _ptr(v), degree, float(radius),
                                                          _ptr(ce), int(bool(integrate_to_one)), _ptr(out)))
    return out
