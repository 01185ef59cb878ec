"""ORACLE (test infrastructure only).  ctypes wrapper of oracle/cpu_ref.cpp — the C++/OpenMP
restatement of the reference's unfused SystemMatrix::vmult (operators.h:536-559) used as
`cpu_baseline` in bench.py and as a second, faster oracle for larger parity cases."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libstfem_cpu_ref.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "cpu_ref.cpp")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        # -march=native code must be rebuilt on the machine that runs it
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.oracle_max_threads.restype = C.c_int
        _lib.oracle_system_vmult.restype = C.c_int
    return _lib


def max_threads():
    return lib().oracle_max_threads()


def system_vmult(space, Alpha, Beta, src, transpose=False, coeff_q=None, n_threads=0, general=None):
    """space: oracle.spatial.Space; src [nb, N] float64 -> dst [nb, N]."""
    L = lib()
    mesh = space.mesh
    dim = mesh.dim
    nb = src.shape[0]
    A = np.ascontiguousarray(Alpha, np.float64)
    B = np.ascontiguousarray(Beta, np.float64)
    S = np.ascontiguousarray(space.S)
    D = np.ascontiguousarray(space.D)
    w = np.ascontiguousarray(space.wq)
    h = np.array([(mesh.upper[d] - mesh.lower[d]) / mesh.n[d] for d in range(dim)])
    use_metric = (not mesh.cartesian) if general is None else general
    metric = None
    if use_metric:
        G, JxW, _ = space.geometry()
        iu = np.triu_indices(dim)
        metric = np.ascontiguousarray(np.concatenate([G[:, :, iu[0], iu[1]], JxW[..., None]], axis=-1))
    cq = None if coeff_q is None else np.ascontiguousarray(coeff_q, np.float64)
    src = np.ascontiguousarray(src, np.float64)
    dst = np.zeros_like(src)
    n = (C.c_int * dim)(*mesh.n)
    sp = (C.c_void_p * nb)(*[src[b].ctypes.data for b in range(nb)])
    dp = (C.c_void_p * nb)(*[dst[b].ctypes.data for b in range(nb)])
    mask = 0x3f if dim == 3 else 0xf

    def p(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    rc = L.oracle_system_vmult(C.c_int(dim), n, C.c_int(space.k), C.c_int(nb), p(A), p(B), C.c_int(int(transpose)),
                               p(S), p(D), p(w), p(h), p(metric), p(cq), C.c_uint(mask), sp, dp, C.c_int(n_threads))
    assert rc == 0
    return dst
