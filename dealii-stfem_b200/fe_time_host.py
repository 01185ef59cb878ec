"""Python view of the product's host-side time algebra (csrc/fe_time.hpp through the C ABI).
Same call shapes as the reference functions of include/fe_time.h / fe_time.cc."""
import ctypes as C

import numpy as np

from . import capi

CGP, DG = "CGP", "DG"
_T = {"CGP": 1, "DG": 2, 1: 1, 2: 2}
_P = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_P)


def n_blocks(ttype, r, nts=1):
    return capi.lib().stfem_fe_time_n_blocks(_T[ttype], r, nts)


def get_fe_time_weights(ttype, r, tau, n_timesteps_at_once=1):
    nb = n_blocks(ttype, r, n_timesteps_at_once)
    A, B = np.zeros((nb, nb)), np.zeros((nb, nb))
    G, Z = np.zeros((nb, 1)), np.zeros((nb, 1))
    capi.check(capi.lib().stfem_fe_time_weights(_T[ttype], r, C.c_double(tau), n_timesteps_at_once, _p(A), _p(B), _p(G), _p(Z)))
    return [A, B, G, Z]


def get_fe_time_weights_wave(ttype, Alpha, Beta, Gamma, Zeta, n_timesteps_at_once=1):
    A = np.ascontiguousarray(Alpha, np.float64)
    B = np.ascontiguousarray(Beta, np.float64)
    G = np.ascontiguousarray(Gamma, np.float64)
    Z = np.ascontiguousarray(Zeta, np.float64) if Zeta is not None and np.size(Zeta) else np.zeros((A.shape[0], 1))
    nd = A.shape[0]
    nt = nd * n_timesteps_at_once
    out = [np.zeros((nt, nt)), np.zeros((nt, nt)), np.zeros((nt, 1)), np.zeros((nt, 1)), np.zeros((nt, 1))]
    capi.check(capi.lib().stfem_fe_time_weights_wave(_T[ttype], nd, _p(A), _p(B), _p(G), _p(Z), n_timesteps_at_once,
                                                     *[_p(o) for o in out]))
    return out


def _transfer(kind, ttype, a, b, nts):
    rows, cols = C.c_int(), C.c_int()
    capi.check(capi.lib().stfem_time_transfer_matrix(kind, _T[ttype], a, b, nts, None, 0, C.byref(rows), C.byref(cols)))
    M = np.zeros((rows.value, cols.value))
    capi.check(capi.lib().stfem_time_transfer_matrix(kind, _T[ttype], a, b, nts, _p(M), M.size, C.byref(rows), C.byref(cols)))
    return M


def get_time_projection_matrix(ttype, r_src, r_dst, n_timesteps_at_once):
    return _transfer(0, ttype, r_src, r_dst, n_timesteps_at_once)


def get_time_prolongation_matrix(ttype, r, n_timesteps_at_once=2):
    return _transfer(1, ttype, r, 0, n_timesteps_at_once)


def get_time_restriction_matrix(ttype, r, n_timesteps_at_once=2):
    return _transfer(2, ttype, r, 0, n_timesteps_at_once)


def get_poly_mg_sequence(k_max, k_min, p_seq="bisect"):
    code = {"bisect": 0, "decrease_by_one": 1, "go_to_one": 2}[p_seq]
    out = (C.c_int * 64)()
    cnt = C.c_int()
    capi.check(capi.lib().stfem_poly_mg_sequence(k_max, k_min, code, out, 64, C.byref(cnt)))
    return [out[i] for i in range(cnt.value)]


def get_mg_sequence(n_sp_lvl, k_seq, p_seq, n_timesteps_at_once, n_timesteps_at_once_min=1, lower_lvl="k",
                    coarsening_type="space_and_time", time_before_space=False, use_p_multigrid_space=False,
                    zip_from_back=True):
    lower = {"tau": "t"}.get(lower_lvl, lower_lvl)
    buf = C.create_string_buffer(256)
    capi.check(capi.lib().stfem_mg_sequence(n_sp_lvl, len(k_seq), len(p_seq), n_timesteps_at_once, n_timesteps_at_once_min,
                                            C.c_char(lower.encode()), 0 if coarsening_type == "space_or_time" else 1,
                                            int(time_before_space), int(use_p_multigrid_space), int(zip_from_back), buf, 256))
    return list(buf.value.decode())


def get_precondition_stmg_types(mg_type_level, coarsening_type, time_before_space, zip_from_back=True, smoother=1):
    seq = "".join(mg_type_level).encode()
    out = (C.c_int * (len(mg_type_level) + 1))()
    capi.check(capi.lib().stfem_precondition_stmg_types(seq, 0 if coarsening_type == "space_or_time" else 1,
                                                        int(time_before_space), smoother, out))
    return list(out)


def quadrature_rule(kind, n):
    x, w = np.zeros(n), np.zeros(n)
    capi.check(capi.lib().stfem_quadrature_rule({"gauss": 0, "lobatto": 1, "radau": 2}[kind], n, _p(x), _p(w)))
    return x, w


def get_time_evaluation_matrix(ttype, r, samples_per_interval):
    """get_time_evaluation_matrix(get_time_basis(type, r), samples) (fe_time.h:307-326, fe_time.cc:152-179): the Lagrange
    basis on the GLL(r+1) (CGP) / right Radau(r+1) (DG) points of [0,1] at samples_per_interval equidistant times."""
    nodes = quadrature_rule("lobatto" if ttype == CGP else "radau", r + 1)[0]
    t = np.arange(samples_per_interval) / max(samples_per_interval - 1.0, 1.0)
    M = np.ones((samples_per_interval, r + 1))
    for j in range(r + 1):
        for m in range(r + 1):
            if m != j:
                M[:, j] *= (t - nodes[m]) / (nodes[j] - nodes[m])
    return M
