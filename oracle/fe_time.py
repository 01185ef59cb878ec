"""ORACLE (test infrastructure only — never imported by the product path).

numpy restatement of the reference's time-discretisation algebra,
include/fe_time.h and include/fe_time.cc of immaaane/dealii-stfem.
Every function cites the reference lines it follows.  deal.II pieces the
reference calls (QGauss*, Lagrange bases, FETools::get_projection_matrix,
FE_Q/FE_DGQArbitraryNodes prolongation/restriction matrices) are restated from
their published definitions; they are pinned by the
reference's own golden outputs tests/tp_02.output and tests/transfer_02.output
(tests/golden/*.json, see tests/test_oracle_fe_time.py).
"""
import math

import numpy as np

from . import quadrature as Q

CGP, DG = "CGP", "DG"


# ----------------------------------------------------------------------------- bases
def time_nodes(ttype, r):
    """get_time_quad: fe_time.cc:152-161 (DG: Radau-right(r+1), CGP: Lobatto(r+1))."""
    if ttype == DG:
        return Q.gauss_radau_right(r + 1)[0]
    return Q.gauss_lobatto(r + 1)[0]


def get_cg_weights(r):
    """fe_time.h:643-696. Returns (matrix, matrix_der), each r x (r+1)."""
    trial = Q.gauss_lobatto(r + 1)[0]
    test = trial[1:]
    xq, wq = Q.gauss(r + 2)
    Lt = Q.lagrange_eval(trial, xq)
    dLt = Q.lagrange_deriv(trial, xq)
    Ls = Q.lagrange_eval(test, xq)
    matrix = np.einsum("q,iq,jq->ij", wq, Ls, Lt)
    matrix_der = np.einsum("q,iq,jq->ij", wq, Ls, dLt)
    return matrix, matrix_der


def get_dg_weights(r):
    """fe_time.h:698-744. Returns (mass, derivative+jump, jump vector (r+1)x1)."""
    nodes = Q.gauss_radau_right(r + 1)[0]
    xq, wq = Q.gauss(r + 2)
    Lv = Q.lagrange_eval(nodes, xq)
    dLv = Q.lagrange_deriv(nodes, xq)
    L0 = Q.lagrange_eval(nodes, [0.0])[:, 0]
    lhs = np.einsum("q,iq,jq->ij", wq, Lv, Lv)
    lhs_der = np.outer(L0, L0) + np.einsum("q,iq,jq->ij", wq, Lv, dLv)
    jump = L0.reshape(-1, 1).copy()
    return lhs, lhs_der, jump


def split_lhs_rhs_cg(w):
    """fe_time.h:485-503."""
    m, md = w
    return [m[:, 1:].copy(), md[:, 1:].copy(), -m[:, 0:1].copy(), -md[:, 0:1].copy()]


def split_lhs_rhs_dg(w):
    """fe_time.h:505-514."""
    return [w[0].copy(), w[1].copy(), w[2].copy(), np.zeros((w[2].shape[0], 1))]


def get_fe_time_weights(ttype, r, tau, n_timesteps_at_once=1):
    """fe_time.h:351-409. Returns [Alpha*tau, Beta, Gamma, Zeta] for the multi-step system."""
    if ttype == CGP:
        tmp = split_lhs_rhs_cg(get_cg_weights(r))
        tmp[2] = tmp[2] * tau
    else:
        tmp = split_lhs_rhs_dg(get_dg_weights(r))
        tmp[3] = tmp[2]
        tmp[2] = np.zeros_like(tmp[3])
    tmp[0] = tmp[0] * tau
    nd = tmp[0].shape[0]
    nt = nd * n_timesteps_at_once
    ret = [np.zeros((nt, nt)), np.zeros((nt, nt)), np.zeros((nt, 1)), np.zeros((nt, 1))]
    for it in range(n_timesteps_at_once):
        for i in range(nd):
            if it < n_timesteps_at_once - 1 and i == nd - 1:
                for j in range(nd):
                    ret[0][j + (it + 1) * nd, i + it * nd] = -tmp[2][j, 0]
                    ret[1][j + (it + 1) * nd, i + it * nd] = -tmp[3][j, 0]
            for j in range(nd):
                ret[0][i + it * nd, j + it * nd] = tmp[0][i, j]
                ret[1][i + it * nd, j + it * nd] = tmp[1][i, j]
    for i in range(nd):
        ret[2][i, 0] = tmp[2 if ttype == CGP else 3][i, 0]
        ret[3][i, 0] = tmp[3 if ttype == CGP else 2][i, 0]
    return ret


def get_fe_time_weights_wave(ttype, Alpha, Beta, Gamma, Zeta, n_timesteps_at_once=1):
    """fe_time.h:157-305: second-order elimination for the wave equation.

    Returns [lhs_uK, lhs_uM, rhs_uK, rhs_uM, rhs_vM]."""
    Ainv = np.linalg.inv(Alpha)
    BxAixB = Beta @ Ainv @ Beta
    BxAixG = Beta @ Ainv @ Gamma
    m = Gamma.shape[0]
    gxai = Gamma[m - 1, 0] / Alpha[m - 1, m - 1]
    GxAixG = Gamma * gxai
    Beta_row = Beta[m - 1:m, :]
    GxAixB = (Gamma @ Beta_row) / Alpha[m - 1, m - 1]
    nd = Alpha.shape[0]
    nt = nd * n_timesteps_at_once
    ret = [np.zeros((nt, nt)), np.zeros((nt, nt)), np.zeros((nt, 1)), np.zeros((nt, 1)), np.zeros((nt, 1))]
    if ttype == CGP:
        BxAixZ = Beta @ Ainv @ Zeta
        ZmBxAixG = Zeta - BxAixG
        ZmBxAixB = (ZmBxAixG @ Beta_row) / Alpha[m - 1, m - 1]
        zxai = Zeta[m - 1, 0] / Alpha[m - 1, m - 1]
        for it in range(n_timesteps_at_once):
            for jt in range(it + 1):
                for i in range(nd):
                    if it == 0 and jt == 0:
                        ret[2][i, 0] = Gamma[i, 0]
                        ret[3][i, 0] = BxAixZ[i, 0]
                        ret[4][i, 0] = ZmBxAixG[i, 0]
                    elif jt == 0:
                        ret[3][i + it * nd, 0] = -zxai * gxai ** (it - 1) * ZmBxAixG[i, 0]
                        ret[4][i + it * nd, 0] = gxai ** it * ZmBxAixG[i, 0]
                    if it == jt + 1:
                        ret[0][i + it * nd, nd - 1 + jt * nd] = -Gamma[i, 0]
                        ret[1][i + it * nd, nd - 1 + jt * nd] = -BxAixZ[i, 0]
                    if it == jt:
                        for j in range(nd):
                            ret[0][i + it * nd, j + it * nd] = Alpha[i, j]
                            ret[1][i + it * nd, j + it * nd] = BxAixB[i, j]
                    else:
                        for j in range(nd):
                            extra = 0.0
                            if it > 1 and it - 1 > jt and j == nd - 1:
                                extra = gxai ** (it - jt - 2) * zxai * ZmBxAixG[i, 0]
                            ret[1][i + it * nd, j + jt * nd] += -gxai ** (it - jt - 1) * ZmBxAixB[i, j] + extra
    else:
        for it in range(n_timesteps_at_once):
            for i in range(nd):
                if it == 0:
                    ret[3][i, 0] = BxAixG[i, 0]
                    ret[4][i, 0] = Gamma[i, 0]
                if it == 1:
                    ret[3][nd + i, 0] = -GxAixG[i, 0]
                if it < n_timesteps_at_once - 1:
                    for j in range(nd):
                        ret[1][j + (it + 1) * nd, i + it * nd] = \
                            -GxAixB[j, i] - (BxAixG[j, 0] if i == nd - 1 else 0.0)
                if it < n_timesteps_at_once - 2 and i == nd - 1:
                    for j in range(nd):
                        ret[1][j + (it + 2) * nd, i + it * nd] = GxAixG[j, 0]
                for j in range(nd):
                    ret[0][i + it * nd, j + it * nd] = Alpha[i, j]
                    ret[1][i + it * nd, j + it * nd] = BxAixB[i, j]
    return ret


def get_fe_time_weights_levels(ttype, tau, n_timesteps_at_once, mg_type_level, poly_time_sequence):
    """fe_time.h:411-442: weights per multigrid level (coarse -> fine order like the reference)."""
    out = [None] * (len(mg_type_level) + 1)
    p = len(poly_time_sequence) - 1
    nts = n_timesteps_at_once
    out[-1] = get_fe_time_weights(ttype, poly_time_sequence[p], tau, nts)
    idx = len(out) - 2
    for mgt in reversed(mg_type_level):
        if mgt == "k":
            p -= 1
        elif mgt == "t":
            nts //= 2
            tau *= 2
        out[idx] = get_fe_time_weights(ttype, poly_time_sequence[p], tau, nts)
        idx -= 1
    return out


def get_fe_time_weights_wave_levels(ttype, tau, n_timesteps_at_once, mg_type_level, poly_time_sequence):
    """fe_time.h:444-474.  NOTE (faithful to the reference): the per-level wave weights are
    built from the *multi-step* heat weights with n_timesteps_at_once=1 passed to the
    elimination, i.e. get_fe_time_weights_wave(type, A, B, G, Z) on the full nb x nb matrices."""
    tw = get_fe_time_weights_levels(ttype, tau, n_timesteps_at_once, mg_type_level, poly_time_sequence)
    return [get_fe_time_weights_wave(ttype, w[0], w[1], w[2], w[3]) for w in tw]


# ----------------------------------------------------------------------------- time transfer
def _l2_projection(nodes_src, nodes_dst):
    """FETools::get_projection_matrix(fe_src, fe_dst): M_dst^{-1} (phi_dst, phi_src)."""
    n = max(len(nodes_src), len(nodes_dst)) + 1
    xq, wq = Q.gauss(n)
    Ls = Q.lagrange_eval(nodes_src, xq)
    Ld = Q.lagrange_eval(nodes_dst, xq)
    M = np.einsum("q,iq,jq->ij", wq, Ld, Ld)
    B = np.einsum("q,iq,jq->ij", wq, Ld, Ls)
    return np.linalg.solve(M, B)


def _fill(dst, src, di, dj, si, sj):
    """FullMatrix::fill(src, dst_offset_i, dst_offset_j, src_offset_i, src_offset_j)."""
    rows = min(dst.shape[0] - di, src.shape[0] - si)
    cols = min(dst.shape[1] - dj, src.shape[1] - sj)
    if rows > 0 and cols > 0:
        dst[di:di + rows, dj:dj + cols] = src[si:si + rows, sj:sj + cols]


def get_time_projection_matrix(ttype, r_src, r_dst, n_timesteps_at_once):
    """fe_time.h:749-805 (k-transfer between time degrees)."""
    nd_dst = r_dst + 1 if ttype == DG else r_dst
    nd_src = r_src + 1 if ttype == DG else r_src
    n_dst = n_timesteps_at_once * (r_dst + 1) if ttype == DG else n_timesteps_at_once * r_dst + 1
    n_src = n_timesteps_at_once * (r_src + 1) if ttype == DG else n_timesteps_at_once * r_src + 1
    proj = _l2_projection(time_nodes(ttype, r_src), time_nodes(ttype, r_dst))
    pn = np.zeros((n_dst, n_src))
    for it in range(n_timesteps_at_once):
        _fill(pn, proj, it * nd_dst, it * nd_src, 0, 0)
    if ttype == CGP:
        return pn[1:, 1:].copy()
    return pn


def _child_embedding(nodes, child):
    """FE::get_prolongation_matrix(child) for a nodal 1D element: P(i,j) = l_j((x_i+child)/2)."""
    x = 0.5 * (np.asarray(nodes) + child)
    return Q.lagrange_eval(nodes, x).T.copy()


def _dg_child_restriction(nodes, child):
    """FE_DGQArbitraryNodes::get_restriction_matrix(child) (FETools::compute_projection_matrices):
    the child's share of the L2 projection fine -> coarse:  M_c^{-1} int_child phi^c_i phi^f_j."""
    n = len(nodes)
    xq, wq = Q.gauss(n + 1)
    Lc_all = Q.lagrange_eval(nodes, xq)
    M = np.einsum("q,iq,jq->ij", wq, Lc_all, Lc_all)
    xc = 0.5 * (xq + child)           # child quadrature points in coarse coordinates
    Lc = Q.lagrange_eval(nodes, xc)
    Lf = Q.lagrange_eval(nodes, xq)   # fine basis in child coordinates
    B = 0.5 * np.einsum("q,iq,jq->ij", wq, Lc, Lf)
    return np.linalg.solve(M, B)


def _q_child_restriction(nodes, child):
    """FE_Q<1>::get_restriction_matrix(child): nodal interpolation; row i is non-zero only if the
    coarse support point i lies inside the child, and then holds the child shape values there."""
    nodes = np.asarray(nodes)
    n = len(nodes)
    R = np.zeros((n, n))
    for i, x in enumerate(nodes):
        xc = 2.0 * x - child
        if -1e-12 <= xc <= 1.0 + 1e-12:
            R[i, :] = Q.lagrange_eval(nodes, [xc])[:, 0]
    R[np.abs(R) < 1e-14] = 0.0
    return R


def get_time_prolongation_matrix(ttype, r, n_timesteps_at_once=2):
    """fe_time.h:807-851 (tau-transfer: one coarse step -> two fine steps)."""
    nodes = time_nodes(ttype, r)
    left, right = _child_embedding(nodes, 0), _child_embedding(nodes, 1)
    if ttype == DG:
        prol = np.zeros((2 * (r + 1), r + 1))
        _fill(prol, left, 0, 0, 0, 0)
        _fill(prol, right, r + 1, 0, 0, 0)
        nd = r + 1
    else:
        prol = np.zeros((2 * r, r))
        _fill(prol, left, 0, 0, 1, 1)
        _fill(prol, right, r, 0, 1, 1)
        nd = r
    pn = np.zeros((nd * n_timesteps_at_once, nd * n_timesteps_at_once // 2))
    for it in range(n_timesteps_at_once // 2):
        _fill(pn, prol, it * 2 * nd, it * nd, 0, 0)
    return pn


def get_time_restriction_matrix(ttype, r, n_timesteps_at_once=2):
    """fe_time.h:853-898."""
    nodes = time_nodes(ttype, r)
    if ttype == DG:
        left, right = _dg_child_restriction(nodes, 0), _dg_child_restriction(nodes, 1)
        rest = np.zeros((r + 1, 2 * (r + 1)))
        _fill(rest, left, 0, 0, 0, 0)
        _fill(rest, right, 0, r + 1, 0, 0)
        nd = r + 1
    else:
        left, right = _q_child_restriction(nodes, 0), _q_child_restriction(nodes, 1)
        rest = np.zeros((r, 2 * r))
        _fill(rest, left, 0, 0, 1, 1)
        _fill(rest, right, 0, r, 1, 1)
        nd = r
    rn = np.zeros((nd * n_timesteps_at_once // 2, nd * n_timesteps_at_once))
    for it in range(n_timesteps_at_once // 2):
        _fill(rn, rest, it * nd, it * 2 * nd, 0, 0)
    return rn


# ----------------------------------------------------------------------------- MG sequences
def create_next_polynomial_coarsening_degree(prev, p_sequence, k_min=0):
    """fe_time.cc:16-38."""
    if p_sequence == "bisect":
        return max(prev // 2, 0)
    if p_sequence == "decrease_by_one":
        return max(prev - 1, 0)
    if p_sequence == "go_to_one":
        return k_min
    raise NotImplementedError(p_sequence)


def get_poly_mg_sequence(k_max, k_min, p_seq="bisect"):
    """fe_time.cc:40-56."""
    degrees = [k_max]
    if degrees[-1] == k_min:
        return degrees
    while degrees[-1] > k_min:
        degrees.append(create_next_polynomial_coarsening_degree(degrees[-1], p_seq, k_min))
    return degrees[::-1]


def get_mg_sequence(n_sp_lvl, k_seq, p_seq, n_timesteps_at_once, n_timesteps_at_once_min=1,
                    lower_lvl="k", coarsening_type="space_and_time", time_before_space=False,
                    use_p_multigrid_space=False, zip_from_back=True):
    """fe_time.cc:58-127. Level types as 1-char strings 't','k','h','p', coarse -> fine."""
    lower_lvl = {"tau": "t"}.get(lower_lvl, lower_lvl)
    n_k_lvl = len(k_seq) - 1
    n_t_lvl = int(math.log2(n_timesteps_at_once // n_timesteps_at_once_min)) \
        if n_timesteps_at_once // max(n_timesteps_at_once_min, 1) > 0 else 0
    upper_lvl = "t" if lower_lvl == "k" else "k"
    lower_s = "p" if lower_lvl == "k" else "h"
    upper_s = "h" if lower_lvl == "k" else "p"
    n_ll = n_k_lvl if lower_lvl == "k" else n_t_lvl
    n_ul = n_t_lvl if lower_lvl == "k" else n_k_lvl
    n_p_lvl = len(p_seq) - 1 if use_p_multigrid_space else 0
    n_ll_s = n_p_lvl if lower_lvl == "k" else n_sp_lvl - 1
    n_ul_s = n_sp_lvl - 1 if lower_lvl == "k" else n_p_lvl
    time_levels = [lower_lvl] * n_ll + [upper_lvl] * n_ul
    space_levels = [lower_s] * n_ll_s + [upper_s] * n_ul_s
    first = time_levels if time_before_space else space_levels
    second = space_levels if time_before_space else time_levels
    out = []
    if coarsening_type == "space_or_time":
        if zip_from_back:
            out = first[::-1] + second[::-1]
        else:
            out = first + second
    else:
        for i in range(max(len(time_levels), len(space_levels))):
            if i < len(first):
                out.append(first[len(first) - 1 - i] if zip_from_back else first[i])
            if i < len(second):
                out.append(second[len(second) - 1 - i] if zip_from_back else second[i])
        if zip_from_back:
            out = out[::-1]
    return out


def is_space_lvl(m):
    return m in ("h", "p")


def is_time_lvl(m):
    return m in ("t", "k")


def get_precondition_stmg_types(mg_type_level, coarsening_type, time_before_space, zip_from_back=True,
                                smoother=1):
    """fe_time.cc:129-150. 0 = Identity, 1 = Relaxation, 2 = Chebyshev."""
    ret = [smoother] * (len(mg_type_level) + 1)
    if coarsening_type == "space_or_time":
        return ret
    i = 0
    while i < len(mg_type_level) - 1:
        a, b = mg_type_level[i], mg_type_level[i + 1]
        hit = (is_space_lvl(a) and is_time_lvl(b)) if time_before_space else (is_time_lvl(a) and is_space_lvl(b))
        if hit:
            ret[i] = smoother
            ret[i + 1] = 0
            i += 1
        i += 1
    return ret


class BlockSlice:
    """block_indexing / BlockSlice, fe_time.h:901-1017 (variable-major default :1015)."""

    def __init__(self, n_timesteps_at_once=1, n_variables=1, n_timedofs=1, variable_major=True):
        self.nt, self.nv, self.nd, self.vm = n_timesteps_at_once, n_variables, n_timedofs, variable_major

    def n_blocks(self):
        return self.nt * self.nv * self.nd

    def index(self, timestep, variable, timedof):
        if self.vm:
            return timestep * (self.nv * self.nd) + variable * self.nd + timedof
        return timestep * (self.nv * self.nd) + timedof * self.nv + variable

    def decompose(self, index):
        t = index // (self.nv * self.nd)
        rem = index % (self.nv * self.nd)
        if self.vm:
            return t, rem // self.nd, rem % self.nd
        return t, rem % self.nv, rem // self.nv


def get_blk_indices(ttype, n_timesteps_at_once, n_variables, n_levels, mg_type_level, poly_time_sequence):
    """include/stmg.h:460-501: block structure per level (coarse -> fine)."""
    blk = [None] * n_levels
    p = len(poly_time_sequence) - 1
    nts = n_timesteps_at_once
    i = n_levels - 1
    for mgt in reversed(mg_type_level):
        nd = poly_time_sequence[p] + 1 if ttype == DG else poly_time_sequence[p]
        blk[i] = BlockSlice(nts, n_variables, nd)
        if mgt == "k":
            p -= 1
        elif mgt == "t":
            nts //= 2
        i -= 1
    nd = poly_time_sequence[p] + 1 if ttype == DG else poly_time_sequence[p]
    blk[0] = BlockSlice(nts, n_variables, nd)
    return blk
