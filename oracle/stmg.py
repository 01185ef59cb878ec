"""ORACLE (test infrastructure only — never imported by the product path).

numpy/scipy restatement of the space-time multigrid preconditioned FGMRES solve of the reference:
  MGTwoLevelTransferTime / MGTwoLevelBlockTransfer    include/stmg.h:38-247
  build_stmg_transfers / get_blk_indices              include/stmg.h:460-617
  PreconditionVanka (cell-patch additive Schwarz)     include/stmg.h:745-872,
      restrict_to_full_matrices_ (valence row scaling) include/compute_block_matrix.h:50-139
  GMG (smoothers, coarse solver, V-cycle, float<->double copies)  include/stmg.h:1047-1344
deal.II pieces restated from SURVEY.md App. A: MGTwoLevelTransfer (A.5), Multigrid::level_v_step and
MGSmootherPrecondition (A.6), PreconditionRelaxation with power-iteration estimate (A.7),
PreconditionChebyshev (A.8), SolverFGMRES (A.9).
Block vectors are arrays [nb, N]; level operators are assembled sparse matrices with constrained
rows/columns zeroed, which is exactly the matrix-free operator (tests/test_oracle_cpu_ref.py).
"""
import numpy as np
import scipy.sparse as sp

from . import fe_time as ft
from . import quadrature as Q
from . import spatial as S


# ----------------------------------------------------------------------------- level operator
class LevelOperator:
    """SystemMatrix on one level (include/operators.h:536-559) with assembled K, M."""

    def __init__(self, space, Alpha, Beta, dtype, coeff=None):
        self.space, self.dtype = space, dtype
        self.Alpha = np.asarray(Alpha, dtype=np.float64)
        self.Beta = np.asarray(Beta, dtype=np.float64)
        Kop = S.MatrixFreeOperator(space, 0.0, 1.0)
        Mop = S.MatrixFreeOperator(space, 1.0, 0.0)
        if coeff is not None:
            Kop.evaluate_coefficient(coeff)
        self.K_full = Kop.compute_system_matrix()          # with the constrained diagonal (for Vanka)
        self.M_full = Mop.compute_system_matrix()
        free = (~space.constrained).astype(np.float64)
        Dm = sp.diags(free)
        self.K = (Dm @ self.K_full @ Dm).tocsr().astype(dtype)   # operator: constrained rows/cols zero
        self.M = (Dm @ self.M_full @ Dm).tocsr().astype(dtype)
        self.nb = self.Alpha.shape[0]
        self.n = space.n_dofs

    def vmult(self, src):
        Ks = (self.K @ src.T).T
        Ms = (self.M @ src.T).T
        return (self.Alpha.astype(self.dtype) @ Ks + self.Beta.astype(self.dtype) @ Ms).astype(self.dtype)


# ----------------------------------------------------------------------------- space transfer
def _p1d_h(k, n_coarse):
    """1D nodal embedding coarse (n_coarse cells) -> fine (2 n_coarse cells), FE_Q(k) on GLL nodes."""
    gll = Q.gauss_lobatto(k + 1)[0]
    P = np.zeros((k * 2 * n_coarse + 1, k * n_coarse + 1))
    for c in range(n_coarse):
        for child in range(2):
            x = 0.5 * (gll + child)
            vals = Q.lagrange_eval(gll, x)           # [j, i] = l_j(x_i)
            for i in range(k + 1):
                row = k * (2 * c + child) + i
                P[row, k * c:k * c + k + 1] = vals[:, i]
    P[np.abs(P) < 1e-15] = 0.0
    return P


def _p1d_p(k_coarse, k_fine, n_cells):
    """1D degree embedding on the same mesh."""
    gc, gf = Q.gauss_lobatto(k_coarse + 1)[0], Q.gauss_lobatto(k_fine + 1)[0]
    vals = Q.lagrange_eval(gc, gf)                   # [j, i] = l_j^c(x_i^f)
    P = np.zeros((k_fine * n_cells + 1, k_coarse * n_cells + 1))
    for c in range(n_cells):
        for i in range(k_fine + 1):
            P[k_fine * c + i, k_coarse * c:k_coarse * c + k_coarse + 1] = vals[:, i]
    P[np.abs(P) < 1e-15] = 0.0
    return P


def space_prolongation(coarse, fine, dtype):
    """MGTwoLevelTransfer::reinit(fine, coarse, constraints) as a sparse matrix (App. A.5): nodal
    embedding; constrained fine rows and constrained coarse columns are zero."""
    d = fine.dim
    mats = []
    for a in range(d):
        if fine.mesh.n[a] == 2 * coarse.mesh.n[a]:
            assert fine.k == coarse.k
            mats.append(sp.csr_matrix(_p1d_h(fine.k, coarse.mesh.n[a])))
        else:
            assert fine.mesh.n[a] == coarse.mesh.n[a]
            mats.append(sp.csr_matrix(_p1d_p(coarse.k, fine.k, fine.mesh.n[a])))
    P = mats[0]
    for a in range(1, d):
        P = sp.kron(mats[a], P, format="csr")
    P = sp.diags((~fine.constrained).astype(float)) @ P @ sp.diags((~coarse.constrained).astype(float))
    return P.tocsr().astype(dtype)


class SpaceTransfer:
    """MGTwoLevelBlockTransfer (stmg.h:38-112): the same spatial transfer for every block."""

    def __init__(self, coarse, fine, dtype):
        self.P = space_prolongation(coarse, fine, dtype)
        self.R = self.P.T.tocsr()

    def prolongate_and_add(self, dst, src):
        dst += (self.P @ src.T).T

    def restrict_and_add(self, dst, src):
        dst += (self.R @ src.T).T


class TimeTransfer:
    """MGTwoLevelTransferTime (stmg.h:114-247)."""

    def __init__(self, ttype, blk_hi, blk_lo, restrict_is_transpose_prolongate, mg_type, dtype):
        k_mg = mg_type == "k"
        r = blk_hi.nd - 1 if ttype == ft.DG else blk_hi.nd
        r_lo = blk_lo.nd - 1 if ttype == ft.DG else blk_lo.nd
        nts = blk_hi.nt
        if k_mg:
            self.P = ft.get_time_projection_matrix(ttype, r_lo, r, nts)
            down = ft.get_time_projection_matrix(ttype, r, r_lo, nts)
        else:
            self.P = ft.get_time_prolongation_matrix(ttype, r, nts)
            down = ft.get_time_restriction_matrix(ttype, r, nts)
        self.R = self.P.T.copy() if restrict_is_transpose_prolongate else down
        self.P = self.P.astype(dtype)
        self.R = self.R.astype(dtype)
        self.down = down.astype(dtype)

    def prolongate_and_add(self, dst, src):
        dst += self.P @ src

    def restrict_and_add(self, dst, src):
        dst += self.R @ src


# ----------------------------------------------------------------------------- Vanka
class PointJacobi:
    """Point-Jacobi inner preconditioner (NOT a reference configuration; BASELINE.json's north_star names it as the cheap
    smoother): the inverse of SystemMatrix::get_matrix_diagonal (operators.h:613-625), diag_b = Alpha(b,b) diag K +
    Beta(b,b) diag M with the constrained rows of diag K, diag M zero, and the reference's rule for vanishing entries
    (operators.h:1105-1109: |d| > sqrt(eps) ? 1/d : 1).  Same interface as PreconditionVanka (GMG(vanka=[...]))."""

    def __init__(self, levelop, dtype):
        free = ~levelop.space.constrained
        dK = np.where(free, levelop.K_full.diagonal(), 0.0)
        dM = np.where(free, levelop.M_full.diagonal(), 0.0)
        nb = levelop.nb
        d = np.stack([levelop.Alpha[b, b] * dK + levelop.Beta[b, b] * dM for b in range(nb)])
        tol = np.sqrt(np.finfo(dtype).eps)
        self.diag = d
        self.inv = np.where(np.abs(d) > tol, 1.0 / np.where(d == 0, 1.0, d), 1.0).astype(dtype)
        self.dtype = dtype

    def vmult(self, src):
        return (self.inv * src).astype(self.dtype)


class PreconditionVanka:
    """stmg.h:745-872: per cell, B = Beta (x) M_c + Alpha (x) K_c from the ASSEMBLED matrices restricted to
    the cell's DoFs (rows scaled by the valence, compute_block_matrix.h:135-136), inverted."""

    def __init__(self, levelop, dtype):
        s = levelop.space
        self.space, self.dtype = s, dtype
        cd = s.cell_dofs
        nc = cd.shape[1]
        valence = np.zeros(s.n_dofs)
        np.add.at(valence, cd.reshape(-1), 1.0)
        Kf, Mf = levelop.K_full.tocsr(), levelop.M_full.tocsr()
        A, Bt = levelop.Alpha, levelop.Beta
        nb = A.shape[0]
        self.nb, self.nc = nb, nc
        blocks = np.zeros((cd.shape[0], nb * nc, nb * nc), dtype=dtype)
        # identical patches (Cartesian interior cells) are inverted once
        cache = {}
        for c in range(cd.shape[0]):
            idx = cd[c]
            Kc = Kf[idx][:, idx].toarray() * valence[idx][:, None]
            Mc = Mf[idx][:, idx].toarray() * valence[idx][:, None]
            key = (Kc.round(14).tobytes(), Mc.round(14).tobytes()) if s.mesh.cartesian else None
            if key is not None and key in cache:
                blocks[c] = cache[key]
                continue
            Bm = np.kron(Bt, Mc) + np.kron(A, Kc)
            inv = np.linalg.inv(Bm).astype(dtype)
            blocks[c] = inv
            if key is not None:
                cache[key] = inv
        self.blocks = blocks

    def vmult(self, src):
        s = self.space
        cd = s.cell_dofs
        loc = src[:, cd]                                   # [nb, C, nc]
        loc = np.transpose(loc, (1, 0, 2)).reshape(cd.shape[0], -1)
        out = np.einsum("cij,cj->ci", self.blocks, loc).reshape(cd.shape[0], self.nb, self.nc)
        dst = np.zeros_like(src)
        for b in range(self.nb):
            np.add.at(dst[b], cd.reshape(-1), out[:, b, :].reshape(-1))
        return dst


# ----------------------------------------------------------------------------- smoothers
def _initial_guess(nb, n, dtype):
    """deal.II set_initial_guess for distributed vectors: (i mod 11) minus the mean, per block (A.7)."""
    v = (np.arange(n) % 11).astype(np.float64)
    v -= v.mean()
    return np.tile(v.astype(dtype), (nb, 1))


def power_iteration(A, P, nb, n, n_iterations, dtype):
    """largest eigenvalue of P^-1 A by power iteration (A.7)."""
    v = _initial_guess(nb, n, dtype)
    nrm = np.linalg.norm(v.astype(np.float64))
    if nrm == 0.0:
        return 1.0
    v = (v / nrm).astype(dtype)
    ev = 0.0
    for _ in range(n_iterations):
        w = P.vmult(A.vmult(v))
        ev = float(np.vdot(v.astype(np.float64), w.astype(np.float64)))
        nw = np.linalg.norm(w.astype(np.float64))
        if nw == 0.0 or not np.isfinite(nw):
            return 1.0                                     # guard (SURVEY App. C.2 hazard)
        v = (w / nw).astype(dtype)
    return ev


class Relaxation:
    """PreconditionRelaxation<A, Vanka> (A.7): x = w P^-1 b, then n_iterations-1 Richardson steps."""

    def __init__(self, A, P, n_iterations=1, relaxation=0.0, smoothing_range=1.0, eig_n_iterations=20):
        self.A, self.P, self.n_it = A, P, n_iterations
        if relaxation == 0.0:
            lam = power_iteration(A, P, A.nb, A.n, eig_n_iterations, A.dtype)
            lmax = 1.2 * lam
            lmin = lam / smoothing_range if smoothing_range > 1.0 else lam
            alpha = lmax / smoothing_range if smoothing_range > 1.0 else min(0.9 * lmax, lmin)
            relaxation = 2.0 / (alpha + lmax)
            self.lambda_estimate = lam
        self.omega = relaxation

    def vmult(self, b):
        w = self.A.dtype(self.omega)
        x = w * self.P.vmult(b)
        for _ in range(self.n_it - 1):
            x = x + w * self.P.vmult(b - self.A.vmult(x))
        return x


class Chebyshev:
    """PreconditionChebyshev<A, V, Vanka> of degree `degree` on [lmax/range', lmax] (A.8),
    eigenvalue by power iteration, 1.2 safety factor."""

    def __init__(self, A, P, degree=1, smoothing_range=1.0, eig_n_iterations=20):
        self.A, self.P, self.degree = A, P, degree
        lam = power_iteration(A, P, A.nb, A.n, eig_n_iterations, A.dtype)
        lmax = 1.2 * lam
        rng = smoothing_range if smoothing_range > 1.0 else 1.0
        lmin = lam / rng if smoothing_range > 1.0 else min(0.9 * lmax, lam)
        self.theta = 0.5 * (lmax + lmin)
        self.delta = 0.5 * (lmax - lmin)
        self.lambda_estimate = lam

    def vmult(self, b):
        # first-kind Chebyshev iteration started from x0 = 0
        A, P = self.A, self.P
        th, de = self.theta, self.delta
        dt = A.dtype
        d = dt(1.0 / th) * P.vmult(b)
        x = d.copy()
        if self.degree < 2:
            return x
        sigma = th / de if de != 0 else np.inf
        rho_old = 1.0 / sigma if np.isfinite(sigma) else 0.0
        for _ in range(self.degree - 1):
            rho = 1.0 / (2.0 * sigma - rho_old) if np.isfinite(sigma) else 0.0
            r = b - A.vmult(x)
            d = dt(rho * rho_old) * d + dt(2.0 * rho / de if de != 0 else 1.0 / th) * P.vmult(r)
            x = x + d
            rho_old = rho
        return x


class Identity:
    def vmult(self, b):
        return b.copy()


# ----------------------------------------------------------------------------- multigrid
class GMG:
    """GMG (stmg.h:1047-1344): V-cycle with MGSmootherPrecondition(1, variable, false, false) and
    MGCoarseGridApplySmoother, run in `dtype` (float for test<double,float>, tp_01.cc:801-804)."""

    def __init__(self, ttype, level_ops, spaces, mg_type_level, poly_time_sequence, n_timesteps_at_once,
                 precondition_types, dtype, smoothing_steps=1, relaxation=0.0, smoothing_range=1.0,
                 eig_n_iterations=20, variable=True, restrict_is_transpose_prolongate=True,
                 vanka=None, coarse_gmres=None):
        self.ops, self.dtype = level_ops, dtype
        nl = len(level_ops)
        self.nl = nl
        self.blk = ft.get_blk_indices(ttype, n_timesteps_at_once, 1, nl, mg_type_level, poly_time_sequence)
        self.transfers = [None] * nl
        for l in range(1, nl):
            t = mg_type_level[l - 1]
            if t in ("h", "p"):
                self.transfers[l] = SpaceTransfer(spaces[l - 1], spaces[l], dtype)
            else:
                self.transfers[l] = TimeTransfer(ttype, self.blk[l], self.blk[l - 1],
                                                 restrict_is_transpose_prolongate, t, dtype)
        self.vanka = vanka if vanka is not None else [None] * nl
        self.smoothers = []
        for l in range(nl):
            if precondition_types[l] == 0:
                self.smoothers.append(Identity())
                continue
            if self.vanka[l] is None:
                self.vanka[l] = PreconditionVanka(level_ops[l], dtype)
            if precondition_types[l] == 1:
                self.smoothers.append(Relaxation(level_ops[l], self.vanka[l], smoothing_steps, relaxation,
                                                 smoothing_range, eig_n_iterations))
            else:
                self.smoothers.append(Chebyshev(level_ops[l], self.vanka[l], smoothing_steps, smoothing_range,
                                                eig_n_iterations))
        self.steps = [(1 << (nl - 1 - l)) if variable else 1 for l in range(nl)]
        # coarseGridSmootherType != "Smoother" (stmg.h:1240-1302): MGCoarseGridIterativeSolver with SolverGMRES (left
        # preconditioning) + IterationNumberControl(coarse_grid_maxiter, coarse_grid_abstol) and the coarse smoother as
        # preconditioner; coarse_gmres = (maxiter, abstol)
        self.coarse_gmres = coarse_gmres

    def _apply(self, l, rhs):
        """MGSmootherPrecondition::apply (zero initial guess)."""
        A, Pm = self.ops[l], self.smoothers[l]
        u = Pm.vmult(rhs)
        for _ in range(self.steps[l] - 1):
            u = u + Pm.vmult(rhs - A.vmult(u))
        return u

    def _smooth(self, l, u, rhs):
        A, Pm = self.ops[l], self.smoothers[l]
        for _ in range(self.steps[l]):
            u = u + Pm.vmult(rhs - A.vmult(u))
        return u

    def _v(self, l, defect):
        """Multigrid::level_v_step (A.6)."""
        if l == 0:
            if self.coarse_gmres is not None:
                return coarse_gmres(self.ops[0].vmult, self.smoothers[0].vmult, defect, self.coarse_gmres[0], self.coarse_gmres[1],
                                    self.dtype)
            return self._apply(0, defect)
        A = self.ops[l]
        sol = self._apply(l, defect)
        t = defect - A.vmult(sol)
        dc = np.zeros((self.ops[l - 1].nb, self.ops[l - 1].n), dtype=self.dtype)
        self.transfers[l].restrict_and_add(dc, t)
        sc = self._v(l - 1, dc)
        t = np.zeros_like(sol)
        self.transfers[l].prolongate_and_add(t, sc)
        sol = sol + t
        return self._smooth(l, sol, defect)

    def vmult(self, src):
        """GMG::vmult (stmg.h:1331-1344): double -> level precision -> double."""
        out = self._v(self.nl - 1, src.astype(self.dtype))
        return out.astype(np.float64)


# ----------------------------------------------------------------------------- coarse-grid GMRES
def coarse_gmres(A, P, b, maxiter, abstol, dtype):
    """SolverGMRES (deal.II default: LEFT preconditioning) with IterationNumberControl(maxiter, abstol) and
    max_n_tmp_vectors = maxiter, zero start vector: min || P (b - A x) || over the Krylov space of P A
    (stmg.h:1240-1250, 1267-1275).  Vectors in `dtype`, inner products accumulated in double."""
    x = np.zeros_like(b)
    r = P(b)
    beta = float(np.sqrt(np.vdot(r.astype(np.float64), r.astype(np.float64))))
    if beta == 0.0:
        return x
    V = [(r / dtype(beta)).astype(dtype)]
    H = np.zeros((maxiter + 1, maxiter))
    g = np.zeros(maxiter + 1)
    g[0] = beta
    cs, sn = np.zeros(maxiter), np.zeros(maxiter)
    jd = 0
    for j in range(maxiter):
        w = P(A(V[j])).astype(dtype)
        for _ in range(2):
            h = np.array([np.vdot(v.astype(np.float64), w.astype(np.float64)) for v in V])
            for hv, v in zip(h, V):
                w = (w - dtype(hv) * v).astype(dtype)
            H[:j + 1, j] += h
        hn = float(np.sqrt(np.vdot(w.astype(np.float64), w.astype(np.float64))))
        H[j + 1, j] = hn
        for i in range(j):
            t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
            H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
            H[i, j] = t
        den = np.hypot(H[j, j], H[j + 1, j])
        if den == 0.0:
            break
        cs[j], sn[j] = H[j, j] / den, H[j + 1, j] / den
        H[j, j] = den
        H[j + 1, j] = 0.0
        g[j + 1] = -sn[j] * g[j]
        g[j] = cs[j] * g[j]
        jd = j + 1
        if abs(g[j + 1]) < abstol or hn == 0.0:
            break
        V.append((w / dtype(hn)).astype(dtype))
    if jd == 0:
        return x
    y = np.linalg.solve(np.triu(H[:jd, :jd]), g[:jd])
    for yj, v in zip(y, V):
        x = (x + dtype(yj) * v).astype(dtype)
    return x


# ----------------------------------------------------------------------------- FGMRES
def fgmres(A, x, b, M, max_basis=100, max_iter=200, abstol=1e-12, reduce=1e-12):
    """SolverFGMRES with ReductionControl(200, 1e-12, 1e-12) (time_integrators.h:56-59, A.9).
    A, M: callables on [nb, N] arrays.  Returns (x, iterations, residual history)."""
    it = 0
    hist = []
    r0 = None
    while True:
        r = b - A(x)
        beta = np.linalg.norm(r)
        if r0 is None:
            r0 = beta
            hist.append(beta)
            tol = max(abstol, reduce * r0)
            if beta <= tol:
                return x, 0, hist
        V = [r / beta]
        Z = []
        H = np.zeros((max_basis + 1, max_basis))
        g = np.zeros(max_basis + 1)
        g[0] = beta
        cs, sn = np.zeros(max_basis), np.zeros(max_basis)
        j_done = 0
        converged = False
        for j in range(max_basis):
            z = M(V[j])
            w = A(z)
            Z.append(z)
            # classical Gram-Schmidt with one re-orthogonalisation
            for _ in range(2):
                h = np.array([np.vdot(v, w) for v in V])
                for hv, v in zip(h, V):
                    w = w - hv * v
                H[:j + 1, j] += h
            H[j + 1, j] = np.linalg.norm(w)
            if H[j + 1, j] != 0:
                V.append(w / H[j + 1, j])
            for i in range(j):
                t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                H[i, j] = t
            den = np.hypot(H[j, j], H[j + 1, j])
            cs[j], sn[j] = H[j, j] / den, H[j + 1, j] / den
            H[j, j] = den
            H[j + 1, j] = 0.0
            g[j + 1] = -sn[j] * g[j]
            g[j] = cs[j] * g[j]
            it += 1
            j_done = j + 1
            res = abs(g[j + 1])
            hist.append(res)
            if res <= tol or it >= max_iter:
                converged = res <= tol
                break
        y = np.linalg.solve(np.triu(H[:j_done, :j_done]), g[:j_done])
        for yj, z in zip(y, Z):
            x = x + yj * z
        if converged or it >= max_iter:
            return x, it, hist
