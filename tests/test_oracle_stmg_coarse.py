"""Oracle of the coarse-grid GMRES option (oracle/stmg.py: coarse_gmres; reference include/stmg.h:1240-1302): the
left-preconditioned GMRES with IterationNumberControl(maxiter, abstol) minimises || P (b - A x) || over the Krylov space of
P A - checked against a dense least-squares solution, and inside the V-cycle against the exact coarse solve."""
import numpy as np

from golden_util import load
from oracle import stmg, tp_01

G = load("tp_01")


def test_coarse_gmres_minimises_the_preconditioned_residual():
    rng = np.random.RandomState(3)
    n = 12
    Am = np.eye(n) * 4 + rng.uniform(-1, 1, (n, n))
    Pm = np.linalg.inv(np.diag(np.diag(Am)))
    b = rng.uniform(-1, 1, (1, n))
    A = lambda v: (Am @ v.reshape(-1)).reshape(1, n)
    P = lambda v: (Pm @ v.reshape(-1)).reshape(1, n)
    for m in (1, 3, 6, 12):
        x = stmg.coarse_gmres(A, P, b, m, 1e-20, np.float64)
        # Krylov basis of P A applied to P b
        K = [Pm @ b.reshape(-1)]
        for _ in range(m - 1):
            K.append(Pm @ Am @ K[-1])
        K = np.stack(K, axis=1)
        y, *_ = np.linalg.lstsq(Pm @ Am @ K, Pm @ b.reshape(-1), rcond=None)
        assert np.abs(x.reshape(-1) - K @ y).max() < 1e-9 * np.abs(K @ y).max()
    assert np.abs(Am @ stmg.coarse_gmres(A, P, b, n, 1e-20, np.float64).reshape(-1) - b.reshape(-1)).max() < 1e-10


def test_vcycle_with_coarse_gmres_reduces_the_error_like_the_exact_coarse_solve():
    p = tp_01.parse_parameters(G["params"]["tf03"], 2)
    lv = tp_01.build_levels(p, 2, 2, p["feDegree"], p["endTime"] * 2.0 ** -3, np.float64)
    kw = dict(smoothing_steps=p["smoothingSteps"], relaxation=p["relaxation"], smoothing_range=p["smoothingRange"],
              eig_n_iterations=p["smoothingEigCgNIterations"], variable=p["variable"])
    mg_s = stmg.GMG(p["timeType"], lv["ops"], lv["spaces"], lv["mg_type_level"], lv["poly_time"], p["nTimestepsAtOnce"], lv["ptypes"],
                    np.float64, **kw)
    mg_g = stmg.GMG(p["timeType"], lv["ops"], lv["spaces"], lv["mg_type_level"], lv["poly_time"], p["nTimestepsAtOnce"], lv["ptypes"],
                    np.float64, vanka=mg_s.vanka, coarse_gmres=(10, 1e-20), **kw)
    top = lv["ops"][-1]
    x = np.random.RandomState(5).uniform(-1, 1, (top.nb, top.n))
    x[:, lv["spaces"][-1].constrained] = 0
    b = top.vmult(x)
    e_s = np.linalg.norm(x - mg_s.vmult(b))
    e_g = np.linalg.norm(x - mg_g.vmult(b))
    assert e_g < 0.9 * np.linalg.norm(x)                 # a contraction
    assert e_g < 1.5 * e_s                               # and not worse than the smoother as coarse solver
