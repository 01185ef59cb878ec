"""ORACLE (test infrastructure only — never imported by the product path).

1D quadrature rules and Lagrange bases on [0,1], restating what the reference
obtains from deal.II (un-vendored dependency, >= 9.6):
  QGauss<1>(n), QGaussLobatto<1>(n), QGaussRadau<1>(n, right)   -- reference
  include/fe_time.cc:152-169, include/fe_time.h:652-673,714, tests/tp_01.cc:77-78
  Polynomials::generate_complete_Lagrange_basis(points)         -- fe_time.cc:163-169
All rules are the textbook ones (Legendre roots / Lobatto / Radau), computed by
Newton iteration in double precision.
"""
import numpy as np
from numpy.polynomial import legendre as L


def _leg(n, x):
    """P_n(x), P_n'(x) on [-1,1]."""
    c = np.zeros(n + 1)
    c[n] = 1.0
    p = L.legval(x, c)
    dp = L.legval(x, L.legder(c))
    return p, dp


def gauss(n):
    """QGauss<1>(n): n points, exact to degree 2n-1, on [0,1]."""
    x, w = L.leggauss(n)
    # polish roots with Newton (leggauss is already accurate to ~1e-16)
    for _ in range(2):
        p, dp = _leg(n, x)
        x = x - p / dp
    _, dp = _leg(n, x)
    w = 2.0 / ((1.0 - x * x) * dp * dp)
    return 0.5 * (x + 1.0), 0.5 * w


def gauss_lobatto(n):
    """QGaussLobatto<1>(n): n points including both end points, on [0,1]."""
    assert n >= 2
    if n == 2:
        return np.array([0.0, 1.0]), np.array([0.5, 0.5])
    m = n - 1
    # interior nodes: roots of P'_m
    c = np.zeros(m + 1)
    c[m] = 1.0
    xi = np.sort(np.real(L.legroots(L.legder(c))))
    d1 = L.legder(c)
    d2 = L.legder(c, 2)
    for _ in range(3):
        xi = xi - L.legval(xi, d1) / L.legval(xi, d2)
    x = np.concatenate(([-1.0], xi, [1.0]))
    pm = L.legval(x, c)
    w = 2.0 / (m * (m + 1) * pm * pm)
    return 0.5 * (x + 1.0), 0.5 * w


def gauss_radau_right(n):
    """QGaussRadau<1>(n, EndPoint::right): n points including x=1, on [0,1]."""
    if n == 1:
        return np.array([1.0]), np.array([1.0])
    # left Radau on [-1,1]: x0=-1, others roots of (P_{n-1}+P_n)/(1+x); mirror for the right rule
    cn = np.zeros(n + 1)
    cn[n] = 1.0
    cm = np.zeros(n + 1)
    cm[n - 1] = 1.0
    q = cn + cm
    r = np.sort(np.real(L.legroots(q)))
    r = r[1:]  # drop the root at -1
    dq = L.legder(q)
    for _ in range(3):
        r = r - L.legval(r, q) / L.legval(r, dq)
    xl = np.concatenate(([-1.0], r))
    pm = L.legval(xl, cm)
    wl = (1.0 - xl) / (n * n * pm * pm)
    wl[0] = 2.0 / (n * n)
    x = -xl[::-1]
    w = wl[::-1]
    return 0.5 * (x + 1.0), 0.5 * w


def lagrange_eval(nodes, x):
    """values[i, q] = l_i(x_q) of the Lagrange basis on `nodes` (product form)."""
    nodes = np.asarray(nodes, dtype=float)
    x = np.atleast_1d(np.asarray(x, dtype=float))
    n = len(nodes)
    v = np.ones((n, len(x)))
    for i in range(n):
        for j in range(n):
            if j != i:
                v[i] *= (x - nodes[j]) / (nodes[i] - nodes[j])
    return v


def lagrange_deriv(nodes, x):
    """derivs[i, q] = l_i'(x_q)."""
    nodes = np.asarray(nodes, dtype=float)
    x = np.atleast_1d(np.asarray(x, dtype=float))
    n = len(nodes)
    d = np.zeros((n, len(x)))
    for i in range(n):
        for m in range(n):
            if m == i:
                continue
            t = np.ones(len(x)) / (nodes[i] - nodes[m])
            for j in range(n):
                if j != i and j != m:
                    t *= (x - nodes[j]) / (nodes[i] - nodes[j])
            d[i] += t
    return d
