// Host-side 1D quadrature rules and Lagrange bases on [0,1] (long double Newton iterations).
// Product code: feeds the kernels' shape matrices and the fe_time algebra.  The deal.II objects
// the reference obtains these from are QGauss / QGaussLobatto / QGaussRadau and
// Polynomials::generate_complete_Lagrange_basis (reference include/fe_time.cc:152-169,
// tests/tp_01.cc:77-78).
#pragma once
#include <cmath>
#include <vector>

namespace stfem
{
  using ld = long double;

  // Legendre P_n(x) and P_n'(x) on [-1,1]
  inline void legendre(int n, ld x, ld &p, ld &dp)
  {
    ld p0 = 1, p1 = x;
    if (n == 0) { p = 1; dp = 0; return; }
    for (int k = 2; k <= n; ++k)
      {
        ld pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
        p0 = p1; p1 = pk;
      }
    p = p1;
    dp = n * (x * p1 - p0) / (x * x - 1);
  }

  struct Rule { std::vector<double> x, w; };

  inline Rule gauss(int n)
  {
    Rule r; r.x.resize(n); r.w.resize(n);
    const ld pi = acosl(-1.0L);
    for (int i = 0; i < n; ++i)
      {
        ld x = -cosl(pi * (i + 0.75L) / (n + 0.5L));
        for (int it = 0; it < 100; ++it)
          {
            ld p, dp; legendre(n, x, p, dp);
            ld dx = p / dp; x -= dx;
            if (fabsl(dx) < 1e-19L) break;
          }
        ld p, dp; legendre(n, x, p, dp);
        r.x[i] = (double)(0.5L * (x + 1));
        r.w[i] = (double)(1.0L / ((1 - x * x) * dp * dp));
      }
    return r;
  }

  inline Rule gauss_lobatto(int n)
  {
    Rule r; r.x.resize(n); r.w.resize(n);
    const int m = n - 1;
    const ld pi = acosl(-1.0L);
    std::vector<ld> xs(n);
    xs[0] = -1; xs[m] = 1;
    for (int i = 1; i < m; ++i)
      {
        // roots of P'_m: Newton on q(x) = P'_m, q' from the Legendre ODE
        ld x = -cosl(pi * i / m);
        for (int it = 0; it < 100; ++it)
          {
            ld p, dp; legendre(m, x, p, dp);
            ld d2p = (2 * x * dp - m * (m + 1) * p) / (1 - x * x);
            ld dx = dp / d2p; x -= dx;
            if (fabsl(dx) < 1e-19L) break;
          }
        xs[i] = x;
      }
    for (int i = 0; i < n; ++i)
      {
        ld p, dp;
        if (i == 0) p = (m % 2 == 0) ? 1 : -1; else if (i == m) p = 1; else legendre(m, xs[i], p, dp);
        r.x[i] = (double)(0.5L * (xs[i] + 1));
        r.w[i] = (double)(1.0L / (m * (m + 1) * p * p));
      }
    return r;
  }

  // n points including the right end point 1
  inline Rule gauss_radau_right(int n)
  {
    Rule r; r.x.resize(n); r.w.resize(n);
    if (n == 1) { r.x[0] = 1; r.w[0] = 1; return r; }
    // left rule on [-1,1]: x0 = -1, others roots of f = (P_{n-1}+P_n)/(1+x)
    const ld pi = acosl(-1.0L);
    std::vector<ld> xl(n), wl(n);
    xl[0] = -1; wl[0] = 2.0L / (n * (ld)n);
    for (int i = 1; i < n; ++i)
      {
        ld x = -cosl(pi * (2 * i + 0.5L) / (2 * n - 1 + 0.5L));
        for (int it = 0; it < 200; ++it)
          {
            ld p0, d0, p1, d1; legendre(n - 1, x, p0, d0); legendre(n, x, p1, d1);
            ld f = (p0 + p1) / (1 + x);
            ld df = ((d0 + d1) - f) / (1 + x);
            // deflate the roots already found
            ld s = 0;
            for (int k = 1; k < i; ++k) s += 1 / (x - xl[k]);
            ld dx = f / (df - f * s); x -= dx;
            if (fabsl(dx) < 1e-19L) break;
          }
        xl[i] = x;
        ld p0, d0; legendre(n - 1, x, p0, d0);
        wl[i] = (1 - x) / (n * (ld)n * p0 * p0);
      }
    // sort ascending (deflation order is not guaranteed monotone)
    for (int i = 1; i < n; ++i)
      for (int j = i + 1; j < n; ++j)
        if (xl[j] < xl[i]) { std::swap(xl[i], xl[j]); std::swap(wl[i], wl[j]); }
    for (int i = 0; i < n; ++i)
      {
        r.x[i] = (double)(0.5L * (1 - xl[n - 1 - i]));
        r.w[i] = (double)(0.5L * wl[n - 1 - i]);
      }
    return r;
  }

  // l_i(x) of the Lagrange basis on `nodes`
  inline double lagrange_value(const std::vector<double> &nodes, int i, double x)
  {
    ld v = 1;
    for (size_t j = 0; j < nodes.size(); ++j)
      if ((int)j != i) v *= ((ld)x - nodes[j]) / ((ld)nodes[i] - nodes[j]);
    return (double)v;
  }

  inline double lagrange_deriv(const std::vector<double> &nodes, int i, double x)
  {
    ld d = 0;
    for (size_t m = 0; m < nodes.size(); ++m)
      {
        if ((int)m == i) continue;
        ld t = 1 / ((ld)nodes[i] - nodes[m]);
        for (size_t j = 0; j < nodes.size(); ++j)
          if ((int)j != i && j != m) t *= ((ld)x - nodes[j]) / ((ld)nodes[i] - nodes[j]);
        d += t;
      }
    return (double)d;
  }

  // Shape data of FE_Q(k) with QGauss(k+1): S[q][i] = phi_i(x_q), D[q][i] = phi_i'(x_q),
  // Dc[q][p] = derivative of the Lagrange basis ON the Gauss points (collocation), D = Dc * S.
  struct ShapeHost
  {
    int n1;
    std::vector<double> gll, xq, wq, S, D, Dc;
    explicit ShapeHost(int degree) : n1(degree + 1)
    {
      Rule g = gauss(n1), l = gauss_lobatto(n1);
      gll = l.x; xq = g.x; wq = g.w;
      S.resize(n1 * n1); D.resize(n1 * n1); Dc.resize(n1 * n1);
      for (int q = 0; q < n1; ++q)
        for (int i = 0; i < n1; ++i)
          {
            S[q * n1 + i] = lagrange_value(gll, i, xq[q]);
            D[q * n1 + i] = lagrange_deriv(gll, i, xq[q]);
            Dc[q * n1 + i] = lagrange_deriv(xq, i, xq[q]);
          }
    }
  };
} // namespace stfem
