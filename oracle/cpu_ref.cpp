// ORACLE / CPU BASELINE (test infrastructure only — never linked into the product library).
//
// Plain C++/OpenMP restatement of the reference's CPU algorithm for the operator path, *as written*:
//   SystemMatrix::vmult            reference include/operators.h:536-559
//        for each block i:  tmp = K src_i ; dst_j += Alpha(j,i) tmp ;  tmp = M src_i ; dst_j += Beta(j,i) tmp
//   MatrixFreeOperator::vmult      operators.h:1013-1018  (cell_loop, zero_dst = true)
//   do_cell_integral_range/local   operators.h:1112-1173  (gather, evaluate, q-loop, integrate, scatter)
// i.e. 2*nb separate sum-factorised cell loops plus nb + nnz(Beta) vector updates (UNFUSED), which is
// what the GPU kernel is compared with in bench.py's cpu_baseline.  deal.II's FEEvaluation is restated
// as straightforward sum factorisation (S = GLL->Gauss values, D = GLL->Gauss derivatives).
// Threading: OpenMP over cells with an 2^dim colouring (deal.II uses MPI ranks / partition_partition).
// Parity of this file itself is pinned against oracle/spatial.py in tests/test_oracle_cpu_ref.py.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace
{
  struct Setup
  {
    int           dim, n[3], np[3], n1;
    long long     n_cells, N;
    const double *S, *D, *w;        // n1*n1, n1*n1, n1
    const double *metric;           // per cell per q: nsym + 1 entries (G sym, JxW), or null (Cartesian)
    const double *coeff_q;          // per cell per q Laplace coefficient or null
    double        h[3];
    unsigned      dirichlet;
  };

  inline bool constrained(const Setup &s, int ix, int iy, int iz)
  {
    const unsigned d = s.dirichlet;
    if ((d & 1u) && ix == 0) return true;
    if ((d & 2u) && ix == s.np[0] - 1) return true;
    if ((d & 4u) && iy == 0) return true;
    if ((d & 8u) && iy == s.np[1] - 1) return true;
    if (s.dim == 3)
      {
        if ((d & 16u) && iz == 0) return true;
        if ((d & 32u) && iz == s.np[2] - 1) return true;
      }
    return false;
  }

  // y[q] = sum_i M[q*n+i] x[i] along direction `dir` of an n^DIM tensor (TR: transpose of M)
  template <int DIM, int N, bool TR, bool ADD>
  inline void sweep(const double *M, const double *in, double *out, int dir)
  {
    constexpr int NT     = (DIM == 3) ? N * N * N : N * N;
    const int     stride = dir == 0 ? 1 : (dir == 1 ? N : N * N);
    for (int outer = 0; outer < NT / (N * stride); ++outer)
      for (int inner = 0; inner < stride; ++inner)
        {
          const int base = outer * N * stride + inner;
          for (int q = 0; q < N; ++q)
            {
              double s = 0;
              for (int i = 0; i < N; ++i) s += (TR ? M[i * N + q] : M[q * N + i]) * in[base + i * stride];
              if (ADD)
                out[base + q * stride] += s;
              else
                out[base + q * stride] = s;
            }
        }
  }

  // one scalar operator application  dst = (mass ? M : K) src,   cell loop with gather/scatter
  template <int DIM, int N>
  void cell_loop(const Setup &s, bool mass, const double *src, double *dst)
  {
    constexpr int NT   = (DIM == 3) ? N * N * N : N * N;
    constexpr int NSYM = DIM * (DIM + 1) / 2;
    constexpr int K    = N - 1;
    std::memset(dst, 0, sizeof(double) * s.N);
    const int ncol = 1 << DIM;
    for (int colour = 0; colour < ncol; ++colour)
      {
#pragma omp parallel for schedule(static)
        for (long long cell = 0; cell < s.n_cells; ++cell)
          {
            const int cx = (int)(cell % s.n[0]), cy = (int)((cell / s.n[0]) % s.n[1]);
            const int cz = DIM == 3 ? (int)(cell / ((long long)s.n[0] * s.n[1])) : 0;
            if (((cx & 1) | ((cy & 1) << 1) | ((cz & 1) << 2)) != colour) continue;
            double u[NT], t0[NT], t1[NT], g[DIM][NT], r[NT];
            // gather (read_dof_values)
            for (int l = 0; l < (DIM == 3 ? N : 1); ++l)
              for (int j = 0; j < N; ++j)
                for (int i = 0; i < N; ++i)
                  {
                    const int ix = cx * K + i, iy = cy * K + j, iz = DIM == 3 ? cz * K + l : 0;
                    const long long gi = ix + (long long)s.np[0] * (iy + (long long)s.np[1] * iz);
                    u[(l * N + j) * N + i] = constrained(s, ix, iy, iz) ? 0.0 : src[gi];
                  }
            const double *mt = s.metric ? s.metric + (size_t)cell * NT * (NSYM + 1) : nullptr;
            const double *cq = s.coeff_q ? s.coeff_q + (size_t)cell * NT : nullptr;
            double        vol = s.h[0] * s.h[1] * (DIM == 3 ? s.h[2] : 1.0);
            if (mass)
              {
                // evaluate(values)
                sweep<DIM, N, false, false>(s.S, u, t0, 0);
                sweep<DIM, N, false, false>(s.S, t0, t1, 1);
                double *uq = t1;
                if (DIM == 3) { sweep<DIM, N, false, false>(s.S, t1, t0, 2); uq = t0; }
                for (int q = 0; q < NT; ++q)
                  {
                    const int qx = q % N, qy = (q / N) % N, qz = q / (N * N);
                    const double jxw = mt ? mt[q * (NSYM + 1) + NSYM] : vol * s.w[qx] * s.w[qy] * (DIM == 3 ? s.w[qz] : 1.0);
                    uq[q] *= jxw; // submit_value
                  }
                // integrate(values)
                double *a = uq, *b = (uq == t0) ? t1 : t0;
                if (DIM == 3) { sweep<DIM, N, true, false>(s.S, a, b, 2); std::swap(a, b); }
                sweep<DIM, N, true, false>(s.S, a, b, 1);
                sweep<DIM, N, true, false>(s.S, b, r, 0);
              }
            else
              {
                // evaluate(gradients):  d/dx = D (x) S (x) S  etc.
                for (int d = 0; d < DIM; ++d)
                  {
                    sweep<DIM, N, false, false>(d == 0 ? s.D : s.S, u, t0, 0);
                    sweep<DIM, N, false, false>(d == 1 ? s.D : s.S, t0, t1, 1);
                    if (DIM == 3)
                      sweep<DIM, N, false, false>(d == 2 ? s.D : s.S, t1, g[d], 2);
                    else
                      std::memcpy(g[d], t1, sizeof(double) * NT);
                  }
                // submit_gradient(coef * grad u): real-space gradient folded with JxW
                for (int q = 0; q < NT; ++q)
                  {
                    const int    qx = q % N, qy = (q / N) % N, qz = q / (N * N);
                    const double c  = cq ? cq[q] : 1.0;
                    double       gi[DIM], to[DIM];
                    for (int d = 0; d < DIM; ++d) gi[d] = g[d][q];
                    if (mt)
                      {
                        const double *m = mt + q * (NSYM + 1);
                        if (DIM == 2)
                          {
                            to[0] = m[0] * gi[0] + m[1] * gi[1];
                            to[1] = m[1] * gi[0] + m[2] * gi[1];
                          }
                        else
                          {
                            to[0] = m[0] * gi[0] + m[1] * gi[1] + m[2] * gi[2];
                            to[1] = m[1] * gi[0] + m[3] * gi[1] + m[4] * gi[2];
                            to[2] = m[2] * gi[0] + m[4] * gi[1] + m[5] * gi[2];
                          }
                      }
                    else
                      {
                        const double wq = vol * s.w[qx] * s.w[qy] * (DIM == 3 ? s.w[qz] : 1.0);
                        for (int d = 0; d < DIM; ++d) to[d] = wq / (s.h[d] * s.h[d]) * gi[d];
                      }
                    for (int d = 0; d < DIM; ++d) g[d][q] = c * to[d];
                  }
                // integrate(gradients)
                std::memset(r, 0, sizeof(r));
                for (int d = 0; d < DIM; ++d)
                  {
                    const double *a = g[d];
                    if (DIM == 3)
                      {
                        sweep<DIM, N, true, false>(d == 2 ? s.D : s.S, a, t0, 2);
                        a = t0;
                      }
                    sweep<DIM, N, true, false>(d == 1 ? s.D : s.S, a, t1, 1);
                    sweep<DIM, N, true, true>(d == 0 ? s.D : s.S, t1, r, 0);
                  }
              }
            // scatter (distribute_local_to_global)
            for (int l = 0; l < (DIM == 3 ? N : 1); ++l)
              for (int j = 0; j < N; ++j)
                for (int i = 0; i < N; ++i)
                  {
                    const int ix = cx * K + i, iy = cy * K + j, iz = DIM == 3 ? cz * K + l : 0;
                    if (constrained(s, ix, iy, iz)) continue;
                    const long long gi = ix + (long long)s.np[0] * (iy + (long long)s.np[1] * iz);
                    dst[gi] += r[(l * N + j) * N + i];
                  }
          }
      }
  }

  template <int DIM>
  void cell_loop_deg(const Setup &s, bool mass, const double *src, double *dst)
  {
    switch (s.n1)
      {
        case 2: cell_loop<DIM, 2>(s, mass, src, dst); break;
        case 3: cell_loop<DIM, 3>(s, mass, src, dst); break;
        case 4: cell_loop<DIM, 4>(s, mass, src, dst); break;
        case 5: cell_loop<DIM, 5>(s, mass, src, dst); break;
        case 6: cell_loop<DIM, 6>(s, mass, src, dst); break;
        case 7: cell_loop<DIM, 7>(s, mass, src, dst); break;
        default: std::abort();
      }
  }

  void axpy(long long n, double a, const double *x, double *y)
  {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i) y[i] += a * x[i];
  }
} // namespace

extern "C" {

int oracle_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// SystemMatrix::vmult (transpose != 0: Tvmult), reference include/operators.h:536-583.
// S, D: (degree+1)^2 row-major [q][i]; w: Gauss weights on [0,1];
// metric: NULL (Cartesian box with spacings h) or per cell per q (nsym G entries, JxW);
// src/dst: nb pointers to N doubles.
int oracle_system_vmult(int dim, const int *n_cells, int degree, int nb, const double *Alpha, const double *Beta,
                        int transpose, const double *S, const double *D, const double *w, const double *h,
                        const double *metric, const double *coeff_q, unsigned dirichlet, const double *const *src,
                        double *const *dst, int n_threads)
{
  Setup s;
  s.dim = dim;
  s.n1  = degree + 1;
  s.n_cells = 1;
  s.N       = 1;
  for (int d = 0; d < 3; ++d)
    {
      s.n[d]  = d < dim ? n_cells[d] : 1;
      s.np[d] = d < dim ? degree * n_cells[d] + 1 : 1;
      s.h[d]  = d < dim ? h[d] : 1.0;
      s.n_cells *= s.n[d];
      s.N *= s.np[d];
    }
  s.S = S; s.D = D; s.w = w; s.metric = metric; s.coeff_q = coeff_q; s.dirichlet = dirichlet;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
  std::vector<double> tmp(s.N);
  for (int j = 0; j < nb; ++j) std::memset(dst[j], 0, sizeof(double) * s.N);
  for (int i = 0; i < nb; ++i)
    {
      if (dim == 2) cell_loop_deg<2>(s, false, src[i], tmp.data()); else cell_loop_deg<3>(s, false, src[i], tmp.data());
      for (int j = 0; j < nb; ++j)
        {
          const double a = transpose ? Alpha[i * nb + j] : Alpha[j * nb + i];
          if (a != 0.0) axpy(s.N, a, tmp.data(), dst[j]);
        }
      if (dim == 2) cell_loop_deg<2>(s, true, src[i], tmp.data()); else cell_loop_deg<3>(s, true, src[i], tmp.data());
      for (int j = 0; j < nb; ++j)
        {
          const double b = transpose ? Beta[i * nb + j] : Beta[j * nb + i];
          if (b != 0.0) axpy(s.N, b, tmp.data(), dst[j]);
        }
    }
  return 0;
}

} // extern "C"
