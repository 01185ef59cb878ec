// Device kernels around the solve ("next" rows of SURVEY.md §8f): source-term integration, nodal
// interpolation of the analytic functions and the space-time error functional.
//   integrate  : VectorTools::create_right_hand_side as used by tests/tp_01.cc:382-392
//   interpolate: VectorTools::interpolate                      tests/tp_01.cc:393-408
//   error      : ErrorCalculator::evaluate_error               include/exact_solution.h:533-633
// Analytic functions (include/exact_solution.h:27-197), selected by id:
//   0 zero, 1 ExactSolution (u), 2 RHSFunction (heat), 3 wave::ExactSolutionV, 4 wave::RHSFunction
#pragma once
#include "op.hpp"
#include "vec.cuh"

namespace stfem
{
  constexpr double ST_PI = 3.14159265358979323846;

  __device__ inline double analytic_value(int fid, int dim, const double *x, double t, double f)
  {
    if (fid == 0) return 0.0;
    double prod = 1.0;
    for (int d = 0; d < dim; ++d) prod *= sin(2 * ST_PI * f * x[d]);
    switch (fid)
      {
        case 1: return sin(2 * ST_PI * f * t) * prod;
        case 2: return (dim * 4 * ST_PI * ST_PI * f * f * sin(2 * ST_PI * f * t) + 2 * ST_PI * f * cos(2 * ST_PI * f * t)) * prod;
        case 3: return 2 * ST_PI * f * cos(2 * ST_PI * f * t) * prod;
        case 4: return pow(2.0, (double)dim) * (ST_PI * f) * (ST_PI * f) * sin(2 * ST_PI * f * t) * prod;
        default: return 0.0;
      }
  }

  // all analytic functions are  amp(t) * prod_d sin(2 pi f x_d): the time factor alone
  __host__ __device__ inline double analytic_amp(int fid, int dim, double t, double f)
  {
    switch (fid)
      {
        case 1: return sin(2 * ST_PI * f * t);
        case 2: return dim * 4 * ST_PI * ST_PI * f * f * sin(2 * ST_PI * f * t) + 2 * ST_PI * f * cos(2 * ST_PI * f * t);
        case 3: return 2 * ST_PI * f * cos(2 * ST_PI * f * t);
        case 4: return pow(2.0, (double)dim) * (ST_PI * f) * (ST_PI * f) * sin(2 * ST_PI * f * t);
        default: return 0.0;
      }
  }

  __device__ inline void exact_gradient(int dim, const double *x, double t, double f, double *g)
  {
    const double tv = 2 * ST_PI * f * sin(2 * ST_PI * f * t);
    for (int i = 0; i < dim; ++i)
      {
        double v = tv;
        for (int j = 0; j < dim; ++j) v *= (i == j ? cos(2 * ST_PI * f * x[j]) : sin(2 * ST_PI * f * x[j]));
        g[i] = v;
      }
  }

  struct AsmGeom
  {
    int           dim, n1, nq1; // FE nodes per direction, quadrature points per direction
    int           n[3], np[3];
    long long     n_cells;
    unsigned      dirichlet;
    double        lower[3], h[3];
    const double *vertices; // device, or null for a Cartesian box
    double        gll[7];   // FE support points
    double        xq[8], wq[8];
    double        S[56], D[56]; // [q*n1+i] values / derivatives of the GLL basis at xq
    // Cartesian meshes: sin(2 pi f x) at the quadrature coordinates of every cell, per direction
    // ([cell * nq1 + q]); null = evaluate the function at every quadrature point
    const double *sin_tab[3] = {nullptr, nullptr, nullptr};
  };

  // MappingQ1: point and Jacobian at reference coordinates xi of a cell
  __device__ inline void map_q1(const AsmGeom &g, const int *c, const double *xi, double *x, double (*J)[3])
  {
    const int dim = g.dim;
    if (!g.vertices)
      {
        for (int a = 0; a < dim; ++a)
          {
            x[a] = g.lower[a] + g.h[a] * (c[a] + xi[a]);
            for (int b = 0; b < dim; ++b) J[a][b] = a == b ? g.h[a] : 0.0;
          }
        return;
      }
    for (int a = 0; a < dim; ++a)
      {
        x[a] = 0;
        for (int b = 0; b < dim; ++b) J[a][b] = 0;
      }
    for (int v = 0; v < (1 << dim); ++v)
      {
        const int vb[3] = {v & 1, (v >> 1) & 1, (v >> 2) & 1};
        long long vid   = (long long)(c[0] + vb[0]) + (long long)(g.n[0] + 1) * (c[1] + vb[1]);
        if (dim == 3) vid += (long long)(g.n[0] + 1) * (g.n[1] + 1) * (c[2] + vb[2]);
        double N[3], dN[3];
        for (int a = 0; a < 3; ++a)
          {
            N[a]  = a < dim ? (vb[a] ? xi[a] : 1 - xi[a]) : 1.0;
            dN[a] = vb[a] ? 1.0 : -1.0;
          }
        const double sh = N[0] * N[1] * N[2];
        for (int a = 0; a < dim; ++a) x[a] += g.vertices[vid * dim + a] * sh;
        for (int b = 0; b < dim; ++b)
          {
            double s = dN[b];
            for (int cc = 0; cc < dim; ++cc)
              if (cc != b) s *= N[cc];
            for (int a = 0; a < dim; ++a) J[a][b] += g.vertices[vid * dim + a] * s;
          }
      }
  }

  __device__ inline double det_inv(int dim, double (*J)[3], double (*inv)[3])
  {
    if (dim == 2)
      {
        const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        inv[0][0] = J[1][1] / det; inv[0][1] = -J[0][1] / det; inv[1][0] = -J[1][0] / det; inv[1][1] = J[0][0] / det;
        return det;
      }
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2], c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    inv[0][0] = c00 / det; inv[1][0] = c01 / det; inv[2][0] = c02 / det;
    inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
    inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
    inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
    inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
    inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
    inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
    return det;
  }

  __device__ inline bool asm_constrained(const AsmGeom &g, int ix, int iy, int iz)
  {
    const unsigned d = g.dirichlet;
    if ((d & 1u) && ix == 0) return true;
    if ((d & 2u) && ix == g.np[0] - 1) return true;
    if ((d & 4u) && iy == 0) return true;
    if ((d & 8u) && iy == g.np[1] - 1) return true;
    if (g.dim == 3 && (((d & 16u) && iz == 0) || ((d & 32u) && iz == g.np[2] - 1))) return true;
    return false;
  }

  // rhs_i += scale * sum_q f(x_q, t) phi_i(x_q) JxW ; one CTA per cell: f JxW is evaluated ONCE per quadrature
  // point into shared memory, then integrated by sum factorisation (one tensor direction per pass);
  // constrained rows skipped
  static __global__ void k_integrate_function(AsmGeom g, int fid, double t, double freq, double scale, double *__restrict__ rhs)
  {
    const int dim = g.dim, n1 = g.n1, nq1 = g.nq1, k = n1 - 1;
    const int nq = dim == 3 ? nq1 * nq1 * nq1 : nq1 * nq1;
    const int nc = dim == 3 ? n1 * n1 * n1 : n1 * n1;
    __shared__ double bufA[512], bufB[512]; // nq1, n1 <= 8
    for (long long cell = blockIdx.x; cell < g.n_cells; cell += gridDim.x)
      {
        const int c[3] = {(int)(cell % g.n[0]), (int)((cell / g.n[0]) % g.n[1]), dim == 3 ? (int)(cell / ((long long)g.n[0] * g.n[1])) : 0};
        __syncthreads();
        for (int q = threadIdx.x; q < nq; q += blockDim.x)
          {
            const int    qi[3] = {q % nq1, (q / nq1) % nq1, dim == 3 ? q / (nq1 * nq1) : 0};
            const double xi[3] = {g.xq[qi[0]], g.xq[qi[1]], dim == 3 ? g.xq[qi[2]] : 0.0};
            const double w = g.wq[qi[0]] * g.wq[qi[1]] * (dim == 3 ? g.wq[qi[2]] : 1.0);
            if (g.sin_tab[0])
              {
                double v = (fid < 0 ? 1.0 : analytic_amp(fid, dim, t, freq)) * g.h[0] * g.h[1] * (dim == 3 ? g.h[2] : 1.0); // fid < 0: spatial factor only
                for (int d = 0; d < dim; ++d) v *= g.sin_tab[d][c[d] * nq1 + qi[d]];
                bufA[q] = v * w;
                continue;
              }
            double       x[3], J[3][3], inv[3][3];
            map_q1(g, c, xi, x, J);
            const double det = det_inv(dim, J, inv);
            bufA[q]          = analytic_value(fid, dim, x, t, freq) * det * w;
          }
        __syncthreads();
        // x: A[qz][qy][qx] -> B[qz][qy][lx]
        const int rows_x = dim == 3 ? nq1 * nq1 : nq1;
        for (int o = threadIdx.x; o < rows_x * n1; o += blockDim.x)
          {
            const int row = o / n1, lx = o % n1;
            double    s   = 0;
            for (int qx = 0; qx < nq1; ++qx) s += g.S[qx * n1 + lx] * bufA[row * nq1 + qx];
            bufB[o] = s;
          }
        __syncthreads();
        // y: B[qz][qy][lx] -> A[qz][ly][lx]
        const int nz_q = dim == 3 ? nq1 : 1;
        for (int o = threadIdx.x; o < nz_q * n1 * n1; o += blockDim.x)
          {
            const int lx = o % n1, ly = (o / n1) % n1, qz = o / (n1 * n1);
            double    s  = 0;
            for (int qy = 0; qy < nq1; ++qy) s += g.S[qy * n1 + ly] * bufB[(qz * nq1 + qy) * n1 + lx];
            bufA[o] = s;
          }
        __syncthreads();
        for (int l = threadIdx.x; l < nc; l += blockDim.x)
          {
            const int li[3] = {l % n1, (l / n1) % n1, dim == 3 ? l / (n1 * n1) : 0};
            double    s;
            if (dim == 3)
              {
                s = 0;
                for (int qz = 0; qz < nq1; ++qz) s += g.S[qz * n1 + li[2]] * bufA[(qz * n1 + li[1]) * n1 + li[0]];
              }
            else
              s = bufA[l];
            const int gi[3] = {c[0] * k + li[0], c[1] * k + li[1], c[2] * k + li[2]};
            if (asm_constrained(g, gi[0], gi[1], gi[2])) continue;
            const long long dof = (long long)gi[0] + (long long)g.np[0] * (gi[1] + (long long)g.np[1] * gi[2]);
            atomicAdd(rhs + dof, scale * s);
          }
      }
  }

  // dst_i = f(support point i, t) ; one thread per global dof
  static __global__ void k_interpolate_function(AsmGeom g, int fid, double t, double freq, double *__restrict__ dst)
  {
    const int       dim = g.dim, k = g.n1 - 1;
    const long long N   = (long long)g.np[0] * g.np[1] * g.np[2];
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < N; gid += (long long)gridDim.x * blockDim.x)
      {
        const int gi[3] = {(int)(gid % g.np[0]), (int)((gid / g.np[0]) % g.np[1]), dim == 3 ? (int)(gid / ((long long)g.np[0] * g.np[1])) : 0};
        int       c[3], li[3];
        double    xi[3] = {0, 0, 0};
        for (int a = 0; a < dim; ++a)
          {
            c[a] = gi[a] / k;
            if (c[a] > g.n[a] - 1) c[a] = g.n[a] - 1;
            li[a] = gi[a] - c[a] * k;
            xi[a] = g.gll[li[a]];
          }
        if (dim == 2) { c[2] = 0; }
        double x[3], J[3][3];
        map_q1(g, c, xi, x, J);
        dst[gid] = analytic_value(fid, dim, x, t, freq);
      }
  }

  // error of ONE spatial function u_h (already combined in time) against ExactSolution at time t:
  // out[0] += sum (u_h-u)^2 JxW, out[1] = max |u_h-u|, out[2] += sum |grad u_h - grad u|^2 JxW
  static __global__ void k_error(AsmGeom g, const double *__restrict__ u, double t, double freq, double *__restrict__ out)
  {
    const int dim = g.dim, n1 = g.n1, nq1 = g.nq1, k = n1 - 1;
    const int nq = dim == 3 ? nq1 * nq1 * nq1 : nq1 * nq1;
    const long long total = g.n_cells * nq;
    double l2 = 0, h1 = 0, l8 = 0;
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x)
      {
        const long long cell = gid / nq;
        const int       q    = (int)(gid % nq);
        const int c[3]  = {(int)(cell % g.n[0]), (int)((cell / g.n[0]) % g.n[1]), dim == 3 ? (int)(cell / ((long long)g.n[0] * g.n[1])) : 0};
        const int qi[3] = {q % nq1, (q / nq1) % nq1, dim == 3 ? q / (nq1 * nq1) : 0};
        const double xi[3] = {g.xq[qi[0]], g.xq[qi[1]], dim == 3 ? g.xq[qi[2]] : 0.0};
        double       x[3], J[3][3], inv[3][3];
        map_q1(g, c, xi, x, J);
        const double det = det_inv(dim, J, inv);
        double       w   = g.wq[qi[0]] * g.wq[qi[1]] * (dim == 3 ? g.wq[qi[2]] : 1.0);
        double       uh = 0, gr[3] = {0, 0, 0};
        const int    nz = dim == 3 ? n1 : 1;
        for (int lz = 0; lz < nz; ++lz)
          for (int ly = 0; ly < n1; ++ly)
            for (int lx = 0; lx < n1; ++lx)
              {
                const long long dof = (long long)(c[0] * k + lx) + (long long)g.np[0] * ((c[1] * k + ly) + (long long)g.np[1] * (c[2] * k + lz));
                const double    val = u[dof];
                const double sx = g.S[qi[0] * n1 + lx], sy = g.S[qi[1] * n1 + ly], sz = dim == 3 ? g.S[qi[2] * n1 + lz] : 1.0;
                const double dx = g.D[qi[0] * n1 + lx], dy = g.D[qi[1] * n1 + ly], dz = dim == 3 ? g.D[qi[2] * n1 + lz] : 0.0;
                uh += val * sx * sy * sz;
                gr[0] += val * dx * sy * sz;
                gr[1] += val * sx * dy * sz;
                gr[2] += val * sx * sy * dz;
              }
        double ge[3];
        exact_gradient(dim, x, t, freq, ge);
        const double diff = uh - analytic_value(1, dim, x, t, freq);
        double       gd2  = 0;
        for (int a = 0; a < dim; ++a)
          {
            double gra = 0;
            for (int b = 0; b < dim; ++b) gra += inv[b][a] * gr[b]; // J^-T grad_ref
            gd2 += (gra - ge[a]) * (gra - ge[a]);
          }
        l2 += diff * diff * det * w;
        h1 += gd2 * det * w;
        l8 = fmax(l8, fabs(diff));
      }
    for (int o = 16; o > 0; o >>= 1)
      {
        l2 += __shfl_xor_sync(0xffffffffu, l2, o);
        h1 += __shfl_xor_sync(0xffffffffu, h1, o);
        l8 = fmax(l8, __shfl_xor_sync(0xffffffffu, l8, o));
      }
    if ((threadIdx.x & 31) == 0)
      {
        atomicAdd(out + 0, l2);
        atomicAdd(out + 2, h1);
        // max of non-negative doubles via integer compare
        atomicMax((unsigned long long *)(out + 1), (unsigned long long)__double_as_longlong(l8));
      }
  }


  // Diagonals of the spatial operators (MatrixFreeTools::compute_diagonal as used at include/operators.h:1092-1110):
  //   dM_i = sum_cells sum_q phi_i(x_q)^2 JxW,   dK_i = sum_cells sum_q c(x_q) |J^-T grad phi_i(x_q)|^2 JxW
  // One CTA per cell: the metric c J^-1 J^-T JxW (6 numbers) and JxW of every quadrature point go to shared memory once,
  // then one thread per local node sums over the quadrature points.  Constrained rows stay 0.  Set-up work (called once
  // per operator), not a hot kernel.  coeff_cell / coeff_q: optional Laplace coefficient (per cell / per cell and q-point).
  static __global__ void k_diagonal(AsmGeom g, const double *__restrict__ coeff_cell, const double *__restrict__ coeff_q,
                                    double *__restrict__ dK, double *__restrict__ dM)
  {
    const int dim = g.dim, n1 = g.n1, nq1 = g.nq1, k = n1 - 1;
    const int nq = dim == 3 ? nq1 * nq1 * nq1 : nq1 * nq1;
    const int nc = dim == 3 ? n1 * n1 * n1 : n1 * n1;
    __shared__ double met[512 * 7]; // nq1 <= 8: [q][Gxx Gxy Gxz Gyy Gyz Gzz JxW]
    for (long long cell = blockIdx.x; cell < g.n_cells; cell += gridDim.x)
      {
        const int c[3] = {(int)(cell % g.n[0]), (int)((cell / g.n[0]) % g.n[1]), dim == 3 ? (int)(cell / ((long long)g.n[0] * g.n[1])) : 0};
        __syncthreads();
        for (int q = threadIdx.x; q < nq; q += blockDim.x)
          {
            const int    qi[3] = {q % nq1, (q / nq1) % nq1, dim == 3 ? q / (nq1 * nq1) : 0};
            const double xi[3] = {g.xq[qi[0]], g.xq[qi[1]], dim == 3 ? g.xq[qi[2]] : 0.0};
            const double w = g.wq[qi[0]] * g.wq[qi[1]] * (dim == 3 ? g.wq[qi[2]] : 1.0);
            double       x[3], J[3][3], inv[3][3];
            map_q1(g, c, xi, x, J);
            const double jxw = det_inv(dim, J, inv) * w;
            double       cl  = coeff_cell ? coeff_cell[cell] : 1.0;
            if (coeff_q) cl *= coeff_q[cell * nq + q];
            double *m = met + q * 7;
            int     o = 0;
            for (int a = 0; a < 3; ++a)
              for (int b = a; b < 3; ++b, ++o)
                {
                  double s = 0.0;
                  if (a < dim && b < dim)
                    for (int e = 0; e < dim; ++e) s += inv[a][e] * inv[b][e];
                  m[o] = s * jxw * cl;
                }
            m[6] = jxw;
          }
        __syncthreads();
        for (int l = threadIdx.x; l < nc; l += blockDim.x)
          {
            const int li[3] = {l % n1, (l / n1) % n1, dim == 3 ? l / (n1 * n1) : 0};
            const int gi[3] = {c[0] * k + li[0], c[1] * k + li[1], c[2] * k + li[2]};
            if (asm_constrained(g, gi[0], gi[1], gi[2])) continue;
            double sk = 0.0, sm = 0.0;
            for (int q = 0; q < nq; ++q)
              {
                const int    qi[3] = {q % nq1, (q / nq1) % nq1, dim == 3 ? q / (nq1 * nq1) : 0};
                const double sx = g.S[qi[0] * n1 + li[0]], sy = g.S[qi[1] * n1 + li[1]], sz = dim == 3 ? g.S[qi[2] * n1 + li[2]] : 1.0;
                const double gx = g.D[qi[0] * n1 + li[0]] * sy * sz, gy = sx * g.D[qi[1] * n1 + li[1]] * sz,
                             gz = dim == 3 ? sx * sy * g.D[qi[2] * n1 + li[2]] : 0.0;
                const double *m = met + q * 7;
                const double  v = sx * sy * sz;
                sm += v * v * m[6];
                sk += gx * (m[0] * gx + 2.0 * (m[1] * gy + m[2] * gz)) + gy * (m[3] * gy + 2.0 * m[4] * gz) + m[5] * gz * gz;
              }
            const long long dof = (long long)gi[0] + (long long)g.np[0] * (gi[1] + (long long)g.np[1] * gi[2]);
            atomicAdd(dK + dof, sk);
            atomicAdd(dM + dof, sm);
          }
      }
  }

  // out_i = a dK_i + b dM_i in the operator's number type
  template <typename T>
  static __global__ void k_diag_combine(long long n, double a, double b, const double *__restrict__ dK, const double *__restrict__ dM,
                                        T *__restrict__ out)
  {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      out[i] = (T)(a * dK[i] + b * dM[i]);
  }

  inline void fill_asm_geom(AsmGeom &g, const stfem_mesh *m, int degree, int nq1)
  {
    g.dim = m->dim; g.n1 = degree + 1; g.nq1 = nq1; g.n_cells = m->n_cells; g.dirichlet = m->dirichlet;
    g.vertices = m->d_vertices;
    for (int d = 0; d < 3; ++d)
      {
        g.n[d]     = m->n[d];
        g.np[d]    = d < m->dim ? degree * m->n[d] + 1 : 1;
        g.lower[d] = m->lower[d];
        g.h[d]     = d < m->dim ? (m->upper[d] - m->lower[d]) / m->n[d] : 1.0;
      }
    const Rule l = gauss_lobatto(degree + 1), q = gauss(nq1);
    for (int i = 0; i <= degree; ++i) g.gll[i] = l.x[i];
    for (int i = 0; i < nq1; ++i)
      {
        g.xq[i] = q.x[i];
        g.wq[i] = q.w[i];
        for (int j = 0; j <= degree; ++j)
          {
            g.S[i * g.n1 + j] = lagrange_value(l.x, j, q.x[i]);
            g.D[i * g.n1 + j] = lagrange_deriv(l.x, j, q.x[i]);
          }
      }
  }
} // namespace stfem
