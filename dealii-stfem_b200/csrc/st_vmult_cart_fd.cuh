// st_vmult, Cartesian variant in FAST-DIAGONALISATION form (3D, square time matrices) - opt-in through
// stfem_op_desc::kernel_variant = 60.  Verified on B200 in round 2 (tests/test_cart_fd_gpu.py, gpu marker): 1.01 ms (FP64) /
// 0.60 ms (FP32) on configs[1] against 0.90 / 0.67 ms of the brick kernel; not on any default path.
//
// Same operator as st_vmult_cart.cuh,
//     dst_j += sum_s ( Alpha(j,s) c_cell K_c + Beta(j,s) M_c ) src_s ,
// with the 1D reference matrices Mh = S^T W S, Kh = D^T W D (reference include/operators.h:1112-1173 on an axis-aligned
// cell) written through their generalised eigen-decomposition  Kh s_q = l_q Mh s_q,  S^T Mh S = I:
//     Mh = V^T V,   Kh = V^T diag(l) V,   V = S^-1 = S^T Mh
//     A_c = vol (V^T (x) V^T (x) V^T) [ Beta + c (l_i/hx^2 + l_j/hy^2 + l_k/hz^2) Alpha ]_(per mode) (V (x) V (x) V).
// 6 one-dimensional sweeps + an nb x nb product per mode instead of 8 sweeps + the temporal contraction at the gather:
// Q4, nb = 2: about 860 FMA per thread instead of 1100 + 104 MUL, ONE field per block through shared memory instead of
// two (P, Q) out and one back, and every thread gathers its own source block only (25 loads instead of 50).  The
// kernel organisation is the one of the Kronecker Vanka (vanka_fd.cuh): a thread owns a y-z plane of a (cell, block),
// lanes run along x; the x phase works on all time blocks of a line so that the mode product stays in registers.
// Constrained (Dirichlet) nodes are read as 0 and never written (FEEvaluation::read_dof_values /
// distribute_local_to_global with zero-boundary AffineConstraints, operators.h:1119-1128).
#pragma once
#include <cmath>

#ifndef STFEM_CART_FD_STANDALONE // tests/cpp/cart_fd_host_emulation.cpp compiles this header as plain host C++ with shims
#include <cuda_runtime.h>

#include "st_vmult_cart.cuh"
#endif

namespace stfem
{
  template <typename T, int N1>
  struct CartFdArgs
  {
    T         V[N1 * N1];   // to modes:   m_q = sum_a V[q*N1+a] u_a
    T         Vt[N1 * N1];  // to nodes:   u_a = sum_q Vt[a*N1+q] m_q
    T         Vtx[N1 * N1]; // the same times the cell volume (applied in the x direction)
    T         lam[3][N1];   // l_q / h_d^2
    int       n[3], np[3];
    long long n_cells;
    int       cells_per_cta;
    unsigned  dirichlet;
    const T  *src[STFEM_MAX_BLOCKS];
    T        *dst[STFEM_MAX_BLOCKS];
    const T  *alpha, *beta; // device, NB x NB row-major
    const T  *coeff_cell;   // optional per-cell Laplace coefficient
  };

  template <typename T, int N1>
  __device__ __forceinline__ void cfd_apply(const T (&Mat)[N1 * N1], const T (&in)[N1], T (&out)[N1])
  {
#pragma unroll
    for (int q = 0; q < N1; ++q)
      {
        T s = T(0);
#pragma unroll
        for (int a = 0; a < N1; ++a) s += Mat[q * N1 + a] * in[a];
        out[q] = s;
      }
  }

  // launch bounds (128, 3) = a budget of 170 registers: Q4, nb = 2 in FP64 needs 136 without spills (128 with 16 bytes of
  // spills under (256, 2)); the launcher picks one-warp CTAs (3 cells x 2 blocks x 5 planes), about 15 resident per SM
  template <int N1, int NB, typename T>
  __global__ void __launch_bounds__(128, 3) k_st_vmult_cart_fd(const __grid_constant__ CartFdArgs<T, N1> a)
  {
    using L           = ExchLayout<N1>;
    constexpr int K   = N1 - 1;
    constexpr int LS  = L::LS;
    constexpr int CBS = L::CBS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *buf = reinterpret_cast<T *>(smem_raw);

    const int       tid  = threadIdx.x;
    const int       tpc  = NB * N1;
    const int       slot = tid / tpc;
    const int       rem  = tid - slot * tpc;
    const int       b    = rem / N1;
    const int       i    = rem - b * N1;
    const int       cb   = tid / N1;
    const long long cell = (long long)blockIdx.x * a.cells_per_cta + slot;
    const bool      active = cell < a.n_cells;
    int             c[3] = {0, 0, 0};
    if (active)
      {
        unsigned cc = (unsigned)cell; // < 2^31 cells: 32-bit divisions
        c[0]        = (int)(cc % (unsigned)a.n[0]);
        cc /= (unsigned)a.n[0];
        c[1] = (int)(cc % (unsigned)a.n[1]);
        c[2] = (int)(cc / (unsigned)a.n[1]);
      }
    bool lo_con[3], hi_con[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
      {
        lo_con[d] = c[d] == 0 && ((a.dirichlet >> (2 * d)) & 1u);
        hi_con[d] = c[d] == a.n[d] - 1 && ((a.dirichlet >> (2 * d + 1)) & 1u);
      }
    const bool      plane_con = (i == 0 && lo_con[0]) || (i == K && hi_con[0]);
    const int       sy   = a.np[0];
    const int       sz   = a.np[0] * a.np[1];
    const long long base = (long long)(c[0] * K + i) + (long long)a.np[0] * ((long long)(c[1] * K) + (long long)a.np[1] * (c[2] * K));

    // ---------------- phase A: gather (constrained nodes read as 0), V in y, V in z
    T x[N1][N1]; // [z][y]
    {
      const T *p = a.src[b] + base;
#pragma unroll
      for (int k = 0; k < N1; ++k)
        {
          const bool kc = (k == 0 && lo_con[2]) || (k == K && hi_con[2]);
#pragma unroll
          for (int jy = 0; jy < N1; ++jy)
            {
              const bool cn = plane_con || kc || (jy == 0 && lo_con[1]) || (jy == K && hi_con[1]);
              x[k][jy]      = (active && !cn) ? p[jy * sy + k * sz] : T(0);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < N1; ++k)
      {
        T t[N1];
        cfd_apply<T, N1>(a.V, x[k], t);
#pragma unroll
        for (int q = 0; q < N1; ++q) x[k][q] = t[q];
      }
    {
      T *pb = buf + cb * CBS + i;
#pragma unroll
      for (int jy = 0; jy < N1; ++jy)
        {
          T in[N1], t[N1];
#pragma unroll
          for (int k = 0; k < N1; ++k) in[k] = x[k][jy];
          cfd_apply<T, N1>(a.V, in, t);
#pragma unroll
          for (int q = 0; q < N1; ++q) pb[(q * N1 + jy) * LS] = t[q];
        }
    }
    __syncthreads();

    // ---------------- phase B: per (my, mz) line all NB blocks: V in x, mode product, vol V^T in x
    {
      T al[NB * NB], be[NB * NB];
#pragma unroll
      for (int e = 0; e < NB * NB; ++e)
        {
          al[e] = __ldg(a.alpha + e);
          be[e] = __ldg(a.beta + e);
        }
      const T coef = (active && a.coeff_cell) ? __ldg(a.coeff_cell + cell) : T(1);
      for (int pos = rem; pos < N1 * N1; pos += tpc)
        {
          const int my = pos % N1, mz = pos / N1;
          const T   lyz = a.lam[1][my] + a.lam[2][mz];
          T         t[NB][N1];
#pragma unroll
          for (int bb = 0; bb < NB; ++bb)
            {
              const T *pl = buf + (slot * NB + bb) * CBS + pos * LS;
              T        in[N1];
#pragma unroll
              for (int xx = 0; xx < N1; ++xx) in[xx] = pl[xx];
              cfd_apply<T, N1>(a.V, in, t[bb]);
            }
          T u[NB][N1];
#pragma unroll
          for (int mx = 0; mx < N1; ++mx)
            {
              const T l = coef * (a.lam[0][mx] + lyz);
#pragma unroll
              for (int r = 0; r < NB; ++r)
                {
                  T s = T(0);
#pragma unroll
                  for (int cc = 0; cc < NB; ++cc) s += (be[r * NB + cc] + l * al[r * NB + cc]) * t[cc][mx];
                  u[r][mx] = s;
                }
            }
#pragma unroll
          for (int bb = 0; bb < NB; ++bb)
            {
              T out[N1];
              cfd_apply<T, N1>(a.Vtx, u[bb], out);
              T *pl = buf + (slot * NB + bb) * CBS + pos * LS;
#pragma unroll
              for (int xx = 0; xx < N1; ++xx) pl[xx] = out[xx];
            }
        }
    }
    __syncthreads();

    // ---------------- phase C: V^T in z, V^T in y, scatter-add (constrained rows skipped)
    {
      const T *pb = buf + cb * CBS + i;
#pragma unroll
      for (int jy = 0; jy < N1; ++jy)
        {
          T in[N1], t[N1];
#pragma unroll
          for (int q = 0; q < N1; ++q) in[q] = pb[(q * N1 + jy) * LS];
          cfd_apply<T, N1>(a.Vt, in, t);
#pragma unroll
          for (int k = 0; k < N1; ++k) x[k][jy] = t[k];
        }
    }
    if (active && !plane_con)
      {
        T *d = a.dst[b] + base;
#pragma unroll
        for (int k = 0; k < N1; ++k)
          {
            T t[N1];
            cfd_apply<T, N1>(a.Vt, x[k], t);
            const bool kc = (k == 0 && lo_con[2]) || (k == K && hi_con[2]);
#pragma unroll
            for (int jy = 0; jy < N1; ++jy)
              {
                const bool cn = kc || (jy == 0 && lo_con[1]) || (jy == K && hi_con[1]);
                if (!cn) atomicAdd(d + jy * sy + k * sz, t[jy]);
              }
          }
      }
  }

  // ---- host set-up: V and the eigenvalues of the reference-cell pencil (Kh, Mh); n <= 8, double
  namespace cartfd_host
  {
    // K s = l M s with M SPD: V = S^-1 (rows = modes) and l, through Cholesky M = L L^T and a cyclic Jacobi iteration on
    // L^-1 K L^-T = Q diag(l) Q^T:  S = L^-T Q,  V = S^-1 = Q^T L^T
    inline void pencil_modes(const double *M, const double *Kin, int n, double *V, double *lam)
    {
      double Lc[64] = {0}, Li[64] = {0}, A[64] = {0}, T1[64] = {0}, Q[64] = {0};
      for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j)
          {
            double s = M[i * n + j];
            for (int k = 0; k < j; ++k) s -= Lc[i * n + k] * Lc[j * n + k];
            Lc[i * n + j] = i == j ? std::sqrt(s) : s / Lc[j * n + j];
          }
      for (int c = 0; c < n; ++c)
        for (int i = 0; i < n; ++i)
          {
            double s = i == c ? 1.0 : 0.0;
            for (int k = 0; k < i; ++k) s -= Lc[i * n + k] * Li[k * n + c];
            Li[i * n + c] = s / Lc[i * n + i];
          }
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
          for (int k = 0; k < n; ++k) T1[i * n + j] += Li[i * n + k] * Kin[k * n + j];
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
          for (int k = 0; k < n; ++k) A[i * n + j] += T1[i * n + k] * Li[j * n + k];
      for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) A[i * n + j] = A[j * n + i] = 0.5 * (A[i * n + j] + A[j * n + i]);
      for (int i = 0; i < n; ++i) Q[i * n + i] = 1.0;
      for (int sweep = 0; sweep < 100; ++sweep)
        {
          double off = 0, diag = 0;
          for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) (i == j ? diag : off) += A[i * n + j] * A[i * n + j];
          if (off <= 1e-32 * (diag + 1e-300)) break;
          for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q)
              {
                const double apq = A[p * n + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                const double t     = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < n; ++k)
                  {
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - sn * akq;
                    A[k * n + q] = sn * akp + c * akq;
                  }
                for (int k = 0; k < n; ++k)
                  {
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - sn * aqk;
                    A[q * n + k] = sn * apk + c * aqk;
                  }
                for (int k = 0; k < n; ++k)
                  {
                    const double qkp = Q[k * n + p], qkq = Q[k * n + q];
                    Q[k * n + p] = c * qkp - sn * qkq;
                    Q[k * n + q] = sn * qkp + c * qkq;
                  }
              }
        }
      for (int q = 0; q < n; ++q) lam[q] = A[q * n + q];
      // V = Q^T L^T :  V[q][a] = sum_k Q[k][q] L[a][k]
      for (int q = 0; q < n; ++q)
        for (int a = 0; a < n; ++a)
          {
            double s = 0;
            for (int k = 0; k < n; ++k) s += Q[k * n + q] * Lc[a * n + k];
            V[q * n + a] = s;
          }
    }
  } // namespace cartfd_host
} // namespace stfem
