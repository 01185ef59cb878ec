// st_vmult (generic variant): fused matrix-free application of
//     dst_j = sum_i  Alpha(j,i) K src_i + Beta(j,i) M src_i
// in ONE cell loop.  Replaces the reference's 2*nb cell loops + (nb + nnz(Beta)) vector updates of
// SystemMatrix::vmult (reference include/operators.h:536-559) and the cell kernel of
// MatrixFreeOperator (operators.h:1112-1173: gather, evaluate, q-loop, integrate, scatter).
//
// Generic in dimension (2,3), degree (N1 = k+1 <= 7), precision and number of time blocks (runtime,
// <= STFEM_MAX_BLOCKS).  One thread per (cell, time block, column of the cell's tensor grid); the
// column along the last tensor direction lives in registers, contractions along the other
// directions go through shared memory.  The temporal nb x nb contraction happens at the
// quadrature points, so a cell's space-time block is read once and written once.
#pragma once
#include <cuda_runtime.h>

#include "../../include/stfem_b200.h"

namespace stfem
{
  template <typename T, int N1>
  struct ShapeDev
  {
    T S[N1 * N1];  // S[q*N1+i]   values of GLL Lagrange basis at Gauss points
    T Dc[N1 * N1]; // Dc[q*N1+p]  collocation derivative on Gauss points
    T w[N1];       // Gauss weights on [0,1]
  };

  template <typename T, int N1>
  struct VmultArgs
  {
    ShapeDev<T, N1> sh;
    int             n[3];  // cells per direction
    int             np[3]; // dofs per direction
    long long       n_cells;
    int             nb_src, nb_dst;
    const T        *src[STFEM_MAX_BLOCKS];
    T              *dst[STFEM_MAX_BLOCKS];
    const T        *alpha; // device, nb_dst x nb_src row-major (already transposed for Tvmult)
    const T        *beta;
    int             geom_mode; // 0 Cartesian, 1 precomputed metric
    T               h[3];
    const T        *metric;     // geom_mode 1: per cell, per q: DIM*(DIM+1)/2 entries of J^-1 J^-T JxW, then JxW
    const T        *coeff_cell; // optional per-cell Laplace coefficient
    unsigned        dirichlet;  // bit 2*d+side
    int             cells_per_cta;
    int             nbmax;
  };

  template <typename T>
  __device__ __forceinline__ void atomic_add(T *p, T v)
  {
    atomicAdd(p, v);
  }

  // out[q] = sum_a Mx[q*N1+a] * in[a]   (registers)
  template <typename T, int N1, bool TRANSPOSE>
  __device__ __forceinline__ void apply_reg(const T *__restrict__ Mx, const T (&in)[N1], T (&out)[N1])
  {
#pragma unroll
    for (int q = 0; q < N1; ++q)
      {
        T s = 0;
#pragma unroll
        for (int a = 0; a < N1; ++a)
          s += (TRANSPOSE ? Mx[a * N1 + q] : Mx[q * N1 + a]) * in[a];
        out[q] = s;
      }
  }

  template <int DIM, int N1, typename T>
  __global__ void st_vmult_generic_kernel(const VmultArgs<T, N1> a)
  {
    constexpr int NP  = (DIM == 3) ? N1 * N1 : N1; // threads per (cell, block)
    constexpr int NC  = (DIM == 3) ? N1 * N1 * N1 : N1 * N1;
    constexpr int NSYM = DIM * (DIM + 1) / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);

    const int nbmax = a.nbmax;
    const int tid   = threadIdx.x;
    const int slot  = tid / (NP * nbmax);
    const int b     = (tid / NP) % nbmax;
    const int p     = tid % NP;
    const int i     = p % N1;
    const int j     = (DIM == 3) ? p / N1 : 0;

    // shared memory per cell slot: U[nbmax][NC] + E[DIM+1][nbmax][NC]
    const int per_slot = (DIM + 2) * nbmax * NC;
    T        *U        = smem + (size_t)slot * per_slot + b * NC;
    T        *Ebase    = smem + (size_t)slot * per_slot + nbmax * NC; // E[c][b][q]
    // alpha/beta staged after the slots
    T *sAlpha = smem + (size_t)a.cells_per_cta * per_slot;
    T *sBeta  = sAlpha + a.nb_dst * a.nb_src;
    for (int t = tid; t < a.nb_dst * a.nb_src; t += blockDim.x)
      {
        sAlpha[t] = a.alpha[t];
        sBeta[t]  = a.beta[t];
      }

    const long long cell   = (long long)blockIdx.x * a.cells_per_cta + slot;
    const bool      active = cell < a.n_cells && slot < a.cells_per_cta;
    int             cx = 0, cy = 0, cz = 0;
    if (active)
      {
        long long c = cell;
        cx          = (int)(c % a.n[0]);
        c /= a.n[0];
        cy = (int)(c % a.n[1]);
        cz = (DIM == 3) ? (int)(c / a.n[1]) : 0;
      }
    constexpr int K = N1 - 1;
    // global dof coordinates of this thread's column
    const int gx = cx * K + i;
    const int gy = (DIM == 3) ? cy * K + j : 0; // 2D: column runs along y
    // constrained flags for the fixed coordinates
    bool fixed_constrained = false;
    if (((a.dirichlet >> 0) & 1u) && gx == 0) fixed_constrained = true;
    if (((a.dirichlet >> 1) & 1u) && gx == a.np[0] - 1) fixed_constrained = true;
    if (DIM == 3)
      {
        if (((a.dirichlet >> 2) & 1u) && gy == 0) fixed_constrained = true;
        if (((a.dirichlet >> 3) & 1u) && gy == a.np[1] - 1) fixed_constrained = true;
      }
    const int  last_dir  = DIM - 1;
    const int  clast     = (DIM == 3) ? cz : cy;
    const int  np_last   = a.np[last_dir];
    const bool dir_lo    = (a.dirichlet >> (2 * last_dir)) & 1u;
    const bool dir_hi    = (a.dirichlet >> (2 * last_dir + 1)) & 1u;
    const long long stride_last = (DIM == 3) ? (long long)a.np[0] * a.np[1] : (long long)a.np[0];
    const long long base = (DIM == 3) ? ((long long)gx + (long long)a.np[0] * gy) : (long long)gx;

    T u[N1], t[N1];
    // ---------------- gather (read_dof_values: constrained -> 0)
#pragma unroll
    for (int l = 0; l < N1; ++l)
      {
        u[l]         = 0;
        const int gl = clast * K + l;
        const bool c = fixed_constrained || (dir_lo && gl == 0) || (dir_hi && gl == np_last - 1);
        if (active && b < a.nb_src && !c)
          u[l] = a.src[b][base + stride_last * gl];
      }
    __syncthreads(); // alpha/beta staged

    auto sidx = [&](int l, int jj, int ii) -> int {
      return (DIM == 3) ? (l * N1 + jj) * N1 + ii : l * N1 + ii;
    };

    // ---------------- interpolate to quadrature points: last direction in registers
    apply_reg<T, N1, false>(a.sh.S, u, t);
    // x direction through shared memory
#pragma unroll
    for (int l = 0; l < N1; ++l) U[sidx(l, j, i)] = t[l];
    __syncthreads();
#pragma unroll
    for (int l = 0; l < N1; ++l)
      {
        T s = 0;
#pragma unroll
        for (int q = 0; q < N1; ++q) s += a.sh.S[i * N1 + q] * U[sidx(l, j, q)];
        u[l] = s;
      }
    __syncthreads();
    if (DIM == 3)
      {
#pragma unroll
        for (int l = 0; l < N1; ++l) U[sidx(l, j, i)] = u[l];
        __syncthreads();
#pragma unroll
        for (int l = 0; l < N1; ++l)
          {
            T s = 0;
#pragma unroll
            for (int q = 0; q < N1; ++q) s += a.sh.S[j * N1 + q] * U[sidx(l, q, i)];
            t[l] = s;
          }
        __syncthreads();
#pragma unroll
        for (int l = 0; l < N1; ++l) u[l] = t[l];
      }
    // u[l] = value at quadrature point (i, j, l)

    // ---------------- collocation gradient
    T g[DIM][N1];
#pragma unroll
    for (int l = 0; l < N1; ++l) U[sidx(l, j, i)] = u[l];
    __syncthreads();
#pragma unroll
    for (int l = 0; l < N1; ++l)
      {
        T s0 = 0, s1 = 0;
#pragma unroll
        for (int q = 0; q < N1; ++q)
          {
            s0 += a.sh.Dc[i * N1 + q] * U[sidx(l, j, q)];
            if (DIM == 3) s1 += a.sh.Dc[j * N1 + q] * U[sidx(l, q, i)];
          }
        g[0][l] = s0;
        if (DIM == 3) g[1][l] = s1;
      }
    apply_reg<T, N1, false>(a.sh.Dc, u, g[DIM - 1]);

    // ---------------- geometry at the quadrature points, then publish (m, t_x, t_y, t_z)
    {
      const T coef = (active && a.coeff_cell) ? a.coeff_cell[cell] : T(1);
#pragma unroll
      for (int l = 0; l < N1; ++l)
        {
          const int q = sidx(l, j, i);
          T         m, tx[DIM];
          if (a.geom_mode == 0)
            {
              const T wq = a.sh.w[i] * ((DIM == 3) ? a.sh.w[j] : T(1)) * a.sh.w[l];
              T       vol = a.h[0] * a.h[1] * ((DIM == 3) ? a.h[2] : T(1));
              m           = vol * wq * u[l];
#pragma unroll
              for (int d = 0; d < DIM; ++d) tx[d] = coef * (vol * wq / (a.h[d] * a.h[d])) * g[d][l];
            }
          else
            {
              const T *mt = a.metric + ((size_t)(active ? cell : 0) * NC + q) * (NSYM + 1);
              m           = mt[NSYM] * u[l];
              if (DIM == 2)
                {
                  tx[0] = coef * (mt[0] * g[0][l] + mt[1] * g[1][l]);
                  tx[1] = coef * (mt[1] * g[0][l] + mt[2] * g[1][l]);
                }
              else
                {
                  // symmetric storage: xx xy xz yy yz zz
                  tx[0] = coef * (mt[0] * g[0][l] + mt[1] * g[1][l] + mt[2] * g[2][l]);
                  tx[1] = coef * (mt[1] * g[0][l] + mt[3] * g[1][l] + mt[4] * g[2][l]);
                  tx[2] = coef * (mt[2] * g[0][l] + mt[4] * g[1][l] + mt[5] * g[2][l]);
                }
            }
          Ebase[(0 * nbmax + b) * NC + q] = m;
#pragma unroll
          for (int d = 0; d < DIM; ++d) Ebase[((d + 1) * nbmax + b) * NC + q] = tx[d];
        }
    }
    __syncthreads();

    // ---------------- temporal contraction at the quadrature points
    T mq[N1], tq[DIM][N1];
#pragma unroll
    for (int l = 0; l < N1; ++l)
      {
        const int q = sidx(l, j, i);
        T         sm = 0, st[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) st[d] = 0;
        if (b < a.nb_dst)
          for (int ib = 0; ib < a.nb_src; ++ib)
            {
              const T al = sAlpha[b * a.nb_src + ib];
              const T be = sBeta[b * a.nb_src + ib];
              sm += be * Ebase[(0 * nbmax + ib) * NC + q];
#pragma unroll
              for (int d = 0; d < DIM; ++d) st[d] += al * Ebase[((d + 1) * nbmax + ib) * NC + q];
            }
        mq[l] = sm;
#pragma unroll
        for (int d = 0; d < DIM; ++d) tq[d][l] = st[d];
      }
    __syncthreads();

    // ---------------- integrate: transposed collocation gradient (+ mass part)
    T *E1 = Ebase + (1 * nbmax + b) * NC;
    T *E2 = Ebase + (2 * nbmax + b) * NC;
#pragma unroll
    for (int l = 0; l < N1; ++l)
      {
        E1[sidx(l, j, i)] = tq[0][l];
        if (DIM == 3) E2[sidx(l, j, i)] = tq[1][l];
      }
    apply_reg<T, N1, true>(a.sh.Dc, tq[DIM - 1], t);
    __syncthreads();
#pragma unroll
    for (int l = 0; l < N1; ++l)
      {
        T s = mq[l] + t[l];
#pragma unroll
        for (int q = 0; q < N1; ++q)
          {
            s += a.sh.Dc[q * N1 + i] * E1[sidx(l, j, q)];
            if (DIM == 3) s += a.sh.Dc[q * N1 + j] * E2[sidx(l, q, i)];
          }
        u[l] = s;
      }
    // ---------------- transposed interpolation: x, (y), last
#pragma unroll
    for (int l = 0; l < N1; ++l) U[sidx(l, j, i)] = u[l];
    __syncthreads();
#pragma unroll
    for (int l = 0; l < N1; ++l)
      {
        T s = 0;
#pragma unroll
        for (int q = 0; q < N1; ++q) s += a.sh.S[q * N1 + i] * U[sidx(l, j, q)];
        t[l] = s;
      }
    __syncthreads();
    if (DIM == 3)
      {
#pragma unroll
        for (int l = 0; l < N1; ++l) U[sidx(l, j, i)] = t[l];
        __syncthreads();
#pragma unroll
        for (int l = 0; l < N1; ++l)
          {
            T s = 0;
#pragma unroll
            for (int q = 0; q < N1; ++q) s += a.sh.S[q * N1 + j] * U[sidx(l, q, i)];
            t[l] = s;
          }
      }
    apply_reg<T, N1, true>(a.sh.S, t, u);

    // ---------------- scatter (distribute_local_to_global: constrained rows skipped)
    if (active && b < a.nb_dst && !fixed_constrained)
      {
#pragma unroll
        for (int l = 0; l < N1; ++l)
          {
            const int  gl = clast * K + l;
            const bool c  = (dir_lo && gl == 0) || (dir_hi && gl == np_last - 1);
            if (!c) atomic_add(a.dst[b] + base + stride_last * gl, u[l]);
          }
      }
  }
} // namespace stfem
