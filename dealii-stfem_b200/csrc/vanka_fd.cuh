// Cell-patch Vanka smoother in Kronecker / fast-diagonalisation form (3D, Cartesian levels with a constant
// coefficient).  Same operator as PreconditionVanka of the reference (include/stmg.h:745-872) with the patch
// matrices of restrict_to_full_matrices_ (include/compute_block_matrix.h:50-139), up to rounding:
//
//   B_c = (I (x) D) [ Beta (x) M3 + Alpha (x) K3 ],   D = valence,  M3 = Mx (x) My (x) Mz,
//   K3 = Kx (x) My (x) Mz + Mx (x) Ky (x) Mz + Mx (x) My (x) Kz
// where Md, Kd are the ASSEMBLED 1D matrices restricted to the cell's nodes in direction d (contribution of the
// neighbouring cell on a shared end node included, constrained end nodes decoupled) and D factorises as well.
// With the generalised eigen-decompositions  Sd^T Md Sd = I,  Sd^T Kd Sd = diag(lambda_d):
//   B_c^-1 = (I (x) S3) [ Beta + (lx_i + ly_j + lz_k) Alpha ]^-1_(per mode) (I (x) S3^T) (I (x) D^-1),  S3 = Sx (x) Sy (x) Sz.
// The reference stores and streams the dense (nb n_c)^2 inverse (250 KB per cell for Q4, nb = 2: HBM-bound,
// SURVEY.md §8a); here the apply is 6 one-dimensional sweeps + an nb x nb product per mode: ~8x fewer flops,
// no patch matrix traffic.  Thread mapping as in st_vmult_cart.cuh: a thread owns a y-z plane of a (cell, block),
// one shared-memory exchange to the x lines and one back; the x phase handles all time blocks of a line so the
// nb x nb mode product stays in registers.
// Constrained (Dirichlet) nodes are decoupled from the free ones exactly as in the dense form; their own output
// is 0 when the source is 0 there, which is the only case the V-cycle / FGMRES produce.
#pragma once
#include <cuda_runtime.h>

#include "st_vmult_cart.cuh"

namespace stfem
{
  template <typename T, int N1>
  struct VankaFdArgs
  {
    // [direction][class][N1*N1]; class: 0 first cell, 1 interior, 2 last, 3 single.
    // ST[q*N1+a] = S[a][q] (to modes), S[a*N1+q] (back to nodes)
    T         ST[3][4][N1 * N1];
    T         S[3][4][N1 * N1];
    int       n[3], np[3];
    long long n_cells;
    int       nb, cells_per_cta;
    unsigned  dirichlet;
    unsigned  neighbor_mask; // bit 2*d+side: a neighbouring RANK continues the mesh there (multi-GPU)
    const T  *src;           // block b at src + b*N
    T        *dst;
    long long N;
    const T  *modes;         // [type 64][mz][my][mx][nb*nb]  inverse mode matrices, row-major
    T         scale;         // dst += scale * B^-1 src
  };

  template <typename T, int N1>
  __device__ __forceinline__ void fd_apply(const T (&Mat)[N1 * N1], const T (&in)[N1], T (&out)[N1])
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

  // INTERIOR (all cells of the CTA are interior cells in every direction - the bulk of the mesh): the class-1 matrices
  // are immediate constant-bank operands.  Otherwise the matrix of the thread's class is addressed at run time (LDC per
  // entry): slower, but only the CTAs touching the boundary take this path and the kernel stays small (one unrolled copy
  // of every sweep instead of four).
  template <typename T, int N1, bool INTERIOR>
  __device__ __forceinline__ void fd_apply_cls(const T (&Mats)[4][N1 * N1], int cls, const T (&in)[N1], T (&out)[N1])
  {
    if (INTERIOR)
      fd_apply<T, N1>(Mats[1], in, out);
    else
      {
        const T *M = Mats[cls];
#pragma unroll
        for (int q = 0; q < N1; ++q)
          {
            T s = T(0);
#pragma unroll
            for (int a = 0; a < N1; ++a) s += M[q * N1 + a] * in[a];
            out[q] = s;
          }
      }
  }

  template <int N1, int NB, typename T, bool INTERIOR>
  __device__ __forceinline__ void vanka_fd_body(const VankaFdArgs<T, N1> &a, const int (&c)[3], bool active);

  template <int N1, int NB, typename T>
  __global__ void __launch_bounds__(256, 2) k_vanka_fd(const __grid_constant__ VankaFdArgs<T, N1> a)
  {
    const int       tpc    = NB * N1;
    const int       slot   = threadIdx.x / tpc;
    const long long cell   = (long long)blockIdx.x * a.cells_per_cta + slot;
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
    bool interior = true;
#pragma unroll
    for (int d = 0; d < 3; ++d)
      {
        const bool has_lo = c[d] > 0 || ((a.neighbor_mask >> (2 * d)) & 1u);
        const bool has_hi = c[d] < a.n[d] - 1 || ((a.neighbor_mask >> (2 * d + 1)) & 1u);
        interior          = interior && has_lo && has_hi;
      }
    if (__syncthreads_and(interior || !active ? 1 : 0) && a.n_cells > 0)
      vanka_fd_body<N1, NB, T, true>(a, c, active);
    else
      vanka_fd_body<N1, NB, T, false>(a, c, active);
  }

  template <int N1, int NB, typename T, bool INTERIOR>
  __device__ __forceinline__ void vanka_fd_body(const VankaFdArgs<T, N1> &a, const int (&c)[3], bool active)
  {
    using L           = ExchLayout<N1>;
    constexpr int K   = N1 - 1;
    constexpr int LS  = L::LS;
    constexpr int CBS = L::CBS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *buf = reinterpret_cast<T *>(smem_raw);

    const int tid  = threadIdx.x;
    const int tpc  = NB * N1;
    const int slot = tid / tpc;
    const int rem  = tid - slot * tpc;
    const int b    = rem / N1;
    const int i    = rem - b * N1;
    const int cb   = tid / N1;

    int  cls[3];
    bool lo_shared[3], hi_shared[3], lo_con[3], hi_con[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
      {
        const bool has_lo = c[d] > 0 || ((a.neighbor_mask >> (2 * d)) & 1u);
        const bool has_hi = c[d] < a.n[d] - 1 || ((a.neighbor_mask >> (2 * d + 1)) & 1u);
        cls[d]            = has_lo ? (has_hi ? 1 : 2) : (has_hi ? 0 : 3);
        lo_shared[d]      = has_lo;
        hi_shared[d]      = has_hi;
        lo_con[d]         = !has_lo && ((a.dirichlet >> (2 * d)) & 1u);
        hi_con[d]         = !has_hi && ((a.dirichlet >> (2 * d + 1)) & 1u);
      }
    const int       sy   = a.np[0];
    const int       sz   = a.np[0] * a.np[1];
    const long long base = (long long)(c[0] * K + i) + (long long)a.np[0] * ((long long)(c[1] * K) + (long long)a.np[1] * (c[2] * K));

    // ---------------- phase A: gather, D^-1, S_y^T, S_z^T
    T x[N1][N1]; // [z][y]
    {
      const T  wx = ((i == 0 && lo_shared[0]) || (i == K && hi_shared[0])) ? T(0.5) : T(1);
      const T *p  = a.src + (size_t)b * a.N + base;
#pragma unroll
      for (int k = 0; k < N1; ++k)
#pragma unroll
        for (int jy = 0; jy < N1; ++jy)
          {
            T w = wx;
            if ((jy == 0 && lo_shared[1]) || (jy == K && hi_shared[1])) w *= T(0.5);
            if ((k == 0 && lo_shared[2]) || (k == K && hi_shared[2])) w *= T(0.5);
            x[k][jy] = active ? w * p[jy * sy + k * sz] : T(0);
          }
    }
#pragma unroll
    for (int k = 0; k < N1; ++k)
      {
        T t[N1];
        fd_apply_cls<T, N1, INTERIOR>(a.ST[1], cls[1], x[k], t);
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
          fd_apply_cls<T, N1, INTERIOR>(a.ST[2], cls[2], in, t);
#pragma unroll
          for (int q = 0; q < N1; ++q) pb[(q * N1 + jy) * LS] = t[q];
        }
    }
    __syncthreads();

    // ---------------- phase B: per (my, mz) position all NB lines: S_x^T, mode product, S_x
    {
      int type = cls[0] + 4 * (cls[1] + 4 * cls[2]);
      for (int pos = rem; pos < N1 * N1; pos += tpc)
        {
          T t[NB][N1];
#pragma unroll
          for (int bb = 0; bb < NB; ++bb)
            {
              const T *pl = buf + (slot * NB + bb) * CBS + pos * LS;
              T        in[N1];
#pragma unroll
              for (int xx = 0; xx < N1; ++xx) in[xx] = pl[xx];
              fd_apply_cls<T, N1, INTERIOR>(a.ST[0], cls[0], in, t[bb]);
            }
          const T *cm = a.modes + ((size_t)type * N1 * N1 + pos) * N1 * NB * NB;
          T        u[NB][N1];
#pragma unroll
          for (int mx = 0; mx < N1; ++mx)
            {
#pragma unroll
              for (int r = 0; r < NB; ++r)
                {
                  T s = T(0);
#pragma unroll
                  for (int cc = 0; cc < NB; ++cc) s += __ldg(cm + (mx * NB + r) * NB + cc) * t[cc][mx];
                  u[r][mx] = s;
                }
            }
#pragma unroll
          for (int bb = 0; bb < NB; ++bb)
            {
              T out[N1];
              fd_apply_cls<T, N1, INTERIOR>(a.S[0], cls[0], u[bb], out);
              T *pl = buf + (slot * NB + bb) * CBS + pos * LS;
#pragma unroll
              for (int xx = 0; xx < N1; ++xx) pl[xx] = out[xx];
            }
        }
    }
    __syncthreads();

    // ---------------- phase C: S_z, S_y, scatter-add (constrained rows skipped)
    {
      const T *pb = buf + cb * CBS + i;
#pragma unroll
      for (int jy = 0; jy < N1; ++jy)
        {
          T in[N1], t[N1];
#pragma unroll
          for (int q = 0; q < N1; ++q) in[q] = pb[(q * N1 + jy) * LS];
          fd_apply_cls<T, N1, INTERIOR>(a.S[2], cls[2], in, t);
#pragma unroll
          for (int k = 0; k < N1; ++k) x[k][jy] = t[k];
        }
    }
    const bool plane_con = (i == 0 && lo_con[0]) || (i == K && hi_con[0]);
    if (active && !plane_con)
      {
        T *d = a.dst + (size_t)b * a.N + base;
#pragma unroll
        for (int k = 0; k < N1; ++k)
          {
            T t[N1];
            fd_apply_cls<T, N1, INTERIOR>(a.S[1], cls[1], x[k], t);
            const bool kc = (k == 0 && lo_con[2]) || (k == K && hi_con[2]);
#pragma unroll
            for (int jy = 0; jy < N1; ++jy)
              {
                const bool cn = kc || (jy == 0 && lo_con[1]) || (jy == K && hi_con[1]);
                if (!cn) atomicAdd(d + jy * sy + k * sz, a.scale * t[jy]);
              }
          }
      }
  }
} // namespace stfem
