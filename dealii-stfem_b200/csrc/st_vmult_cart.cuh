// st_vmult, Cartesian variant (3D): fused application of
//     dst_j (+)= sum_s  Alpha(j,s) c_cell K src_s + Beta(j,s) M src_s
// on axis-aligned hexahedra.  Same arithmetic contract as the reference's cell kernel
// (include/operators.h:1112-1173, quadrature QGauss(k+1), FE_Q(k)) and block loop
// (SystemMatrix::vmult, include/operators.h:536-559), restructured for the B200 FP64 pipe:
//
//  * On a Cartesian cell  M_c = vol  Mh (x) Mh (x) Mh  and
//    K_c = vol (Kh/hx^2 (x) Mh (x) Mh + Mh (x) Kh/hy^2 (x) Mh + Mh (x) Mh (x) Kh/hz^2)
//    with the 1D reference matrices Mh = S^T W S, Kh = D^T W D of the Gauss rule, so the
//    interpolate -> q-point -> integrate chain collapses to 1D matrix applications on NODAL
//    values: 8 line sweeps per block instead of 12 + quadrature-point work.
//  * The temporal contraction is done on the nodal values while they are gathered
//    (v = sum_s Beta(j,s) u_s, w = sum_s Alpha(j,s) c u_s): everything after it is
//    per destination block, no exchange between time blocks is needed.
//  * One thread owns one y-z PLANE of a (cell, dst block): N1 x N1 values of v and w in registers,
//    lanes run along x, so global loads / reductions are coalesced (x is the fastest index).
//    The y and z sweeps run entirely in registers with the 1D matrices as compile-time indexed
//    constant-bank operands; ONE exchange through shared memory (fields P, Q) feeds the x sweep.
//        y:  s = Mh v + Ky w,  t = Mh w          z:  P = Mh s + Kz t,  Q = Mh t
//        x:  out = (vol Mh) P + (vol Kh/hx^2) Q
//    Per thread (Q4, nb = 2): ~1100 FMA against ~150 shared-memory accesses, so the kernel is
//    FP64-pipe bound instead of shared-memory bound (the generic kernel needs an LDS per FMA).
//  * The cell's space-time block is gathered once and scattered once (RED.ADD per node; nodes on
//    cell faces get one reduction per adjacent cell, resolved in L2).
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "../../include/stfem_b200.h"

namespace stfem
{
  // shared-memory layout of the exchange buffer: [cell-block][line][x], line stride LS, cell-block
  // stride CBS; `blocked`: x-sweep thread r takes lines N1*r..N1*r+N1-1, else r, r+N1, ...
  // (chosen by a bank-conflict search)
  template <int N1> struct ExchLayout;
  template <> struct ExchLayout<2> { static constexpr int LS = 3, CBS = 14; static constexpr bool blocked = false; };
  template <> struct ExchLayout<3> { static constexpr int LS = 3, CBS = 27; static constexpr bool blocked = true; };
  template <> struct ExchLayout<4> { static constexpr int LS = 5, CBS = 84; static constexpr bool blocked = false; };
  template <> struct ExchLayout<5> { static constexpr int LS = 5, CBS = 125; static constexpr bool blocked = true; };
  template <> struct ExchLayout<6> { static constexpr int LS = 7, CBS = 266; static constexpr bool blocked = false; };
  template <> struct ExchLayout<7> { static constexpr int LS = 7, CBS = 343; static constexpr bool blocked = true; };

  template <typename T, int N1>
  struct CartArgs
  {
    T         M[N1 * N1];                  // Mh (y and z sweeps)
    T         Ky[N1 * N1], Kz[N1 * N1];    // Kh / hy^2, Kh / hz^2
    T         Mx[N1 * N1], Kx[N1 * N1];    // vol * Mh, vol * Kh / hx^2
    int       n[3], np[3];
    int       box_lo[3], box_n[3]; // sub-box of cells this launch processes (whole mesh: lo = 0, n = mesh)
    // further boxes of the same launch (shell of cells along rank interfaces: up to 6 slabs in ONE launch instead of
    // six small ones with their own tails); box b covers the linear cell range [xbox_start[b], xbox_start[b+1])
    int       n_xbox;
    int       xbox_lo[6][3], xbox_n[6][3], xbox_start[7];
    long long n_cells;             // cells in the sub-box
    int       nb_src, nb_dst, cells_per_cta;
    unsigned  dirichlet;
    const T  *src[STFEM_MAX_BLOCKS];
    T        *dst[STFEM_MAX_BLOCKS];
    const T  *alpha, *beta; // device, nb_dst x nb_src row-major
    const T  *coeff_cell;   // optional per-cell Laplace coefficient
  };

  // MAXT / MINB: launch bounds (threads per CTA, resident CTAs per SM) = the register budget ptxas gets
  // EXPERIMENT (tuning only, wrong results): 1 = no scatter, 2 = no gather (constants), 3 = neither, 4 = no x sweep
  // PACKED (FP64, odd N1): P and Q of a node are exchanged as ONE 16-byte word.  With the dense layout
  // [cell-block][line][x] (N1^3 = N1 mod 8 words per cell-block) both the plane-thread stores and the x-line loads are
  // conflict-free 128-bit accesses; the x sweep keeps its N1 x N1 results in registers and, after a barrier, writes them
  // as 8-byte words at  cb*2*N1^3 + delta(cb) + line*N1 + x  with delta chosen such that the plane threads read them
  // back conflict-free too.  Same bytes through shared memory, ~30 % fewer wavefronts than two separate 8-byte fields.
  template <typename T> struct Pair2;
  template <> struct Pair2<double> { using type = double2; };
  template <> struct Pair2<float> { using type = float2; };

  template <int N1, typename T, int MAXT, int MINB, int EXPERIMENT = 0, bool PACKED = false>
  __global__ void __launch_bounds__(MAXT, MINB) st_vmult_cart_kernel(const __grid_constant__ CartArgs<T, N1> a)
  {
    using L           = ExchLayout<N1>;
    using T2          = typename Pair2<T>::type;
    constexpr int K   = N1 - 1;
    constexpr int LS  = L::LS;
    constexpr int CBS = L::CBS;
    constexpr int NC  = N1 * N1 * N1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T  *bufP = reinterpret_cast<T *>(smem_raw);
    T  *bufQ = bufP + (size_t)a.cells_per_cta * a.nb_dst * CBS;
    T2 *bufPQ = reinterpret_cast<T2 *>(smem_raw); // PACKED: [cb][line][x] pairs

    const int tid  = threadIdx.x;
    const int tpc  = a.nb_dst * N1; // threads per cell
    const int slot = tid / tpc;
    const int rem  = tid - slot * tpc;
    const int j    = rem / N1;     // destination block
    const int i    = rem - j * N1; // x index of this thread's plane (phases A, C); line group (phase B)
    const int cb   = tid / N1;     // cell-block slot in shared memory

    const long long cell   = (long long)blockIdx.x * a.cells_per_cta + slot;
    const bool      active = cell < a.n_cells;
    int             cx = 0, cy = 0, cz = 0;
    if (active)
      {
        unsigned c = (unsigned)cell; // < 2^31 cells (checked by the launcher): 32-bit divisions
        if (a.n_xbox > 0)
          {
            int bx = 0;
            while (bx + 1 < a.n_xbox && (int)c >= a.xbox_start[bx + 1]) ++bx;
            c -= (unsigned)a.xbox_start[bx];
            cx = a.xbox_lo[bx][0] + (int)(c % (unsigned)a.xbox_n[bx][0]);
            c /= (unsigned)a.xbox_n[bx][0];
            cy = a.xbox_lo[bx][1] + (int)(c % (unsigned)a.xbox_n[bx][1]);
            cz = a.xbox_lo[bx][2] + (int)(c / (unsigned)a.xbox_n[bx][1]);
          }
        else
          {
            cx = a.box_lo[0] + (int)(c % (unsigned)a.box_n[0]);
            c /= (unsigned)a.box_n[0];
            cy = a.box_lo[1] + (int)(c % (unsigned)a.box_n[1]);
            cz = a.box_lo[2] + (int)(c / (unsigned)a.box_n[1]);
          }
      }
    const unsigned dm  = a.dirichlet;
    const bool     xlo = (dm & 1u) && cx == 0, xhi = (dm & 2u) && cx == a.n[0] - 1;
    const bool     ylo = (dm & 4u) && cy == 0, yhi = (dm & 8u) && cy == a.n[1] - 1;
    const bool     zlo = (dm & 16u) && cz == 0, zhi = (dm & 32u) && cz == a.n[2] - 1;
    const bool     plane_constrained = (xlo && i == 0) || (xhi && i == K);
    const bool     any_yz            = ylo || yhi || zlo || zhi;
    const int      sy                = a.np[0];
    const int      sz                = a.np[0] * a.np[1];
    const long long base = (long long)(cx * K + i) + (long long)a.np[0] * ((long long)(cy * K) + (long long)a.np[1] * (cz * K));

    // ---------------- phase A: gather + temporal contraction (read_dof_values: constrained -> 0)
    T v[N1][N1], w[N1][N1]; // [z][y]
    if (active && !plane_constrained)
      {
        const T coef = a.coeff_cell ? a.coeff_cell[(long long)cx + (long long)a.n[0] * (cy + (long long)a.n[1] * cz)] : T(1);
        // the first source block initialises the accumulators (no zero fill), the others accumulate
        auto gather = [&](int s, auto first) {
          const T  be = a.beta[j * a.nb_src + s];
          const T  al = a.alpha[j * a.nb_src + s] * coef;
          const T *p  = a.src[s] + base;
          if (!any_yz)
            {
#pragma unroll
              for (int k = 0; k < N1; ++k)
#pragma unroll
                for (int jy = 0; jy < N1; ++jy)
                  {
                    const T u = (EXPERIMENT == 2 || EXPERIMENT == 3) ? T(jy + k) + be : p[jy * sy + k * sz];
                    v[k][jy]  = decltype(first)::value ? be * u : v[k][jy] + be * u;
                    w[k][jy]  = decltype(first)::value ? al * u : w[k][jy] + al * u;
                  }
            }
          else
            {
#pragma unroll
              for (int k = 0; k < N1; ++k)
#pragma unroll
                for (int jy = 0; jy < N1; ++jy)
                  {
                    const bool c = (ylo && jy == 0) || (yhi && jy == K) || (zlo && k == 0) || (zhi && k == K);
                    const T    u = c ? T(0) : p[jy * sy + k * sz];
                    v[k][jy]     = decltype(first)::value ? be * u : v[k][jy] + be * u;
                    w[k][jy]     = decltype(first)::value ? al * u : w[k][jy] + al * u;
                  }
            }
        };
        if (a.nb_src == 2 && !any_yz && EXPERIMENT == 0)
          {
            // two source blocks (the common case): all 2 N1^2 loads are issued back to back into the v / w registers,
            // the 2x2 contraction then happens in place -> one exposed load latency instead of two
            const T  be0 = a.beta[j * 2], be1 = a.beta[j * 2 + 1];
            const T  al0 = a.alpha[j * 2] * coef, al1 = a.alpha[j * 2 + 1] * coef;
            const T *p0 = a.src[0] + base, *p1 = a.src[1] + base;
#pragma unroll
            for (int k = 0; k < N1; ++k)
#pragma unroll
              for (int jy = 0; jy < N1; ++jy)
                {
                  v[k][jy] = p0[jy * sy + k * sz];
                  w[k][jy] = p1[jy * sy + k * sz];
                }
#pragma unroll
            for (int k = 0; k < N1; ++k)
#pragma unroll
              for (int jy = 0; jy < N1; ++jy)
                {
                  const T u0 = v[k][jy], u1 = w[k][jy];
                  v[k][jy]   = be0 * u0 + be1 * u1;
                  w[k][jy]   = al0 * u0 + al1 * u1;
                }
          }
        else
          {
            gather(0, std::true_type());
            for (int s = 1; s < a.nb_src; ++s) gather(s, std::false_type());
          }
      }
    else
      {
#pragma unroll
        for (int k = 0; k < N1; ++k)
#pragma unroll
          for (int jy = 0; jy < N1; ++jy) v[k][jy] = w[k][jy] = T(0);
      }

    // ---------------- y sweep (registers):  s = Mh v + Ky w,  t = Mh w
#pragma unroll
    for (int k = 0; k < N1; ++k)
      {
        T s[N1], t[N1];
#pragma unroll
        for (int q = 0; q < N1; ++q)
          {
            T ss = T(0), tt = T(0);
#pragma unroll
            for (int jy = 0; jy < N1; ++jy)
              {
                ss += a.M[q * N1 + jy] * v[k][jy];
                ss += a.Ky[q * N1 + jy] * w[k][jy];
                tt += a.M[q * N1 + jy] * w[k][jy];
              }
            s[q] = ss;
            t[q] = tt;
          }
#pragma unroll
        for (int q = 0; q < N1; ++q)
          {
            v[k][q] = s[q];
            w[k][q] = t[q];
          }
      }

    // ---------------- z sweep (registers) + publish:  P = Mh s + Kz t,  Q = Mh t
    {
      T *pP = bufP + cb * CBS + i;
      T *pQ = bufQ + cb * CBS + i;
#pragma unroll
      for (int jy = 0; jy < N1; ++jy)
#pragma unroll
        for (int q = 0; q < N1; ++q)
          {
            T pp = T(0), qq = T(0);
#pragma unroll
            for (int k = 0; k < N1; ++k)
              {
                pp += a.M[q * N1 + k] * v[k][jy];
                pp += a.Kz[q * N1 + k] * w[k][jy];
                qq += a.M[q * N1 + k] * w[k][jy];
              }
            if (PACKED)
              {
                T2 pr;
                pr.x = pp;
                pr.y = qq;
                bufPQ[cb * NC + (q * N1 + jy) * N1 + i] = pr;
              }
            else
              {
                pP[(q * N1 + jy) * LS] = pp;
                pQ[(q * N1 + jy) * LS] = qq;
              }
          }
    }
    __syncthreads();

    // out words of the packed variant: conflict-free for the plane threads (see above)
    const int out_base = cb * 2 * NC + (((N1 - 2 * NC) * cb) & 15);
    // ---------------- phase B: x sweep on N1 lines of this cell-block, result overwrites P
    if (PACKED)
      {
        T o[N1][N1];
#pragma unroll
        for (int m = 0; m < N1; ++m)
          {
            const T2 *pl = bufPQ + cb * NC + (N1 * i + m) * N1;
            T         P[N1], Q[N1];
#pragma unroll
            for (int x = 0; x < N1; ++x)
              {
                const T2 pr = pl[x];
                P[x]        = pr.x;
                Q[x]        = pr.y;
              }
#pragma unroll
            for (int q = 0; q < N1; ++q)
              {
                T oo = T(0);
#pragma unroll
                for (int x = 0; x < N1; ++x)
                  {
                    oo += a.Mx[q * N1 + x] * P[x];
                    oo += a.Kx[q * N1 + x] * Q[x];
                  }
                o[m][q] = oo;
              }
          }
        __syncthreads(); // every pair has been read: the storage can be reused for the results
#pragma unroll
        for (int m = 0; m < N1; ++m)
#pragma unroll
          for (int q = 0; q < N1; ++q) bufP[out_base + (N1 * i + m) * N1 + q] = o[m][q];
      }
    else if (EXPERIMENT != 4)
    {
#pragma unroll
      for (int m = 0; m < N1; ++m)
        {
          const int line = L::blocked ? N1 * i + m : i + N1 * m;
          T        *pP   = bufP + cb * CBS + line * LS;
          const T  *pQ   = bufQ + cb * CBS + line * LS;
          T         P[N1], Q[N1];
#pragma unroll
          for (int x = 0; x < N1; ++x)
            {
              P[x] = pP[x];
              Q[x] = pQ[x];
            }
#pragma unroll
          for (int q = 0; q < N1; ++q)
            {
              T o = T(0);
#pragma unroll
              for (int x = 0; x < N1; ++x)
                {
                  o += a.Mx[q * N1 + x] * P[x];
                  o += a.Kx[q * N1 + x] * Q[x];
                }
              pP[q] = o;
            }
        }
    }
    __syncthreads();

    // ---------------- phase C: scatter-add (distribute_local_to_global: constrained rows skipped)
    if (active && !plane_constrained)
      {
        T       *d  = a.dst[j] + base;
        const T *pP = PACKED ? bufP + out_base + i : bufP + cb * CBS + i;
        if (EXPERIMENT == 1 || EXPERIMENT == 3)
          {
            T acc = T(0);
#pragma unroll
            for (int k = 0; k < N1; ++k)
#pragma unroll
              for (int jy = 0; jy < N1; ++jy) acc += pP[(k * N1 + jy) * LS];
            if (acc == T(12345.678)) d[0] = acc;
          }
        else if (!any_yz)
          {
#pragma unroll
            for (int k = 0; k < N1; ++k)
#pragma unroll
              for (int jy = 0; jy < N1; ++jy) atomicAdd(d + jy * sy + k * sz, pP[(k * N1 + jy) * LS]);
          }
        else
          {
#pragma unroll
            for (int k = 0; k < N1; ++k)
#pragma unroll
              for (int jy = 0; jy < N1; ++jy)
                {
                  const bool c = (ylo && jy == 0) || (yhi && jy == K) || (zlo && k == 0) || (zhi && k == K);
                  if (!c) atomicAdd(d + jy * sy + k * sz, pP[(k * N1 + jy) * LS]);
                }
          }
      }
  }
  // ------------------------------------------------------------------------------------------------------------
  // Software-pipelined persistent variant (NBS = number of source blocks, compile time, <= 2): every CTA loops over
  // batches of cells; the gather of batch n+1 is issued (raw values into the registers the y/z sweeps have just
  // released) BEFORE the x sweep and the scatter of batch n, so the global-load latency hides behind shared-memory
  // and RED work instead of stalling the FP64 pipe.
  template <typename T>
  struct CartCell
  {
    long long base, cell_lin;
    bool      active, plane_con, any_yz, ylo, yhi, zlo, zhi;
  };

  template <int N1, typename T>
  __device__ __forceinline__ void cart_decode(const CartArgs<T, N1> &a, long long batch, int slot, int i, CartCell<T> &c)
  {
    constexpr int   K    = N1 - 1;
    const long long cell = batch * a.cells_per_cta + slot;
    c.active             = cell < a.n_cells;
    int cx = 0, cy = 0, cz = 0;
    if (c.active)
      {
        unsigned r = (unsigned)cell;
        cx         = a.box_lo[0] + (int)(r % (unsigned)a.box_n[0]);
        r /= (unsigned)a.box_n[0];
        cy = a.box_lo[1] + (int)(r % (unsigned)a.box_n[1]);
        cz = a.box_lo[2] + (int)(r / (unsigned)a.box_n[1]);
      }
    const unsigned dm  = a.dirichlet;
    const bool     xlo = (dm & 1u) && cx == 0, xhi = (dm & 2u) && cx == a.n[0] - 1;
    c.ylo = (dm & 4u) && cy == 0;
    c.yhi = (dm & 8u) && cy == a.n[1] - 1;
    c.zlo = (dm & 16u) && cz == 0;
    c.zhi = (dm & 32u) && cz == a.n[2] - 1;
    c.plane_con = (xlo && i == 0) || (xhi && i == K);
    c.any_yz    = c.ylo || c.yhi || c.zlo || c.zhi;
    c.base      = (long long)(cx * K + i) + (long long)a.np[0] * ((long long)(cy * K) + (long long)a.np[1] * (cz * K));
    c.cell_lin  = (long long)cx + (long long)a.n[0] * (cy + (long long)a.n[1] * cz);
  }

  // raw values of source block 0 into v, of source block 1 (NBS == 2) into w
  template <int N1, typename T, int NBS>
  __device__ __forceinline__ void cart_gather(const CartArgs<T, N1> &a, const CartCell<T> &c, T (&v)[N1][N1], T (&w)[N1][N1])
  {
    constexpr int K  = N1 - 1;
    const int     sy = a.np[0], sz = a.np[0] * a.np[1];
    if (c.active && !c.plane_con)
      {
#pragma unroll
        for (int s = 0; s < NBS; ++s)
          {
            const T *p = a.src[s] + c.base;
            if (!c.any_yz)
              {
#pragma unroll
                for (int k = 0; k < N1; ++k)
#pragma unroll
                  for (int jy = 0; jy < N1; ++jy) (s == 0 ? v : w)[k][jy] = p[jy * sy + k * sz];
              }
            else
              {
#pragma unroll
                for (int k = 0; k < N1; ++k)
#pragma unroll
                  for (int jy = 0; jy < N1; ++jy)
                    {
                      const bool cn = (c.ylo && jy == 0) || (c.yhi && jy == K) || (c.zlo && k == 0) || (c.zhi && k == K);
                      (s == 0 ? v : w)[k][jy] = cn ? T(0) : p[jy * sy + k * sz];
                    }
              }
          }
      }
    else
      {
#pragma unroll
        for (int k = 0; k < N1; ++k)
#pragma unroll
          for (int jy = 0; jy < N1; ++jy) v[k][jy] = w[k][jy] = T(0);
      }
  }

  // EARLY: prefetch before the x sweep (longest overlap, needs the registers for it: fine in FP32), else after it
  template <int N1, typename T, int NBS, int MAXREG, bool EARLY>
  __global__ void __maxnreg__(MAXREG) st_vmult_cart_pipe_kernel(const __grid_constant__ CartArgs<T, N1> a)
  {
    using L           = ExchLayout<N1>;
    constexpr int K   = N1 - 1;
    constexpr int LS  = L::LS;
    constexpr int CBS = L::CBS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *bufP = reinterpret_cast<T *>(smem_raw);
    T *bufQ = bufP + (size_t)a.cells_per_cta * a.nb_dst * CBS;

    const int tid  = threadIdx.x;
    const int tpc  = a.nb_dst * N1;
    const int slot = tid / tpc;
    const int rem  = tid - slot * tpc;
    const int j    = rem / N1;
    const int i    = rem - j * N1;
    const int cb   = tid / N1;
    const int sy   = a.np[0];
    const int sz   = a.np[0] * a.np[1];

    T be[NBS], al0[NBS];
#pragma unroll
    for (int s = 0; s < NBS; ++s)
      {
        be[s]  = a.beta[j * NBS + s];
        al0[s] = a.alpha[j * NBS + s];
      }
    const long long n_batches = (a.n_cells + a.cells_per_cta - 1) / a.cells_per_cta;
    long long       batch     = blockIdx.x;
    static_assert(NBS == 1 || NBS == 2, "pipelined kernel: one or two source blocks");
    CartCell<T>     c;
    T               v[N1][N1], w[N1][N1]; // [z][y]; hold the raw gathered values between iterations
    cart_decode<N1, T>(a, batch, slot, i, c);
    cart_gather<N1, T, NBS>(a, c, v, w);

    for (; batch < n_batches; batch += gridDim.x)
      {
        // ---------------- temporal contraction of the gathered values
        {
          const T coef = (a.coeff_cell && c.active) ? a.coeff_cell[c.cell_lin] : T(1);
#pragma unroll
          for (int k = 0; k < N1; ++k)
#pragma unroll
            for (int jy = 0; jy < N1; ++jy)
              {
                const T u0 = v[k][jy], u1 = NBS == 2 ? w[k][jy] : T(0);
                T       vv = be[0] * u0, ww = (al0[0] * coef) * u0;
                if (NBS == 2)
                  {
                    vv += be[1] * u1;
                    ww += (al0[1] * coef) * u1;
                  }
                v[k][jy] = vv;
                w[k][jy] = ww;
              }
        }
        // ---------------- y sweep
#pragma unroll
        for (int k = 0; k < N1; ++k)
          {
            T s_[N1], t_[N1];
#pragma unroll
            for (int q = 0; q < N1; ++q)
              {
                T ss = T(0), tt = T(0);
#pragma unroll
                for (int jy = 0; jy < N1; ++jy)
                  {
                    ss += a.M[q * N1 + jy] * v[k][jy];
                    ss += a.Ky[q * N1 + jy] * w[k][jy];
                    tt += a.M[q * N1 + jy] * w[k][jy];
                  }
                s_[q] = ss;
                t_[q] = tt;
              }
#pragma unroll
            for (int q = 0; q < N1; ++q)
              {
                v[k][q] = s_[q];
                w[k][q] = t_[q];
              }
          }
        // ---------------- z sweep + publish
        {
          T *pP = bufP + cb * CBS + i;
          T *pQ = bufQ + cb * CBS + i;
#pragma unroll
          for (int jy = 0; jy < N1; ++jy)
#pragma unroll
            for (int q = 0; q < N1; ++q)
              {
                T pp = T(0), qq = T(0);
#pragma unroll
                for (int k = 0; k < N1; ++k)
                  {
                    pp += a.M[q * N1 + k] * v[k][jy];
                    pp += a.Kz[q * N1 + k] * w[k][jy];
                    qq += a.M[q * N1 + k] * w[k][jy];
                  }
                pP[(q * N1 + jy) * LS] = pp;
                pQ[(q * N1 + jy) * LS] = qq;
              }
        }
        __syncthreads();
        // ---------------- prefetch: gather of the next batch into the registers the sweeps released
        const CartCell<T> cur = c;
        if (EARLY && batch + gridDim.x < n_batches)
          {
            cart_decode<N1, T>(a, batch + gridDim.x, slot, i, c);
            cart_gather<N1, T, NBS>(a, c, v, w);
          }
        // ---------------- x sweep
#pragma unroll
        for (int m = 0; m < N1; ++m)
          {
            const int line = L::blocked ? N1 * i + m : i + N1 * m;
            T        *pP   = bufP + cb * CBS + line * LS;
            const T  *pQ   = bufQ + cb * CBS + line * LS;
            T         P[N1], Q[N1];
#pragma unroll
            for (int x = 0; x < N1; ++x)
              {
                P[x] = pP[x];
                Q[x] = pQ[x];
              }
#pragma unroll
            for (int q = 0; q < N1; ++q)
              {
                T o = T(0);
#pragma unroll
                for (int x = 0; x < N1; ++x)
                  {
                    o += a.Mx[q * N1 + x] * P[x];
                    o += a.Kx[q * N1 + x] * Q[x];
                  }
                pP[q] = o;
              }
          }
        if (!EARLY && batch + gridDim.x < n_batches)
          {
            cart_decode<N1, T>(a, batch + gridDim.x, slot, i, c);
            cart_gather<N1, T, NBS>(a, c, v, w);
          }
        __syncthreads();
        // ---------------- scatter-add
        if (cur.active && !cur.plane_con)
          {
            T       *d  = a.dst[j] + cur.base;
            const T *pP = bufP + cb * CBS + i;
            if (!cur.any_yz)
              {
#pragma unroll
                for (int k = 0; k < N1; ++k)
#pragma unroll
                  for (int jy = 0; jy < N1; ++jy) atomicAdd(d + jy * sy + k * sz, pP[(k * N1 + jy) * LS]);
              }
            else
              {
#pragma unroll
                for (int k = 0; k < N1; ++k)
#pragma unroll
                  for (int jy = 0; jy < N1; ++jy)
                    {
                      const bool cn = (cur.ylo && jy == 0) || (cur.yhi && jy == K) || (cur.zlo && k == 0) || (cur.zhi && k == K);
                      if (!cn) atomicAdd(d + jy * sy + k * sz, pP[(k * N1 + jy) * LS]);
                    }
              }
          }
        __syncthreads(); // bufP is rewritten by the next batch
      }
  }
} // namespace stfem

// ----------------------------------------------------------------------------------------------------------------
// TMA variant of the Cartesian kernel (sm_90+/sm_100a: cp.async.bulk + mbarrier).  Persistent CTAs; a batch is
// CPC x-adjacent cells of one mesh row, whose source values are CPC*K+1 consecutive numbers per (block, y, z): the
// 25*nb_src rows of a batch are fetched by bulk asynchronous copies (global -> shared, completion counted on an
// mbarrier) issued by warp 0 as soon as the previous batch has been consumed, i.e. they land while the x sweep and
// the scatter of the previous batch run.  No register staging and no LSU gather traffic: phase A reads the rows
// from shared memory (conflict-free, lanes = consecutive x).  Rows are only 8-byte aligned in the block vectors
// (odd numbers of DoFs per line); every copy starts at the enclosing 16-byte boundary and moves one extra element,
// the parity of each row is recomputed when it is read.  The last element of the last row may lie one number past
// the end of a block vector: allocations are padded (stfem_dev_alloc).
namespace stfem
{
  __device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
  __device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  }
  __device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
  {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  }
  __device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
  {
    unsigned ok;
    do
      {
        asm volatile("{\n"
                     ".reg .pred p;\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     "selp.u32 %0, 1, 0, p;\n"
                     "}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
      }
    while (!ok);
  }
  __device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar)
  {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
  }

  // CPC cells per batch (compile time: row length), requires n_x % CPC == 0 and a launch over whole mesh rows
  template <int N1, typename T, int CPC, int MAXT, int MINB>
  __global__ void __launch_bounds__(MAXT, MINB) st_vmult_cart_tma_kernel(const __grid_constant__ CartArgs<T, N1> a)
  {
    using L              = ExchLayout<N1>;
    constexpr int K      = N1 - 1;
    constexpr int LS     = L::LS;
    constexpr int CBS    = L::CBS;
    constexpr int EPV    = 16 / (int)sizeof(T);                    // elements per 16 bytes
    constexpr int ROWLEN = CPC * K + 1;                            // numbers per row
    constexpr int COPY   = ((ROWLEN + EPV - 1 + EPV - 1) / EPV) * EPV; // copied numbers: worst misalignment included
    constexpr int RSTR   = COPY;                                   // row stride in the tile (multiple of 16 bytes)
    extern __shared__ __align__(128) unsigned char smem_tma[];
    unsigned long long *bar  = reinterpret_cast<unsigned long long *>(smem_tma);
    T                  *tile = reinterpret_cast<T *>(smem_tma + 16);
    const int           n_rows = N1 * N1 * a.nb_src;
    T                  *bufP = tile + (size_t)n_rows * RSTR;
    T                  *bufQ = bufP + (size_t)CPC * a.nb_dst * CBS;

    const int tid  = threadIdx.x;
    const int tpc  = a.nb_dst * N1;
    const int slot = tid / tpc;
    const int rem  = tid - slot * tpc;
    const int j    = rem / N1;
    const int i    = rem - j * N1;
    const int cb   = tid / N1;
    const bool worker = slot < CPC; // threads beyond CPC*tpc (warp padding) only take part in barriers
    const int sy   = a.np[0];
    const int sz   = a.np[0] * a.np[1];
    const int bpr  = a.n[0] / CPC;                                 // batches per mesh row
    const long long n_batches = (long long)bpr * a.n[1] * a.n[2];

    // row r = (s, k, jy) of batch b: source address, rounded down to 16 bytes
    auto issue = [&](long long b) {
      const int bx = (int)(b % bpr);
      long long r2 = b / bpr;
      const int cy = (int)(r2 % a.n[1]), cz = (int)(r2 / a.n[1]);
      const long long e0 = (long long)bx * CPC * K + (long long)sy * (cy * K) + (long long)sz * (cz * K);
      if (tid == 0) mbar_expect_tx(bar, (unsigned)(n_rows * COPY * sizeof(T)));
      // the copies are warp-uniform instructions (one per lane, serialised): spread the rows over all warps
      const int n_warps = blockDim.x >> 5;
      for (int r = (tid >> 5) + n_warps * (tid & 31); r < n_rows; r += n_warps * 32)
        {
          const int s = r / (N1 * N1), kj = r - s * (N1 * N1);
          const int k = kj / N1, jy = kj - k * N1;
          const unsigned long long addr = (unsigned long long)(a.src[s] + e0 + (long long)jy * sy + (long long)k * sz);
          bulk_g2s(tile + (size_t)r * RSTR, (const void *)(addr & ~15ull), COPY * sizeof(T), bar);
        }
    };

    if (tid == 0)
      {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
    __syncthreads();
    long long batch = blockIdx.x;
    if (batch < n_batches) issue(batch);
    unsigned phase = 0;

    for (; batch < n_batches; batch += gridDim.x)
      {
        const int bx = (int)(batch % bpr);
        long long r2 = batch / bpr;
        const int cy = (int)(r2 % a.n[1]), cz = (int)(r2 / a.n[1]);
        const int cx = bx * CPC + slot;
        const unsigned dm  = a.dirichlet;
        const bool     xlo = (dm & 1u) && cx == 0, xhi = (dm & 2u) && cx == a.n[0] - 1;
        const bool     ylo = (dm & 4u) && cy == 0, yhi = (dm & 8u) && cy == a.n[1] - 1;
        const bool     zlo = (dm & 16u) && cz == 0, zhi = (dm & 32u) && cz == a.n[2] - 1;
        const bool     plane_constrained = (xlo && i == 0) || (xhi && i == K);
        const bool     any_yz            = ylo || yhi || zlo || zhi;
        const long long e0   = (long long)bx * CPC * K + (long long)sy * (cy * K) + (long long)sz * (cz * K);
        const long long base = e0 + slot * K + i;

        // ---------------- phase A: contraction straight from the tile
        T v[N1][N1], w[N1][N1];
#pragma unroll
        for (int k = 0; k < N1; ++k)
#pragma unroll
          for (int jy = 0; jy < N1; ++jy) v[k][jy] = w[k][jy] = T(0);
        mbar_wait(bar, phase);
        phase ^= 1u;
        if (worker && !plane_constrained)
          {
            const T coef = a.coeff_cell ? a.coeff_cell[(long long)cx + (long long)a.n[0] * (cy + (long long)a.n[1] * cz)] : T(1);
            for (int s = 0; s < a.nb_src; ++s)
              {
                const T be = a.beta[j * a.nb_src + s];
                const T al = a.alpha[j * a.nb_src + s] * coef;
                // misalignment (in numbers) of the row (k = 0, jy = 0) and its change per row
                const unsigned long long a0 = (unsigned long long)(a.src[s] + e0);
                const int                m0 = (int)((a0 & 15ull) / sizeof(T));
                const T                 *tp = tile + (size_t)s * (N1 * N1) * RSTR + slot * K + i;
#pragma unroll
                for (int k = 0; k < N1; ++k)
#pragma unroll
                  for (int jy = 0; jy < N1; ++jy)
                    {
                      const int  mis = (m0 + jy * sy + k * sz) & (EPV - 1);
                      const bool c   = any_yz && ((ylo && jy == 0) || (yhi && jy == K) || (zlo && k == 0) || (zhi && k == K));
                      const T    u   = c ? T(0) : tp[(k * N1 + jy) * RSTR + mis];
                      v[k][jy] += be * u;
                      w[k][jy] += al * u;
                    }
              }
          }
        // ---------------- y sweep
#pragma unroll
        for (int k = 0; k < N1; ++k)
          {
            T s_[N1], t_[N1];
#pragma unroll
            for (int q = 0; q < N1; ++q)
              {
                T ss = T(0), tt = T(0);
#pragma unroll
                for (int jy = 0; jy < N1; ++jy)
                  {
                    ss += a.M[q * N1 + jy] * v[k][jy];
                    ss += a.Ky[q * N1 + jy] * w[k][jy];
                    tt += a.M[q * N1 + jy] * w[k][jy];
                  }
                s_[q] = ss;
                t_[q] = tt;
              }
#pragma unroll
            for (int q = 0; q < N1; ++q)
              {
                v[k][q] = s_[q];
                w[k][q] = t_[q];
              }
          }
        // ---------------- z sweep + publish
        if (worker)
          {
            T *pP = bufP + cb * CBS + i;
            T *pQ = bufQ + cb * CBS + i;
#pragma unroll
            for (int jy = 0; jy < N1; ++jy)
#pragma unroll
              for (int q = 0; q < N1; ++q)
                {
                  T pp = T(0), qq = T(0);
#pragma unroll
                  for (int k = 0; k < N1; ++k)
                    {
                      pp += a.M[q * N1 + k] * v[k][jy];
                      pp += a.Kz[q * N1 + k] * w[k][jy];
                      qq += a.M[q * N1 + k] * w[k][jy];
                    }
                  pP[(q * N1 + jy) * LS] = pp;
                  pQ[(q * N1 + jy) * LS] = qq;
                }
          }
        __syncthreads(); // tile consumed by everybody, P/Q complete
        if (batch + gridDim.x < n_batches) issue(batch + gridDim.x);
        // ---------------- x sweep
        if (worker)
          {
#pragma unroll
            for (int m = 0; m < N1; ++m)
              {
                const int line = L::blocked ? N1 * i + m : i + N1 * m;
                T        *pP   = bufP + cb * CBS + line * LS;
                const T  *pQ   = bufQ + cb * CBS + line * LS;
                T         P[N1], Q[N1];
#pragma unroll
                for (int x = 0; x < N1; ++x)
                  {
                    P[x] = pP[x];
                    Q[x] = pQ[x];
                  }
#pragma unroll
                for (int q = 0; q < N1; ++q)
                  {
                    T o = T(0);
#pragma unroll
                    for (int x = 0; x < N1; ++x)
                      {
                        o += a.Mx[q * N1 + x] * P[x];
                        o += a.Kx[q * N1 + x] * Q[x];
                      }
                    pP[q] = o;
                  }
              }
          }
        __syncthreads();
        // ---------------- scatter-add
        if (worker && !plane_constrained)
          {
            T       *d  = a.dst[j] + base;
            const T *pP = bufP + cb * CBS + i;
#pragma unroll
            for (int k = 0; k < N1; ++k)
#pragma unroll
              for (int jy = 0; jy < N1; ++jy)
                {
                  const bool cn = any_yz && ((ylo && jy == 0) || (yhi && jy == K) || (zlo && k == 0) || (zhi && k == K));
                  if (!cn) atomicAdd(d + jy * sy + k * sz, pP[(k * N1 + jy) * LS]);
                }
          }
        __syncthreads(); // P/Q are rewritten by the next batch
      }
  }
} // namespace stfem
