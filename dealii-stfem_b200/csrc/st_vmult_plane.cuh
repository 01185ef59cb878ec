// st_vmult, general-geometry variant (3D): fused application of
//     dst_j (+)= sum_s  Alpha(j,s) K src_s + Beta(j,s) M src_s
// on arbitrary (MappingQ1) hexahedra with per-cell / per-quadrature-point Laplace coefficients, using the
// precomputed metric  G = c J^-1 J^-T |J| w  (6 numbers) and |J| w  per quadrature point.
// Arithmetic contract = the reference's cell kernel (include/operators.h:1112-1173: evaluate values + gradients,
// q-loop submit_value / submit_gradient, integrate) inside SystemMatrix::vmult's block loop (operators.h:536-559).
//
// Same thread organisation as st_vmult_cart.cuh (a thread owns a y-z plane of a (cell, destination block), lanes
// run along x, 1D matrices are compile-time indexed constant-bank operands), with the collocation structure of the
// Gauss points exploited so that only TWO shared-memory exchanges are needed:
//   A  (plane threads)  gather + temporal contraction at the nodes: v = sum Beta u (mass field), w = sum Alpha c u
//                       (stiffness field); interpolate both to the Gauss points in y and z; y/z collocation
//                       derivatives of w  ->  fields  v, w, gy, gz  to shared memory
//   B  (x-line threads) interpolate the four fields in x, gx = Dc w; quadrature-point operation with the metric;
//                       transposed x operations  ->  fields  F0 = S^T (|J|w v + Dc^T g'x),  F1 = S^T g'y,  F2 = S^T g'z
//   C  (plane threads)  out = Sy^T Sz^T (F0 + Dc_y^T F1 + Dc_z^T F2);  RED.ADD scatter
// The metric is stored in the order phase B reads it ([cell][qy][qx][pair][qz][2]) so that the lanes of a warp
// read contiguous 16-byte pieces.
#pragma once
#include <cuda_runtime.h>

#include "st_vmult_cart.cuh"

namespace stfem
{
  template <typename T, int N1>
  struct PlaneArgs
  {
    T         S[N1 * N1];  // S[q*N1+a]  GLL basis a at Gauss point q
    T         Dc[N1 * N1]; // Dc[q*N1+p] derivative of the Lagrange basis on the Gauss points (collocation)
    int       n[3], np[3];
    int       box_lo[3], box_n[3];
    long long n_cells;
    int       nb_src, nb_dst, cells_per_cta;
    unsigned  dirichlet;
    const T  *src[STFEM_MAX_BLOCKS];
    T        *dst[STFEM_MAX_BLOCKS];
    const T  *alpha, *beta;
    const T  *coeff_cell;
    const T  *metric; // [cell][qy][qx][4][qz][2]: (Gxx Gxy) (Gxz Gyy) (Gyz Gzz) (JxW 0)
    // on-the-fly geometry (OTF): the metric is computed in phase B from the cell's 8 vertices instead of being streamed
    const double *vertices; // (n0+1)(n1+1)(n2+1) points, lexicographic, xyz interleaved
    T             xq[N1], wq[N1]; // Gauss points / weights on [0,1]
  };

  template <typename T> struct Vec2T;
  template <> struct Vec2T<double> { using type = double2; };
  template <> struct Vec2T<float> { using type = float2; };

  // out[q] = sum_a Mat[q*N1+a] in[a]  (TR: out[a] = sum_q Mat[q*N1+a] in[q])
  template <typename T, int N1, bool TR>
  __device__ __forceinline__ void pl_apply(const T (&Mat)[N1 * N1], const T (&in)[N1], T (&out)[N1])
  {
#pragma unroll
    for (int q = 0; q < N1; ++q)
      {
        T s = T(0);
#pragma unroll
        for (int a = 0; a < N1; ++a) s += (TR ? Mat[a * N1 + q] : Mat[q * N1 + a]) * in[a];
        out[q] = s;
      }
  }
  template <typename T, int N1, bool TR>
  __device__ __forceinline__ void pl_apply_add(const T (&Mat)[N1 * N1], const T (&in)[N1], T (&out)[N1])
  {
#pragma unroll
    for (int q = 0; q < N1; ++q)
      {
        T s = out[q];
#pragma unroll
        for (int a = 0; a < N1; ++a) s += (TR ? Mat[a * N1 + q] : Mat[q * N1 + a]) * in[a];
        out[q] = s;
      }
  }

  // OTF: MappingQ1 geometry on the fly (reference include/operators.h:973-1004 builds MatrixFree with
  // MappingQ1 and lets it store J^-1 and JxW per quadrature point; here nothing per quadrature point is stored):
  // J(xi) = [c0(eta,zeta) | c1(xi,zeta) | c2(xi,eta)], each column a bilinear blend of the four cell edges of that
  // direction.  Per cell the 3 (k+1)^2 column samples go to shared memory once; phase B then forms, per quadrature point,
  // the cross products (= rows of det J * J^-1), det J and  G = J^-1 J^-T det J w = (cross_r . cross_s) w / det J.
  template <int N1, typename T, int MAXT, int MINB, bool OTF = false>
  __global__ void __launch_bounds__(MAXT, MINB) st_vmult_plane_kernel(const __grid_constant__ PlaneArgs<T, N1> a)
  {
    using L           = ExchLayout<N1>;
    using V2          = typename Vec2T<T>::type;
    constexpr int K   = N1 - 1;
    constexpr int LS  = L::LS;
    constexpr int CBS = L::CBS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T        *buf = reinterpret_cast<T *>(smem_raw);
    const int FS  = a.cells_per_cta * a.nb_dst * CBS; // field stride
    constexpr int JT = 3 * N1 * N1 * 3;               // OTF: column samples of one cell, [column][i][j][component]
    T        *jtab   = buf + 4 * FS;                  // [cell slot][JT]
    T        *jedges = jtab + a.cells_per_cta * JT;   // [cell slot][12 edges][3]

    const int tid  = threadIdx.x;
    const int tpc  = a.nb_dst * N1;
    const int slot = tid / tpc;
    const int rem  = tid - slot * tpc;
    const int j    = rem / N1;
    const int i    = rem - j * N1;
    const int cb   = tid / N1;

    const long long cell_in_box = (long long)blockIdx.x * a.cells_per_cta + slot;
    const bool      active      = cell_in_box < a.n_cells;
    int             cx = 0, cy = 0, cz = 0;
    if (active)
      {
        unsigned c = (unsigned)cell_in_box; // < 2^31 cells (checked by the launcher): 32-bit divisions
        cx          = a.box_lo[0] + (int)(c % (unsigned)a.box_n[0]);
        c /= (unsigned)a.box_n[0];
        cy = a.box_lo[1] + (int)(c % (unsigned)a.box_n[1]);
        cz = a.box_lo[2] + (int)(c / (unsigned)a.box_n[1]);
      }
    const long long cell = (long long)cx + (long long)a.n[0] * (cy + (long long)a.n[1] * cz);
    const unsigned  dm   = a.dirichlet;
    const bool      xlo = (dm & 1u) && cx == 0, xhi = (dm & 2u) && cx == a.n[0] - 1;
    const bool      ylo = (dm & 4u) && cy == 0, yhi = (dm & 8u) && cy == a.n[1] - 1;
    const bool      zlo = (dm & 16u) && cz == 0, zhi = (dm & 32u) && cz == a.n[2] - 1;
    const bool      plane_constrained = (xlo && i == 0) || (xhi && i == K);
    const bool      any_yz            = ylo || yhi || zlo || zhi;
    const int       sy                = a.np[0];
    const int       sz                = a.np[0] * a.np[1];
    const long long base = (long long)(cx * K + i) + (long long)a.np[0] * ((long long)(cy * K) + (long long)a.np[1] * (cz * K));

    if (OTF)
      {
        // edge vectors of the cell: edge e = 4 d + (b1 + 2 b2) runs in direction d between the vertices whose other two
        // coordinates (in increasing direction order) are b1, b2
        const int tpc_ = a.nb_dst * N1, lt = tid - slot * tpc_;
        if (active)
          for (int e = lt; e < 12; e += tpc_)
            {
              const int d = e >> 2, b1 = e & 1, b2 = (e >> 1) & 1;
              int       v0[3];
              v0[d]                    = 0;
              v0[d == 0 ? 1 : 0]       = b1;
              v0[d == 2 ? 1 : 2]       = b2;
              const long long vid0 = (long long)(cx + v0[0]) + (long long)(a.n[0] + 1) * ((cy + v0[1]) + (long long)(a.n[1] + 1) * (cz + v0[2]));
              const long long str  = d == 0 ? 1 : (d == 1 ? (long long)(a.n[0] + 1) : (long long)(a.n[0] + 1) * (a.n[1] + 1));
#pragma unroll
              for (int comp = 0; comp < 3; ++comp)
                jedges[(slot * 12 + e) * 3 + comp] = (T)(__ldg(a.vertices + (vid0 + str) * 3 + comp) - __ldg(a.vertices + vid0 * 3 + comp));
            }
        __syncthreads();
        // column d at (i, j): blend over (b1, b2) with weights L(b1, xq[i]) L(b2, xq[j]); (i, j) = the Gauss indices of the two
        // other directions in increasing order
        if (active)
          for (int row = lt; row < 3 * N1; row += tpc_)
            {
              const int d = row / N1, i = row - d * N1;
              const T   xi = a.xq[i];
              T         f0[3], f1[3];
#pragma unroll
              for (int comp = 0; comp < 3; ++comp)
                {
                  const T *e = jedges + (slot * 12 + 4 * d) * 3 + comp;
                  f0[comp]   = e[0] + xi * (e[3] - e[0]);  // b2 = 0: edges (b1 = 0, 1)
                  f1[comp]   = e[6] + xi * (e[9] - e[6]);  // b2 = 1
                }
#pragma unroll
              for (int j = 0; j < N1; ++j)
                {
                  const T xj = a.xq[j];
#pragma unroll
                  for (int comp = 0; comp < 3; ++comp) jtab[slot * JT + ((d * N1 + i) * N1 + j) * 3 + comp] = f0[comp] + xj * (f1[comp] - f0[comp]);
                }
            }
        // (visible to phase B after the barrier that ends phase A)
      }

    // ---------------- phase A: gather + temporal contraction
    T v[N1][N1], w[N1][N1]; // [z][y]
#pragma unroll
    for (int k = 0; k < N1; ++k)
#pragma unroll
      for (int jy = 0; jy < N1; ++jy) v[k][jy] = w[k][jy] = T(0);
    if (active && !plane_constrained)
      {
        const T coef = a.coeff_cell ? a.coeff_cell[cell] : T(1);
        for (int s = 0; s < a.nb_src; ++s)
          {
            const T  be = a.beta[j * a.nb_src + s];
            const T  al = a.alpha[j * a.nb_src + s] * coef;
            const T *p  = a.src[s] + base;
#pragma unroll
            for (int k = 0; k < N1; ++k)
#pragma unroll
              for (int jy = 0; jy < N1; ++jy)
                {
                  const bool c = any_yz && ((ylo && jy == 0) || (yhi && jy == K) || (zlo && k == 0) || (zhi && k == K));
                  const T    u = c ? T(0) : p[jy * sy + k * sz];
                  v[k][jy] += be * u;
                  w[k][jy] += al * u;
                }
          }
      }
    // interpolation to the Gauss points in y, then z
#pragma unroll
    for (int k = 0; k < N1; ++k)
      {
        T t[N1];
        pl_apply<T, N1, false>(a.S, v[k], t);
#pragma unroll
        for (int q = 0; q < N1; ++q) v[k][q] = t[q];
        pl_apply<T, N1, false>(a.S, w[k], t);
#pragma unroll
        for (int q = 0; q < N1; ++q) w[k][q] = t[q];
      }
    {
      T *p0 = buf + cb * CBS + i; // field 0: v
      T *p1 = p0 + FS;            // field 1: w
      T *p3 = p0 + 3 * FS;        // field 3: gz
#pragma unroll
      for (int jy = 0; jy < N1; ++jy)
        {
          T in[N1], t[N1];
#pragma unroll
          for (int k = 0; k < N1; ++k) in[k] = v[k][jy];
          pl_apply<T, N1, false>(a.S, in, t);
#pragma unroll
          for (int q = 0; q < N1; ++q) p0[(q * N1 + jy) * LS] = t[q];
#pragma unroll
          for (int k = 0; k < N1; ++k) in[k] = w[k][jy];
          pl_apply<T, N1, false>(a.S, in, t);
#pragma unroll
          for (int q = 0; q < N1; ++q)
            {
              w[q][jy]                = t[q];
              p1[(q * N1 + jy) * LS] = t[q];
            }
          // z derivative of w on the Gauss points
          T g[N1];
          pl_apply<T, N1, false>(a.Dc, t, g);
#pragma unroll
          for (int q = 0; q < N1; ++q) p3[(q * N1 + jy) * LS] = g[q];
        }
      // y derivative
      T *p2 = p0 + 2 * FS; // field 2: gy
#pragma unroll
      for (int k = 0; k < N1; ++k)
        {
          T g[N1];
          pl_apply<T, N1, false>(a.Dc, w[k], g);
#pragma unroll
          for (int q = 0; q < N1; ++q) p2[(k * N1 + q) * LS] = g[q];
        }
    }
    __syncthreads();

    // ---------------- phase B: x lines (z = i, y = m): to the Gauss points in x, metric, back
    {
      const V2 *met = reinterpret_cast<const V2 *>(a.metric) + (size_t)cell * (N1 * N1 * 4 * N1);
      // One x line.  The N1 lines of a thread run as a real loop where that is faster (measured, Q4, 96^3 cells: on-the-fly
      // geometry FP64 6.03 -> 5.15 ms, FP32 2.60 -> 2.33 ms; stored metric FP32 2.10 -> 1.91 ms): the unrolled body (N1 times
      // ~600 instructions) overflows the instruction cache (ncu: one "no instruction" stall per issued instruction with the
      // 3 warps per scheduler this kernel runs with).  FP64 with the stored metric keeps the unrolled form (4.22 against 4.39 ms).
      auto x_line = [&](int m)
        {
          const int line = L::blocked ? N1 * i + m : i + N1 * m; // (z, y) = (i, m) for blocked, (m, i) else
          const int qy   = L::blocked ? m : i;
          const int qz   = L::blocked ? i : m;
          T        *pl   = buf + cb * CBS + line * LS;
          T         in[N1], vq[N1], wq[N1], gx[N1], gy[N1], gz[N1];
#pragma unroll
          for (int x = 0; x < N1; ++x) in[x] = pl[x];
          pl_apply<T, N1, false>(a.S, in, vq);
#pragma unroll
          for (int x = 0; x < N1; ++x) in[x] = pl[FS + x];
          pl_apply<T, N1, false>(a.S, in, wq);
          pl_apply<T, N1, false>(a.Dc, wq, gx);
#pragma unroll
          for (int x = 0; x < N1; ++x) in[x] = pl[2 * FS + x];
          pl_apply<T, N1, false>(a.S, in, gy);
#pragma unroll
          for (int x = 0; x < N1; ++x) in[x] = pl[3 * FS + x];
          pl_apply<T, N1, false>(a.S, in, gz);
          // quadrature-point operation
          T c0[3];
          if (OTF)
            {
#pragma unroll
              for (int comp = 0; comp < 3; ++comp) c0[comp] = jtab[slot * JT + ((0 * N1 + qy) * N1 + qz) * 3 + comp]; // column 0 at (eta, zeta)
            }
#pragma unroll
          for (int qx = 0; qx < N1; ++qx)
            {
              T Gxx, Gxy, Gxz, Gyy, Gyz, Gzz, JxW;
              if (OTF)
                {
                  T c1[3], c2[3];
#pragma unroll
                  for (int comp = 0; comp < 3; ++comp)
                    {
                      c1[comp] = jtab[slot * JT + ((1 * N1 + qx) * N1 + qz) * 3 + comp]; // column 1 at (xi, zeta)
                      c2[comp] = jtab[slot * JT + ((2 * N1 + qx) * N1 + qy) * 3 + comp]; // column 2 at (xi, eta)
                    }
                  // rows of det J * J^-1
                  const T r0[3] = {c1[1] * c2[2] - c1[2] * c2[1], c1[2] * c2[0] - c1[0] * c2[2], c1[0] * c2[1] - c1[1] * c2[0]};
                  const T r1[3] = {c2[1] * c0[2] - c2[2] * c0[1], c2[2] * c0[0] - c2[0] * c0[2], c2[0] * c0[1] - c2[1] * c0[0]};
                  const T r2[3] = {c0[1] * c1[2] - c0[2] * c1[1], c0[2] * c1[0] - c0[0] * c1[2], c0[0] * c1[1] - c0[1] * c1[0]};
                  const T det   = c0[0] * r0[0] + c0[1] * r0[1] + c0[2] * r0[2];
                  const T wgt   = a.wq[qx] * a.wq[qy] * a.wq[qz];
                  // 1 / det by two Newton steps from the single-precision reciprocal (46 and 92 bits): a third of the
                  // instructions of the IEEE division
                  T rd = (T)(1.0f / (float)det);
                  if (sizeof(T) == 8)
                    {
                      rd = rd * (T(2) - det * rd);
                      rd = rd * (T(2) - det * rd);
                    }
                  const T sc = wgt * rd;
                  Gxx = (r0[0] * r0[0] + r0[1] * r0[1] + r0[2] * r0[2]) * sc;
                  Gxy = (r0[0] * r1[0] + r0[1] * r1[1] + r0[2] * r1[2]) * sc;
                  Gxz = (r0[0] * r2[0] + r0[1] * r2[1] + r0[2] * r2[2]) * sc;
                  Gyy = (r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]) * sc;
                  Gyz = (r1[0] * r2[0] + r1[1] * r2[1] + r1[2] * r2[2]) * sc;
                  Gzz = (r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2]) * sc;
                  JxW = det * wgt;
                }
              else
                {
                  const V2 *mq  = met + ((size_t)(qy * N1 + qx) * 4) * N1 + qz;
                  const V2  m01 = mq[0], m23 = mq[N1], m45 = mq[2 * N1], m6 = mq[3 * N1];
                  Gxx = m01.x, Gxy = m01.y, Gxz = m23.x, Gyy = m23.y, Gyz = m45.x, Gzz = m45.y, JxW = m6.x;
                }
              const T   a0 = gx[qx], a1 = gy[qx], a2 = gz[qx];
              gx[qx] = Gxx * a0 + Gxy * a1 + Gxz * a2;
              gy[qx] = Gxy * a0 + Gyy * a1 + Gyz * a2;
              gz[qx] = Gxz * a0 + Gyz * a1 + Gzz * a2;
              vq[qx] = JxW * vq[qx];
            }
          // transposed x operations
          pl_apply_add<T, N1, true>(a.Dc, gx, vq); // vq += Dc^T g'x
          T o[N1];
          pl_apply<T, N1, true>(a.S, vq, o);
#pragma unroll
          for (int x = 0; x < N1; ++x) pl[x] = o[x];
          pl_apply<T, N1, true>(a.S, gy, o);
#pragma unroll
          for (int x = 0; x < N1; ++x) pl[FS + x] = o[x];
          pl_apply<T, N1, true>(a.S, gz, o);
#pragma unroll
          for (int x = 0; x < N1; ++x) pl[2 * FS + x] = o[x];
        };
      if constexpr (OTF || sizeof(T) == 4)
        {
#pragma unroll 1
          for (int m = 0; m < N1; ++m) x_line(m);
        }
      else
        {
#pragma unroll
          for (int m = 0; m < N1; ++m) x_line(m);
        }
    }
    __syncthreads();

    // ---------------- phase C: out = Sy^T Sz^T (F0 + Dc_y^T F1 + Dc_z^T F2), scatter-add
    {
      const T *p0 = buf + cb * CBS + i;
#pragma unroll
      for (int k = 0; k < N1; ++k)
        {
          T f1[N1];
#pragma unroll
          for (int q = 0; q < N1; ++q)
            {
              v[k][q] = p0[(k * N1 + q) * LS];
              f1[q]   = p0[FS + (k * N1 + q) * LS];
            }
          pl_apply_add<T, N1, true>(a.Dc, f1, v[k]);
        }
#pragma unroll
      for (int jy = 0; jy < N1; ++jy)
        {
          T f2[N1], col[N1], t[N1];
#pragma unroll
          for (int q = 0; q < N1; ++q)
            {
              f2[q]  = p0[2 * FS + (q * N1 + jy) * LS];
              col[q] = v[q][jy];
            }
          pl_apply_add<T, N1, true>(a.Dc, f2, col);
          pl_apply<T, N1, true>(a.S, col, t); // Sz^T
#pragma unroll
          for (int k = 0; k < N1; ++k) v[k][jy] = t[k];
        }
    }
    if (active && !plane_constrained)
      {
        T *d = a.dst[j] + base;
#pragma unroll
        for (int k = 0; k < N1; ++k)
          {
            T t[N1];
            pl_apply<T, N1, true>(a.S, v[k], t); // Sy^T
            const bool kc = (k == 0 && zlo) || (k == K && zhi);
#pragma unroll
            for (int jy = 0; jy < N1; ++jy)
              {
                const bool cn = kc || (jy == 0 && ylo) || (jy == K && yhi);
                if (!cn) atomicAdd(d + jy * sy + k * sz, t[jy]);
              }
          }
      }
  }

  // metric in the order the plane kernel reads it; one thread per (cell, q)
  template <typename T>
  __global__ void k_metric_to_plane_layout(const T *__restrict__ in, long long n_cells, int n1, T *__restrict__ out)
  {
    const int       nq  = n1 * n1 * n1;
    const long long tot = n_cells * nq;
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < tot; gid += (long long)gridDim.x * blockDim.x)
      {
        const long long cell = gid / nq;
        const int       q    = (int)(gid % nq);
        const int       qx = q % n1, qy = (q / n1) % n1, qz = q / (n1 * n1);
        const T        *m  = in + (size_t)gid * 7; // xx xy xz yy yz zz JxW
        T              *o  = out + (size_t)cell * nq * 8;
#pragma unroll
        for (int p = 0; p < 4; ++p)
          {
            const size_t idx = ((((size_t)qy * n1 + qx) * 4 + p) * n1 + qz) * 2;
            o[idx]     = m[2 * p];
            o[idx + 1] = p < 3 ? m[2 * p + 1] : T(0);
          }
      }
  }
} // namespace stfem
