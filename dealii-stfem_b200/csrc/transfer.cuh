// Two-level transfer operators of the space-time multigrid.
//   SpaceTransfer : deal.II MGTwoLevelTransfer as used by MGTwoLevelBlockTransfer
//                   (reference include/stmg.h:38-112, built at :594-602): nodal embedding of FE_Q between a
//                   coarse and a fine level (h: every cell split in 2^dim, p: degree change on the same mesh),
//                   restriction = exact transpose, constrained DoFs excluded.
//                   Tensor-product structure: one 1D pass per direction, all time blocks in one launch,
//                   no index tables, no atomics (every output entry is produced by exactly one thread).
//   TimeTransfer  : MGTwoLevelTransferTime (stmg.h:114-247): small dense matrix across the time blocks,
//                   one streaming pass (the reference does nnz(P) vector updates, operators.h:252-265).
#pragma once
#include <cstdlib>

#include "basis_host.hpp"
#include "fe_time.hpp"
#include "vec.cuh"

namespace stfem
{
  constexpr int TR_MAX_F = 13; // fine local nodes per coarse cell (2*6+1)
  constexpr int TR_MAX_C = 7;

  template <typename T>
  struct Transfer1D
  {
    T   P[TR_MAX_F * TR_MAX_C]; // P[lf * nc_loc + a]
    int sf, sc;                 // fine / coarse DoF strides per coarse cell
    int n_cells;                // coarse cells in this direction
  };

  // One tensor direction.  Array layout [blocks][n2][n1][n0] with the transfer acting on axis `axis`;
  // sizes given for the OUTPUT array (out_n) and the input extent along the axis (in_len).
  // mode 0: prolongation (out fine), 1: restriction (out coarse).  final: add into out and apply `mask`.
  template <typename T>
  __global__ void k_transfer_1d(Transfer1D<T> tr, int mode, int axis, int on0, int on1, int on2, int in_len, long long blocks,
                                const T *__restrict__ in, T *__restrict__ out, int final_add, unsigned dirichlet, int dim)
  {
    // 3D launch: x = node index along the fastest direction, y = second index, z = (third index, time block):
    // no 64-bit divisions per entry
    const long long out_per_block = (long long)on0 * on1 * on2;
    const int       nc_loc        = tr.sc + 1;
    const int       i0            = blockIdx.x * blockDim.x + threadIdx.x;
    if (i0 >= on0) return;
    const int i1 = blockIdx.y;
    for (int zb = blockIdx.z; zb < on2 * (int)blocks; zb += gridDim.z)
      {
        const int       b   = zb / on2;
        const int       i2  = zb - b * on2;
        const long long gid = (long long)b * out_per_block + (long long)i0 + (long long)on0 * (i1 + (long long)on1 * i2);
        int       idx[3] = {i0, i1, i2};
        const int o      = idx[axis];
        // input strides: same as output except along axis
        const int       in0 = axis == 0 ? in_len : on0, in1 = axis == 1 ? in_len : on1, in2 = axis == 2 ? in_len : on2;
        const long long in_per_block = (long long)in0 * in1 * in2;
        const long long stride       = axis == 0 ? 1 : (axis == 1 ? in0 : (long long)in0 * in1);
        idx[axis]                    = 0;
        const T *src = in + b * in_per_block + (long long)idx[0] + (long long)in0 * (idx[1] + (long long)in1 * idx[2]);
        T        s   = 0;
        if (mode == 0)
          {
            int c = o / tr.sf;
            if (c > tr.n_cells - 1) c = tr.n_cells - 1;
            const int lf = o - c * tr.sf;
            for (int a = 0; a < nc_loc; ++a) s += tr.P[lf * nc_loc + a] * src[stride * (c * tr.sc + a)];
          }
        else
          {
            int c_hi = o / tr.sc;
            int c_lo = (o % tr.sc == 0) ? c_hi - 1 : c_hi;
            if (c_lo < 0) c_lo = 0;
            if (c_hi > tr.n_cells - 1) c_hi = tr.n_cells - 1;
            for (int c = c_lo; c <= c_hi; ++c)
              {
                const int a      = o - c * tr.sc;
                const int lf_end = (c == tr.n_cells - 1) ? tr.sf : tr.sf - 1; // owned fine nodes of cell c
                for (int lf = 0; lf <= lf_end; ++lf) s += tr.P[lf * nc_loc + a] * src[stride * (c * tr.sf + lf)];
              }
          }
        if (final_add)
          {
            bool con = false;
            if (((dirichlet >> 0) & 1u) && i0 == 0) con = true;
            if (((dirichlet >> 1) & 1u) && i0 == on0 - 1) con = true;
            if (((dirichlet >> 2) & 1u) && i1 == 0) con = true;
            if (((dirichlet >> 3) & 1u) && i1 == on1 - 1) con = true;
            if (dim == 3)
              {
                if (((dirichlet >> 4) & 1u) && i2 == 0) con = true;
                if (((dirichlet >> 5) & 1u) && i2 == on2 - 1) con = true;
              }
            if (!con) out[gid] += s;
          }
        else
          out[gid] = s;
      }
  }

  // One tensor direction, one thread per COARSE CELL along the axis (compile-time sizes: SC = coarse intervals per
  // cell, SF = fine intervals per coarse cell; all P entries are immediate constant-bank operands):
  //   prolongation: SC+1 loads -> SF (last cell SF+1) outputs;  restriction: 2 SF+1 loads -> SC (last cell SC+1) outputs
  // 3D launch (x: fastest remaining index or cell, y, z x blocks), no divisions, no atomics.
  template <typename T, int SC, int SF>
  struct TransferLine
  {
    T P[(SF + 1) * (SC + 1)]; // P[lf * (SC+1) + a]
  };

  template <typename T, int SC, int SF, int MODE>
  __global__ void k_transfer_line(const __grid_constant__ TransferLine<T, SC, SF> tr, int axis, int n_cells, int on0, int on1, int on2,
                                  int in_len, int blocks, const T *__restrict__ in, T *__restrict__ out, int final_add, unsigned dirichlet,
                                  int dim)
  {
    // thread coordinates: along `axis` the index is the coarse cell, along the others the node index
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int t1 = blockIdx.y;
    const int ext0 = axis == 0 ? n_cells : on0;
    if (t0 >= ext0) return;
    const int ext2 = axis == 2 ? n_cells : on2;
    const int in0 = axis == 0 ? in_len : on0, in1 = axis == 1 ? in_len : on1, in2 = axis == 2 ? in_len : on2;
    const long long out_per_block = (long long)on0 * on1 * on2, in_per_block = (long long)in0 * in1 * in2;
    const long long os = axis == 0 ? 1 : (axis == 1 ? on0 : (long long)on0 * on1); // output stride along the axis
    const long long is = axis == 0 ? 1 : (axis == 1 ? in0 : (long long)in0 * in1);
    const int       on_axis = axis == 0 ? on0 : (axis == 1 ? on1 : on2);
    for (int zb = blockIdx.z; zb < ext2 * blocks; zb += gridDim.z)
      {
        const int b  = zb / ext2;
        const int t2 = zb - b * ext2;
        const int c  = axis == 0 ? t0 : (axis == 1 ? t1 : t2); // coarse cell along the axis
        int       idx[3] = {t0, t1, t2};
        idx[axis]        = 0;
        const T  *src = in + b * in_per_block + (long long)idx[0] + (long long)in0 * (idx[1] + (long long)in1 * idx[2]);
        T        *dst = out + b * out_per_block + (long long)idx[0] + (long long)on0 * (idx[1] + (long long)on1 * idx[2]);
        const bool last = c == n_cells - 1;
        // constraint of the two fixed coordinates (final pass only)
        bool fixed_con = false;
        if (final_add)
          {
            const int onn[3] = {on0, on1, on2};
#pragma unroll
            for (int d = 0; d < 3; ++d)
              if (d != axis && d < dim)
                {
                  if (((dirichlet >> (2 * d)) & 1u) && idx[d] == 0) fixed_con = true;
                  if (((dirichlet >> (2 * d + 1)) & 1u) && idx[d] == onn[d] - 1) fixed_con = true;
                }
          }
        const bool axis_lo = (dirichlet >> (2 * axis)) & 1u, axis_hi = (dirichlet >> (2 * axis + 1)) & 1u;
        if (MODE == 0)
          {
            T v[SC + 1];
#pragma unroll
            for (int a = 0; a <= SC; ++a) v[a] = src[is * (c * SC + a)];
#pragma unroll
            for (int lf = 0; lf <= SF; ++lf)
              {
                if (lf == SF && !last) break;
                T s = T(0);
#pragma unroll
                for (int a = 0; a <= SC; ++a) s += tr.P[lf * (SC + 1) + a] * v[a];
                const int o = c * SF + lf;
                if (final_add)
                  {
                    const bool con = fixed_con || (axis_lo && o == 0) || (axis_hi && o == on_axis - 1);
                    if (!con) dst[os * o] += s;
                  }
                else
                  dst[os * o] = s;
              }
          }
        else
          {
            T f[SF + 1], left[SF];
#pragma unroll
            for (int lf = 0; lf <= SF; ++lf) f[lf] = src[is * (c * SF + lf)];
            if (c > 0)
              {
#pragma unroll
                for (int lf = 0; lf < SF; ++lf) left[lf] = src[is * ((c - 1) * SF + lf)];
              }
            else
              {
#pragma unroll
                for (int lf = 0; lf < SF; ++lf) left[lf] = T(0);
              }
#pragma unroll
            for (int a = 0; a <= SC; ++a)
              {
                if (a == SC && !last) break;
                T s = T(0);
#pragma unroll
                for (int lf = 0; lf <= SF; ++lf) s += tr.P[lf * (SC + 1) + a] * f[lf];
                if (a == 0)
                  {
#pragma unroll
                    for (int lf = 0; lf < SF; ++lf) s += tr.P[lf * (SC + 1) + SC] * left[lf];
                  }
                const int o = c * SC + a;
                if (final_add)
                  {
                    const bool con = fixed_con || (axis_lo && o == 0) || (axis_hi && o == on_axis - 1);
                    if (!con) dst[os * o] += s;
                  }
                else
                  dst[os * o] = s;
              }
          }
      }
  }

  template <typename T>
  struct SpaceTransfer
  {
    stfem_ctx    *ctx = nullptr;
    int           dim = 3;
    int           npf[3] = {1, 1, 1}, npc[3] = {1, 1, 1};
    unsigned      dirichlet = 0;
    Transfer1D<T> tr[3];
    BlockVec<T>   tmp1, tmp2;
    int           nb_alloc = 0;

    // coarse: n_cells_c, degree kc ; fine: n_cells_f, degree kf
    int init(stfem_ctx *c, int dim_, const int *ncell_c, int kc, const int *ncell_f, int kf, unsigned dirichlet_)
    {
      ctx = c; dim = dim_; dirichlet = dirichlet_;
      const bool h = ncell_f[0] == 2 * ncell_c[0];
      STFEM_REQUIRE(h ? (kc == kf) : (ncell_f[0] == ncell_c[0] && kf >= kc), "space transfer: levels are neither h- nor p-related");
      STFEM_REQUIRE(kf <= 6 && kc <= 6, "space transfer: degree > 6");
      const auto gc = gauss_lobatto(kc + 1).x, gf = gauss_lobatto(kf + 1).x;
      for (int d = 0; d < 3; ++d)
        {
          npf[d] = d < dim ? kf * ncell_f[d] + 1 : 1;
          npc[d] = d < dim ? kc * ncell_c[d] + 1 : 1;
          if (d >= dim) continue;
          if (h) STFEM_REQUIRE(ncell_f[d] == 2 * ncell_c[d], "space transfer: anisotropic refinement");
          Transfer1D<T> &t = tr[d];
          t.n_cells = ncell_c[d];
          t.sc      = kc;
          t.sf      = h ? 2 * kf : kf;
          const int nc_loc = kc + 1;
          for (int lf = 0; lf <= t.sf; ++lf)
            {
              // position of the fine node in the coarse cell
              double x;
              if (h)
                {
                  const int child = lf < kf ? 0 : 1;
                  const int i     = lf - child * kf;
                  x               = 0.5 * (gf[i] + child);
                  if (lf == 2 * kf) x = 1.0;
                }
              else
                x = gf[lf];
              for (int a = 0; a < nc_loc; ++a)
                {
                  double v = lagrange_value(gc, a, x);
                  if (std::fabs(v) < 1e-15) v = 0.0;
                  t.P[lf * nc_loc + a] = (T)v;
                }
            }
        }
      return STFEM_OK;
    }

    int ensure_tmp(int nb)
    {
      if (nb <= nb_alloc) return STFEM_OK;
      const long long nf = (long long)npf[0] * npf[1] * npf[2];
      STFEM_FORWARD(tmp1.alloc(ctx, nb, nf));
      STFEM_FORWARD(tmp2.alloc(ctx, nb, nf));
      nb_alloc = nb;
      return STFEM_OK;
    }

    template <int SC, int SF>
    void launch_line(int mode, int axis, const int *on, int in_len, int nb, const T *in, T *out, bool fin)
    {
      TransferLine<T, SC, SF> tl;
      for (int i = 0; i < (SF + 1) * (SC + 1); ++i) tl.P[i] = tr[axis].P[i];
      const int nc   = tr[axis].n_cells;
      const int ext0 = axis == 0 ? nc : on[0], ext1 = axis == 1 ? nc : on[1], ext2 = axis == 2 ? nc : on[2];
      const int threads = ext0 >= 128 ? 128 : (ext0 >= 64 ? 64 : 32);
      const long long nz = (long long)ext2 * nb;
      const dim3 grid((ext0 + threads - 1) / threads, ext1, (unsigned)(nz < 65535 ? nz : 65535));
      if (mode == 0)
        k_transfer_line<T, SC, SF, 0><<<grid, threads, 0, ctx->stream>>>(tl, axis, nc, on[0], on[1], on[2], in_len, nb, in, out, fin ? 1 : 0, dirichlet, dim);
      else
        k_transfer_line<T, SC, SF, 1><<<grid, threads, 0, ctx->stream>>>(tl, axis, nc, on[0], on[1], on[2], in_len, nb, in, out, fin ? 1 : 0, dirichlet, dim);
      ctx->launches++;
    }

    void launch(int mode, int axis, const int *on, int in_len, int nb, const T *in, T *out, bool fin)
    {
      // compile-time sized per-cell kernel for the h transfers of degree 1..6 and the common p transfers
      const int sc = tr[axis].sc, sf = tr[axis].sf;
      static const bool legacy = std::getenv("STFEM_LEGACY_TRANSFER") != nullptr;
#define STFEM_TL(SC_, SF_) \
  if (!legacy && sc == SC_ && sf == SF_) return launch_line<SC_, SF_>(mode, axis, on, in_len, nb, in, out, fin);
      STFEM_TL(1, 2) STFEM_TL(2, 4) STFEM_TL(3, 6) STFEM_TL(4, 8) STFEM_TL(5, 10) STFEM_TL(6, 12)
      STFEM_TL(1, 2) STFEM_TL(1, 3) STFEM_TL(1, 4) STFEM_TL(2, 3) STFEM_TL(2, 4) STFEM_TL(2, 5) STFEM_TL(3, 4) STFEM_TL(3, 5) STFEM_TL(3, 6)
      STFEM_TL(4, 5) STFEM_TL(4, 6) STFEM_TL(5, 6) STFEM_TL(1, 5) STFEM_TL(1, 6) STFEM_TL(2, 6)
#undef STFEM_TL
      const int  threads = on[0] >= 128 ? 128 : (on[0] >= 64 ? 64 : 32);
      const long long nz = (long long)on[2] * nb;
      const dim3 grid((on[0] + threads - 1) / threads, on[1], (unsigned)(nz < 65535 ? nz : 65535));
      k_transfer_1d<T><<<grid, threads, 0, ctx->stream>>>(tr[axis], mode, axis, on[0], on[1], on[2], in_len, nb, in, out, fin ? 1 : 0,
                                                          dirichlet, dim);
      ctx->launches++;
    }

    // fine += P coarse
    int prolongate_and_add(BlockVec<T> &fine, const BlockVec<T> &coarse)
    {
      const int nb = coarse.nb;
      STFEM_FORWARD(ensure_tmp(nb));
      int       on[3] = {npf[0], npc[1], npc[2]};
      const T  *in    = coarse.d;
      T        *bufs[2] = {tmp1.d, tmp2.d};
      for (int d = 0; d < dim; ++d)
        {
          for (int e = 0; e < 3; ++e) on[e] = e <= d ? npf[e] : npc[e];
          const bool fin = d == dim - 1;
          T         *out = fin ? fine.d : bufs[d & 1];
          launch(0, d, on, npc[d], nb, in, out, fin);
          in = out;
        }
      return STFEM_OK;
    }

    // coarse += P^T fine
    int restrict_and_add(BlockVec<T> &coarse, const BlockVec<T> &fine)
    {
      const int nb = fine.nb;
      STFEM_FORWARD(ensure_tmp(nb));
      int      on[3];
      const T *in = fine.d;
      T       *bufs[2] = {tmp1.d, tmp2.d};
      int      step = 0;
      for (int d = dim - 1; d >= 0; --d, ++step)
        {
          for (int e = 0; e < 3; ++e) on[e] = e >= d ? npc[e] : npf[e];
          const bool fin = d == 0;
          T         *out = fin ? coarse.d : bufs[step & 1];
          launch(1, d, on, npf[d], nb, in, out, fin);
          in = out;
        }
      return STFEM_OK;
    }
  };

  template <typename T>
  struct TimeTransfer
  {
    std::vector<double> P, R; // P: nb_hi x nb_lo, R: nb_lo x nb_hi
    int                 nb_hi = 0, nb_lo = 0;

    // type: CGP/DG ; (nts_hi, nd_hi) / (nts_lo, nd_lo) block structure (get_blk_indices, stmg.h:460-501)
    int init(int type, int nts_hi, int nd_hi, int nts_lo, int nd_lo, bool restrict_is_transpose_prolongate, char mg_type)
    {
      const int r = type == DG ? nd_hi - 1 : nd_hi, r_lo = type == DG ? nd_lo - 1 : nd_lo;
      Mat       Pm, down;
      try
        {
          if (mg_type == 'k')
            {
              Pm   = time_projection_matrix(type, r_lo, r, nts_hi);
              down = time_projection_matrix(type, r, r_lo, nts_hi);
            }
          else
            {
              Pm   = time_prolongation_matrix(type, r, nts_hi);
              down = time_restriction_matrix(type, r, nts_hi);
            }
        }
      catch (const std::exception &e)
        {
          set_error("time transfer: %s", e.what());
          return STFEM_ERR_INVALID;
        }
      nb_hi = Pm.m; nb_lo = Pm.n;
      STFEM_REQUIRE(nb_hi == nts_hi * nd_hi && nb_lo == nts_lo * nd_lo, "time transfer: block structure mismatch (%d x %d vs %d, %d)",
                    nb_hi, nb_lo, nts_hi * nd_hi, nts_lo * nd_lo);
      P = Pm.a;
      R = restrict_is_transpose_prolongate ? Pm.transposed().a : down.a;
      return STFEM_OK;
    }
    void prolongate_and_add(BlockVec<T> &hi, const BlockVec<T> &lo) { v_block_matmul(hi, P, nb_hi, nb_lo, lo, true); }
    void restrict_and_add(BlockVec<T> &lo, const BlockVec<T> &hi) { v_block_matmul(lo, R, nb_lo, nb_hi, hi, true); }
  };
} // namespace stfem
