// Space-time multigrid V-cycle and FGMRES, host drivers over the device kernels.
// Replaces, for the reference's hot path:
//   GMG                               include/stmg.h:1047-1344   (reinit :1193-1329, vmult :1331-1344)
//   PreconditionSTMG                  include/stmg.h:968-1045    (Identity | Relaxation<Vanka> | Chebyshev<Vanka>)
//   STMGTransferBlockMatrixFree       include/stmg.h:304-458
//   deal.II Multigrid::level_v_step, MGSmootherPrecondition, MGCoarseGridApplySmoother,
//   PreconditionRelaxation / PreconditionChebyshev with power-iteration estimate   
//   deal.II SolverFGMRES + ReductionControl  (include/time_integrators.h:56-59, 315)
#pragma once
#include <cmath>
#include <cstdlib>
#include <memory>

#include "fe_time.hpp"
#include "op.hpp"
#include "transfer.cuh"
#include "vanka.cuh"
#include "vec.cuh"

namespace stfem
{
  template <typename T>
  __global__ void k_initial_guess(long long n, int nb, T *__restrict__ v)
  {
    // deal.II set_initial_guess: (i mod 11) minus its mean, per block
    const long long full = n / 11, rem = n % 11;
    const double    mean = (double)(full * 55 + rem * (rem - 1) / 2) / (double)n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * nb; i += (long long)gridDim.x * blockDim.x)
      v[i] = (T)((double)((i % n) % 11) - mean);
  }

  // the same start vector on a partitioned mesh: value from the GLOBAL lexicographic index, so that the copies
  // of an interface DoF agree on all ranks
  template <typename T>
  __global__ void k_initial_guess_part(int np0, int np1, int np2, int off0, int off1, int off2, long long gnp0, long long gnp1, long long gnp2,
                                       int nb, T *__restrict__ v)
  {
    const long long n = (long long)np0 * np1 * np2, ng = gnp0 * gnp1 * gnp2;
    const long long full = ng / 11, rem = ng % 11;
    const double    mean = (double)(full * 55 + rem * (rem - 1) / 2) / (double)ng;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * nb; i += (long long)gridDim.x * blockDim.x)
      {
        long long r  = i % n;
        const int ix = (int)(r % np0);
        r /= np0;
        const int       iy = (int)(r % np1), iz = (int)(r / np1);
        const long long g  = (long long)(ix + off0) + gnp0 * ((long long)(iy + off1) + gnp1 * (long long)(iz + off2));
        v[i] = (T)((double)(g % 11) - mean);
      }
  }

  // point-Jacobi inner preconditioner: inverse of diag_b = Alpha(b,b) diag K + Beta(b,b) diag M with the reference's
  // rule for (near-)zero entries (include/operators.h:1105-1109: |d| > sqrt(eps) ? 1/d : 1), and its application
  template <typename T>
  __global__ void k_inv_diag(long long n, double a, double b, const double *__restrict__ dK, const double *__restrict__ dM, T *__restrict__ out)
  {
    const double tol = sizeof(T) == 8 ? 1.4901161193847656e-08 : 3.4526698300124393e-04; // sqrt(epsilon)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      {
        const double d = a * dK[i] + b * dM[i];
        out[i]         = (T)(fabs(d) > tol ? 1.0 / d : 1.0);
      }
  }
  template <typename T>
  __global__ void k_diag_apply_add(long long n, T scale, const T *__restrict__ inv_diag, const T *__restrict__ src, T *__restrict__ dst)
  {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      dst[i] += scale * inv_diag[i] * src[i];
  }

  template <typename T>
  struct MGLevel
  {
    stfem_op                 *op = nullptr;
    int                       smoother = 1; // 0 identity, 1 relaxation, 2 chebyshev
    std::unique_ptr<Vanka<T>> vanka;
    BlockVec<T>               inv_diag; // point-Jacobi inner preconditioner (MGOptions::inner_preconditioner == 1)
    double                    omega = 1.0, theta = 1.0, delta = 0.0, lambda = 1.0;
    char                      ttype = 0; // transfer from level-1: 'h','p','k','t'
    SpaceTransfer<T>          st;
    TimeTransfer<T>           tt;
    BlockVec<T>               sol, defect, t, r, d, d2;
    int                       steps = 1;
    // coarse-level agglomeration: this level is partitioned, the next coarser one is the GLOBAL mesh held (and
    // solved redundantly) by every rank.  brick = this rank's part of the coarse level.
    bool        agglo = false;
    BlockVec<T> brick;
    int         brick_np[3] = {1, 1, 1}, brick_off[3] = {0, 0, 0};
  };

  struct MGOptions
  {
    int    smoothing_steps = 1;
    double relaxation = 0.0, smoothing_range = 1.0;
    int    eig_n_iterations = 20;
    bool   variable = true, restrict_is_transpose_prolongate = true;
    int    inner_preconditioner = 0; // 0 PreconditionVanka (the reference, stmg.h:1055-1063), 1 point-Jacobi
    int    vanka_storage = 0;        // 0 level precision (the reference: float), 1 FP16 (dense patch inverses only)
    int    coarse_gmres_maxiter = 0; // > 0: GMRES on the coarsest level instead of the smoother (stmg.h:1240-1302)
    double coarse_gmres_abstol = 1e-20;
  };

  struct MGBase
  {
    virtual ~MGBase() = default;
    virtual int vmult(void *const *dst, const void *const *src) = 0; // double block pointers (finest level)
    virtual int level_apply(int level, int what, void *const *dst, const void *const *src) = 0;
    virtual int level_info(int level, double *out) = 0;
    virtual int n_levels() const = 0;
    stfem_ctx *ctx = nullptr;
  };

  template <typename T>
  struct Multigrid : MGBase
  {
    std::vector<MGLevel<T>> L;
    MGOptions               opt;
    DotScratch              sc;
    BlockVec<double>        src64, dst64; // staging when T == float or the caller's blocks are not contiguous

    int n_levels() const override { return (int)L.size(); }

    int A(int l, BlockVec<T> &dst, const BlockVec<T> &src)
    {
      stfem_op *op = L[l].op;
      return op_apply(op, dst.block_ptrs(), src.cblock_ptrs(), op->nb_cols, op->nb_rows, op->d_alpha, op->d_beta, true);
    }

    // r = b - A x in one pass of the operator kernel (negated time matrices, rhs = b).  Partitioned meshes: every rank
    // contributes b / multiplicity at its interface nodes, so the one exchange of the partial sums also restores b.
    int residual(int l, BlockVec<T> &r, const BlockVec<T> &x, const BlockVec<T> &b)
    {
      stfem_op *op = L[l].op;
      return op_apply(op, r.block_ptrs(), x.cblock_ptrs(), op->nb_cols, op->nb_rows, op->d_alpha_neg, op->d_beta_neg, true, b.cblock_ptrs());
    }
    // dst += scale * P^-1 src  (Vanka); partitioned meshes: the increment is summed over ranks before it is added
    int vanka_add(int l, BlockVec<T> &dst, const BlockVec<T> &src, T scale)
    {
      MGLevel<T> &lv = L[l];
      if (lv.inv_diag.d)
        {
          k_diag_apply_add<T><<<grid_for(ctx, dst.size(), 256), 256, 0, ctx->stream>>>(dst.size(), scale, lv.inv_diag.d, src.d, dst.d);
          ctx->launches++;
          return STFEM_OK;
        }
      if (lv.op->mesh->part.active)
        {
          // the copies of an interface DoF are weighted with 1 / multiplicity, every rank adds the increments of its own
          // patches to them, and the exchange sums both: dst + scale * sum of all increments, without a temporary
          static const bool via_tmp = std::getenv("STFEM_VANKA_VIA_TMP") != nullptr; // round-1 path (A/B comparison)
          if (via_tmp)
            {
              if (tmp_part.size() <= (size_t)l) tmp_part.resize(L.size());
              if (!tmp_part[l].d) STFEM_FORWARD(tmp_part[l].alloc(ctx, src.nb, src.n));
              STFEM_FORWARD(lv.vanka->vmult(tmp_part[l], src));
              STFEM_FORWARD(halo_compress_add<T>(ctx, lv.op->mesh->part, lv.op->halo, tmp_part[l].block_ptrs(), src.nb, lv.op->np, lv.op->mesh->dim));
              v_axpy(dst, scale, tmp_part[l]);
              return STFEM_OK;
            }
          STFEM_FORWARD(halo_scale_interfaces<T>(ctx, lv.op->mesh->part, lv.op->halo, dst.block_ptrs(), dst.nb, lv.op->np, lv.op->mesh->dim));
          STFEM_FORWARD(lv.vanka->vmult_add(dst, src, scale));
          return halo_compress_add<T>(ctx, lv.op->mesh->part, lv.op->halo, dst.block_ptrs(), dst.nb, lv.op->np, lv.op->mesh->dim);
        }
      return lv.vanka->vmult_add(dst, src, scale);
    }
    std::vector<BlockVec<T>> tmp_part;

    // PreconditionSTMG::vmult(dst, src)   (stmg.h:1018-1023)
    int smoother_vmult(int l, BlockVec<T> &dst, const BlockVec<T> &src)
    {
      MGLevel<T> &lv = L[l];
      if (lv.smoother == 0) return v_copy(dst, src);
      if (lv.smoother == 1)
        {
          // PreconditionRelaxation: x = w P^-1 b ; n_it-1 times x += w P^-1 (b - A x)
          STFEM_FORWARD(dst.zero());
          STFEM_FORWARD(vanka_add(l, dst, src, (T)lv.omega));
          for (int it = 1; it < opt.smoothing_steps; ++it)
            {
              STFEM_FORWARD(residual(l, lv.d2, dst, src));
              STFEM_FORWARD(vanka_add(l, dst, lv.d2, (T)lv.omega));
            }
          return STFEM_OK;
        }
      // Chebyshev of degree smoothing_steps, zero start vector
      const double th = lv.theta, de = lv.delta;
      STFEM_FORWARD(lv.d2.zero());
      STFEM_FORWARD(vanka_add(l, lv.d2, src, (T)(1.0 / th))); // d
      STFEM_FORWARD(v_copy(dst, lv.d2));
      if (opt.smoothing_steps < 2) return STFEM_OK;
      const bool   fin = de != 0.0;
      const double sigma = fin ? th / de : 0.0;
      double       rho_old = fin ? 1.0 / sigma : 0.0;
      for (int it = 1; it < opt.smoothing_steps; ++it)
        {
          const double rho = fin ? 1.0 / (2.0 * sigma - rho_old) : 0.0;
          STFEM_FORWARD(residual(l, lv.t, dst, src)); // r = b - A x
          // d = rho*rho_old*d + (2 rho/delta) P^-1 r   (delta = 0: plain Richardson with 1/theta)
          v_scale(lv.d2, (T)(rho * rho_old));
          STFEM_FORWARD(vanka_add(l, lv.d2, lv.t, (T)(fin ? 2.0 * rho / de : 1.0 / th)));
          v_axpy(dst, (T)1, lv.d2);
          rho_old = rho;
        }
      return STFEM_OK;
    }

    // MGSmootherPrecondition::apply: u = P b, then steps-1 times u += P (b - A u)
    int mg_apply(int l, BlockVec<T> &u, const BlockVec<T> &rhs)
    {
      MGLevel<T> &lv = L[l];
      STFEM_FORWARD(smoother_vmult(l, u, rhs));
      for (int s = 1; s < lv.steps; ++s) STFEM_FORWARD(smooth_step(l, u, rhs));
      return STFEM_OK;
    }
    int smooth_step(int l, BlockVec<T> &u, const BlockVec<T> &rhs)
    {
      MGLevel<T> &lv = L[l];
      // the smoother uses lv.t/lv.d2/lv.r as scratch: the residual lives in lv.d
      STFEM_FORWARD(residual(l, lv.d, u, rhs));
      if (lv.smoother == 0)
        {
          v_axpy(u, (T)1, lv.d);
          return STFEM_OK;
        }
      if (lv.smoother == 1 && opt.smoothing_steps == 1) // u += w P^-1 r in one fused scatter
        return vanka_add(l, u, lv.d, (T)lv.omega);
      BlockVec<T> &corr = lv.r;
      STFEM_FORWARD(smoother_vmult(l, corr, lv.d));
      v_axpy(u, (T)1, corr);
      return STFEM_OK;
    }

    // coarse (+)= R fine.  Partitioned meshes: interface entries of the (complete) fine vector are weighted by
    // 1/multiplicity first, the local restrictions are then summed over the ranks (space transfers only; the time
    // transfers act DoF-wise).  `fine` is scaled in place.
    int restrict_level(int l, BlockVec<T> &coarse, BlockVec<T> &fine)
    {
      MGLevel<T> &lv = L[l];
      if (lv.ttype != 'h' && lv.ttype != 'p')
        {
          lv.tt.restrict_and_add(coarse, fine);
          return STFEM_OK;
        }
      const PartitionInfo &part = lv.op->mesh->part;
      if (part.active)
        STFEM_FORWARD(halo_scale_interfaces<T>(ctx, part, lv.op->halo, fine.block_ptrs(), fine.nb, lv.op->np, lv.op->mesh->dim));
      if (lv.agglo)
        {
          // restrict into this rank's brick of the coarse level, add the bricks into the global coarse vector
          // (interface partial sums included) with ONE all-reduce; all coarser levels need no communication
          stfem_op *oc  = L[l - 1].op;
          NcclApi  *api = nccl_api();
          STFEM_REQUIRE(api && ctx->nccl_comm, "multigrid: coarse-level agglomeration needs a communicator");
          STFEM_FORWARD(lv.brick.zero());
          STFEM_FORWARD(lv.st.restrict_and_add(lv.brick, fine));
          k_brick_global<T, true><<<grid_for(ctx, lv.brick.size(), 256), 256, 0, ctx->stream>>>(lv.brick.d, coarse.d, coarse.nb, lv.brick_np[0], lv.brick_np[1],
                                                                                               lv.brick_np[2], oc->np[0], oc->np[1], oc->np[2],
                                                                                               lv.brick_off[0], lv.brick_off[1], lv.brick_off[2]);
          ctx->launches++;
          STFEM_NCCL_CHECK(api->AllReduce(coarse.d, coarse.d, (size_t)coarse.size(), std::is_same<T, double>::value ? NcclApi::kDouble : NcclApi::kFloat,
                                          NcclApi::kSum, (nccl_comm_t)ctx->nccl_comm, ctx->stream));
          return STFEM_OK;
        }
      STFEM_FORWARD(lv.st.restrict_and_add(coarse, fine));
      if (part.active)
        {
          stfem_op *oc = L[l - 1].op;
          STFEM_FORWARD(halo_compress_add<T>(ctx, oc->mesh->part, oc->halo, coarse.block_ptrs(), coarse.nb, oc->np, oc->mesh->dim));
        }
      return STFEM_OK;
    }

    // fine += P coarse.  Agglomerated coarse level: rank 0's (redundantly computed) coarse solution is made the common
    // one by a broadcast - the redundant solves differ in the last bits (atomic summation order) and the copies of an
    // interface DoF must stay identical on all ranks - then this rank's brick is cut out and prolongated locally.
    int prolongate_level(int l, BlockVec<T> &fine, BlockVec<T> &coarse)
    {
      MGLevel<T> &lv = L[l];
      if (lv.ttype != 'h' && lv.ttype != 'p')
        {
          lv.tt.prolongate_and_add(fine, coarse);
          return STFEM_OK;
        }
      if (lv.agglo)
        {
          stfem_op *oc  = L[l - 1].op;
          NcclApi  *api = nccl_api();
          STFEM_REQUIRE(api && ctx->nccl_comm, "multigrid: coarse-level agglomeration needs a communicator");
          STFEM_NCCL_CHECK(api->Broadcast(coarse.d, coarse.d, (size_t)coarse.size(), std::is_same<T, double>::value ? NcclApi::kDouble : NcclApi::kFloat, 0,
                                          (nccl_comm_t)ctx->nccl_comm, ctx->stream));
          k_brick_global<T, false><<<grid_for(ctx, lv.brick.size(), 256), 256, 0, ctx->stream>>>(lv.brick.d, coarse.d, coarse.nb, lv.brick_np[0], lv.brick_np[1],
                                                                                                lv.brick_np[2], oc->np[0], oc->np[1], oc->np[2],
                                                                                                lv.brick_off[0], lv.brick_off[1], lv.brick_off[2]);
          ctx->launches++;
          return lv.st.prolongate_and_add(fine, lv.brick);
        }
      return lv.st.prolongate_and_add(fine, coarse);
    }

    // MGCoarseGridIterativeSolver (stmg.h:1267-1275, 1290-1298): SolverGMRES with deal.II's default LEFT preconditioning,
    // IterationNumberControl(maxiter, abstol) (reaching maxiter is success), at most maxiter basis vectors, zero start
    // vector, the coarse smoother as preconditioner:  min || P (b - A x) ||  over  span{P b, (P A) P b, ...}.
    // Vectors in level precision, inner products in double (fused multi-dot), Hessenberg / Givens on the host: this part
    // of the cycle is not captured in a CUDA graph (the cycle is split around it, see cycle()).
    std::vector<BlockVec<T>> cg_V;
    BlockVec<T>              cg_w;
    int coarse_gmres(BlockVec<T> &x, const BlockVec<T> &b)
    {
      MGLevel<T> &lv = L[0];
      const int   m  = opt.coarse_gmres_maxiter;
      if ((int)cg_V.size() < m + 1)
        {
          cg_V.resize(m + 1);
          for (auto &v : cg_V) STFEM_FORWARD(v.alloc(ctx, b.nb, b.n));
          STFEM_FORWARD(cg_w.alloc(ctx, b.nb, b.n));
        }
      sc.set_partition(lv.op->mesh->part, lv.op->np, lv.op->mesh->dim);
      STFEM_FORWARD(x.zero());
      STFEM_FORWARD(smoother_vmult(0, cg_V[0], b)); // r = P b
      double beta2 = 0;
      STFEM_FORWARD(v_dot(sc, cg_V[0], cg_V[0], &beta2));
      const double beta = std::sqrt(beta2);
      if (!(beta > 0)) return STFEM_OK;
      v_scale(cg_V[0], (T)(1.0 / beta));
      std::vector<double> H((size_t)(m + 1) * m, 0.0), g(m + 1, 0.0), cs(m, 0.0), sn(m, 0.0), h(m + 2, 0.0);
      auto Hm = [&](int i, int j) -> double & { return H[(size_t)i * m + j]; };
      g[0]   = beta;
      int jd = 0;
      for (int j = 0; j < m; ++j)
        {
          STFEM_FORWARD(A(0, lv.r, cg_V[j]));
          STFEM_FORWARD(smoother_vmult(0, cg_w, lv.r)); // w = P A v_j   (lv.r: smoother_vmult itself works in lv.d2 / lv.t)
          std::vector<const BlockVec<T> *> basis;
          for (int i = 0; i <= j; ++i) basis.push_back(&cg_V[i]);
          for (int pass = 0; pass < 2; ++pass)
            {
              STFEM_FORWARD(v_multi_dot(sc, cg_w, basis, h.data()));
              std::vector<double> c(j + 1);
              for (int i = 0; i <= j; ++i)
                {
                  c[i] = -h[i];
                  Hm(i, j) += h[i];
                }
              v_multi_axpy(cg_w, basis, c.data());
            }
          double wn2 = 0;
          STFEM_FORWARD(v_dot(sc, cg_w, cg_w, &wn2));
          const double hn = std::sqrt(std::max(wn2, 0.0));
          Hm(j + 1, j)    = hn;
          for (int i = 0; i < j; ++i)
            {
              const double t = cs[i] * Hm(i, j) + sn[i] * Hm(i + 1, j);
              Hm(i + 1, j)   = -sn[i] * Hm(i, j) + cs[i] * Hm(i + 1, j);
              Hm(i, j)       = t;
            }
          const double den = std::hypot(Hm(j, j), Hm(j + 1, j));
          if (den == 0.0) break;
          cs[j]        = Hm(j, j) / den;
          sn[j]        = Hm(j + 1, j) / den;
          Hm(j, j)     = den;
          Hm(j + 1, j) = 0.0;
          g[j + 1]     = -sn[j] * g[j];
          g[j]         = cs[j] * g[j];
          jd           = j + 1;
          if (std::fabs(g[j + 1]) < opt.coarse_gmres_abstol || hn == 0.0) break;
          if (j + 1 < m) v_scale_copy(cg_V[j + 1], (T)(1.0 / hn), cg_w);
        }
      if (jd == 0) return STFEM_OK;
      std::vector<double> y(jd);
      for (int i = jd - 1; i >= 0; --i)
        {
          double s_ = g[i];
          for (int k = i + 1; k < jd; ++k) s_ -= Hm(i, k) * y[k];
          y[i] = s_ / Hm(i, i);
        }
      std::vector<const BlockVec<T> *> vs;
      for (int i = 0; i < jd; ++i) vs.push_back(&cg_V[i]);
      v_multi_axpy(x, vs, y.data());
      return STFEM_OK;
    }
    int coarse_solve()
    {
      if (opt.coarse_gmres_maxiter > 0) return coarse_gmres(L[0].sol, L[0].defect);
      return mg_apply(0, L[0].sol, L[0].defect);
    }
    // the two halves of the V-cycle around the coarse solve (each a fixed launch sequence): down = pre-smoothing, residual
    // and restriction on levels top..1; up = prolongation and post-smoothing on levels 1..top
    int v_down(int top)
    {
      for (int l = top; l >= 1; --l)
        {
          MGLevel<T> &lv = L[l];
          STFEM_FORWARD(mg_apply(l, lv.sol, lv.defect));
          STFEM_FORWARD(residual(l, lv.d, lv.sol, lv.defect));
          STFEM_FORWARD(L[l - 1].defect.zero());
          STFEM_FORWARD(restrict_level(l, L[l - 1].defect, lv.d));
        }
      return STFEM_OK;
    }
    int v_up(int top)
    {
      for (int l = 1; l <= top; ++l)
        {
          MGLevel<T> &lv = L[l];
          STFEM_FORWARD(prolongate_level(l, lv.sol, L[l - 1].sol));
          for (int s_ = 0; s_ < lv.steps; ++s_) STFEM_FORWARD(smooth_step(l, lv.sol, lv.defect));
        }
      return STFEM_OK;
    }

    // Multigrid::level_v_step; defect in L[l].defect, result in L[l].sol
    int v_step(int l)
    {
      MGLevel<T> &lv = L[l];
      if (l == 0) return coarse_solve();
      STFEM_FORWARD(mg_apply(l, lv.sol, lv.defect));
      // t = defect - A sol
      {
        BlockVec<T> &res = lv.d;
        STFEM_FORWARD(residual(l, res, lv.sol, lv.defect));
        MGLevel<T> &lc = L[l - 1];
        STFEM_FORWARD(lc.defect.zero());
        STFEM_FORWARD(restrict_level(l, lc.defect, res));
      }
      STFEM_FORWARD(v_step(l - 1));
      {
        MGLevel<T> &lc = L[l - 1];
        STFEM_FORWARD(prolongate_level(l, lv.sol, lc.sol));
      }
      for (int s = 0; s < lv.steps; ++s) STFEM_FORWARD(smooth_step(l, lv.sol, lv.defect));
      return STFEM_OK;
    }

    int estimate(int l)
    {
      MGLevel<T> &lv = L[l];
      if (lv.smoother == 0) return STFEM_OK;
      if (lv.smoother == 1 && opt.relaxation != 0.0)
        {
          lv.omega = opt.relaxation;
          return STFEM_OK;
        }
      // power iteration on P^-1 A (A.7)
      BlockVec<T> &v = lv.sol, &w = lv.defect, &tmp = lv.t;
      const PartitionInfo &part = lv.op->mesh->part;
      sc.set_partition(part, lv.op->np, lv.op->mesh->dim);
      if (part.active)
        {
          const int k = lv.op->degree, *np = lv.op->np, *nl = lv.op->mesh->n;
          long long gnp[3];
          int       off[3];
          for (int d = 0; d < 3; ++d)
            {
              gnp[d] = d < lv.op->mesh->dim ? (long long)k * nl[d] * part.grid[d] + 1 : 1;
              off[d] = d < lv.op->mesh->dim ? k * nl[d] * part.coords[d] : 0;
            }
          k_initial_guess_part<T><<<grid_for(ctx, v.size(), 256), 256, 0, ctx->stream>>>(np[0], np[1], np[2], off[0], off[1], off[2], gnp[0], gnp[1],
                                                                                        gnp[2], v.nb, v.d);
        }
      else
        k_initial_guess<T><<<grid_for(ctx, v.size(), 256), 256, 0, ctx->stream>>>(v.n, v.nb, v.d);
      ctx->launches++;
      double nrm2 = 0, lam = 1.0;
      STFEM_FORWARD(v_dot(sc, v, v, &nrm2));
      if (nrm2 > 0)
        {
          v_scale(v, (T)(1.0 / std::sqrt(nrm2)));
          for (int it = 0; it < opt.eig_n_iterations; ++it)
            {
              STFEM_FORWARD(A(l, tmp, v));
              STFEM_FORWARD(w.zero());
              STFEM_FORWARD(vanka_add(l, w, tmp, (T)1));
              double                           out[2];
              std::vector<const BlockVec<T> *> V{&v, &w};
              STFEM_FORWARD(v_multi_dot(sc, w, V, out));
              lam = out[0];
              const double nw = std::sqrt(out[1]);
              if (!(nw > 0) || !std::isfinite(nw))
                {
                  lam = 1.0; // guard: start vector in the kernel of A (1-cell coarse levels)
                  break;
                }
              STFEM_FORWARD(v_copy(v, w));
              v_scale(v, (T)(1.0 / nw));
            }
        }
      lv.lambda = lam;
      const double lmax = 1.2 * lam;
      const double rng  = opt.smoothing_range;
      const double lmin = rng > 1.0 ? lam / rng : lam;
      const double alpha = rng > 1.0 ? lmax / rng : std::min(0.9 * lmax, lmin);
      if (lv.smoother == 1)
        lv.omega = 2.0 / (alpha + lmax);
      else
        {
          lv.theta = 0.5 * (lmax + alpha);
          lv.delta = 0.5 * (lmax - alpha);
        }
      return STFEM_OK;
    }

    int init(stfem_ctx *c, const std::vector<stfem_op *> &ops, const std::string &types, const std::vector<int> &smoothers, int time_type,
             int nts, const std::vector<int> &poly_time, const MGOptions &o)
    {
      ctx = c; opt = o;
      const int nl = (int)ops.size();
      STFEM_REQUIRE(nl >= 1 && (int)types.size() == nl - 1 && (int)smoothers.size() == nl, "mg: inconsistent level description");
      L.resize(nl);
      // block structure per level (get_blk_indices, stmg.h:460-501)
      std::vector<int> lv_nts(nl), lv_nd(nl);
      {
        int p = (int)poly_time.size() - 1, n = nts;
        for (int i = nl - 1; i >= 1; --i)
          {
            STFEM_REQUIRE(p >= 0, "mg: poly_time_sequence too short for the number of k levels");
            lv_nts[i] = n;
            lv_nd[i]  = time_type == DG ? poly_time[p] + 1 : poly_time[p];
            if (types[i - 1] == 'k') --p;
            else if (types[i - 1] == 't') n /= 2;
          }
        STFEM_REQUIRE(p >= 0, "mg: poly_time_sequence too short");
        lv_nts[0] = n;
        lv_nd[0]  = time_type == DG ? poly_time[p] + 1 : poly_time[p];
      }
      for (int l = 0; l < nl; ++l)
        {
          MGLevel<T> &lv = L[l];
          lv.op       = ops[l];
          lv.smoother = smoothers[l];
          lv.steps    = opt.variable ? (1 << (nl - 1 - l)) : 1;
          STFEM_REQUIRE(lv.op->nb_rows == lv_nts[l] * lv_nd[l], "mg: level %d operator has %d blocks, level structure says %d x %d", l,
                        lv.op->nb_rows, lv_nts[l], lv_nd[l]);
          const int nb = lv.op->nb_rows;
          for (BlockVec<T> *v : {&lv.sol, &lv.defect, &lv.t, &lv.r, &lv.d, &lv.d2}) STFEM_FORWARD(v->alloc(ctx, nb, lv.op->N));
          if (lv.smoother != 0 && opt.inner_preconditioner == 1)
            {
              double *dK = nullptr, *dM = nullptr;
              STFEM_FORWARD(op_spatial_diagonals(lv.op, &dK, &dM));
              STFEM_FORWARD(lv.inv_diag.alloc(ctx, nb, lv.op->N));
              for (int b = 0; b < nb; ++b)
                {
                  k_inv_diag<T><<<grid_for(ctx, lv.op->N, 256), 256, 0, ctx->stream>>>(lv.op->N, lv.op->Alpha[(size_t)b * nb + b],
                                                                                       lv.op->Beta[(size_t)b * nb + b], dK, dM,
                                                                                       lv.inv_diag.d + (size_t)b * lv.op->N);
                  ctx->launches++;
                }
              STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
              cudaFree(dK);
              cudaFree(dM);
            }
          else if (lv.smoother != 0)
            {
              lv.vanka = std::make_unique<Vanka<T>>();
              STFEM_FORWARD(lv.vanka->setup(lv.op, opt.vanka_storage == 1));
            }
          if (l > 0)
            {
              lv.ttype = types[l - 1];
              stfem_op *oc = ops[l - 1];
              const bool boundary = lv.op->mesh->part.active && !oc->mesh->part.active; // partitioned above, global below
              if (lv.ttype == 'h' || lv.ttype == 'p')
                {
                  STFEM_REQUIRE(oc->nb_rows == nb, "mg: space transfer between levels with different block counts");
                  if (boundary)
                    {
                      // the coarse level is the global mesh on every rank: transfers act on this rank's brick of it
                      const PartitionInfo &part = lv.op->mesh->part;
                      const int            dim  = lv.op->mesh->dim;
                      int                  nloc[3] = {1, 1, 1};
                      long long            nbrick = 1;
                      for (int d = 0; d < 3; ++d)
                        {
                          if (d < dim)
                            {
                              nloc[d] = lv.ttype == 'h' ? lv.op->mesh->n[d] / 2 : lv.op->mesh->n[d];
                              STFEM_REQUIRE(nloc[d] * part.grid[d] == oc->mesh->n[d], "mg: global coarse mesh does not match the partitioned fine level");
                            }
                          lv.brick_np[d]  = d < dim ? oc->degree * nloc[d] + 1 : 1;
                          lv.brick_off[d] = d < dim ? oc->degree * nloc[d] * part.coords[d] : 0;
                          nbrick *= lv.brick_np[d];
                        }
                      lv.agglo = true;
                      STFEM_FORWARD(lv.brick.alloc(ctx, nb, nbrick));
                      STFEM_FORWARD(lv.st.init(ctx, dim, nloc, oc->degree, lv.op->mesh->n, lv.op->degree, lv.op->mesh->dirichlet));
                    }
                  else
                    STFEM_FORWARD(lv.st.init(ctx, lv.op->mesh->dim, oc->mesh->n, oc->degree, lv.op->mesh->n, lv.op->degree, lv.op->mesh->dirichlet));
                }
              else
                {
                  STFEM_REQUIRE(!boundary, "mg: the switch from partitioned to agglomerated levels must be a space (h/p) transfer");
                  STFEM_REQUIRE(oc->N == lv.op->N, "mg: time transfer between levels with different spatial size");
                  STFEM_FORWARD(lv.tt.init(time_type, lv_nts[l], lv_nd[l], lv_nts[l - 1], lv_nd[l - 1], opt.restrict_is_transpose_prolongate, lv.ttype));
                }
            }
        }
      for (int l = 0; l < nl; ++l) STFEM_FORWARD(estimate(l));
      return STFEM_OK;
    }

    static bool contiguous(const void *const *p, int nb, long long n, size_t elem)
    {
      for (int b = 1; b < nb; ++b)
        if ((const char *)p[b] != (const char *)p[0] + (size_t)b * n * elem) return false;
      return true;
    }
    int stage_in(BlockVec<T> &dst, const void *const *src, int nb, long long n)
    {
      if (std::is_same<T, double>::value)
        {
          for (int b = 0; b < nb; ++b)
            STFEM_CUDA_CHECK(cudaMemcpyAsync((char *)dst.d + sizeof(T) * (size_t)b * n, src[b], sizeof(T) * n, cudaMemcpyDeviceToDevice, ctx->stream));
          return STFEM_OK;
        }
      const double *from = (const double *)src[0];
      if (!contiguous(src, nb, n, sizeof(double)))
        {
          if (!src64.d) STFEM_FORWARD(src64.alloc(ctx, nb, n));
          for (int b = 0; b < nb; ++b)
            STFEM_CUDA_CHECK(cudaMemcpyAsync(src64.d + (size_t)b * n, src[b], sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
          from = src64.d;
        }
      k_convert<T, double><<<grid_for(ctx, dst.size(), 256), 256, 0, ctx->stream>>>(dst.size(), from, dst.d);
      ctx->launches++;
      return STFEM_OK;
    }
    int stage_out(void *const *dst, const BlockVec<T> &src, int nb, long long n)
    {
      if (std::is_same<T, double>::value)
        {
          for (int b = 0; b < nb; ++b)
            STFEM_CUDA_CHECK(cudaMemcpyAsync(dst[b], (const char *)src.d + sizeof(T) * (size_t)b * n, sizeof(T) * n, cudaMemcpyDeviceToDevice, ctx->stream));
          return STFEM_OK;
        }
      if (contiguous(dst, nb, n, sizeof(double)))
        {
          k_convert<double, T><<<grid_for(ctx, src.size(), 256), 256, 0, ctx->stream>>>(src.size(), src.d, (double *)dst[0]);
          ctx->launches++;
          return STFEM_OK;
        }
      if (!dst64.d) STFEM_FORWARD(dst64.alloc(ctx, nb, n));
      k_convert<double, T><<<grid_for(ctx, src.size(), 256), 256, 0, ctx->stream>>>(src.size(), src.d, dst64.d);
      ctx->launches++;
      for (int b = 0; b < nb; ++b)
        STFEM_CUDA_CHECK(cudaMemcpyAsync(dst[b], dst64.d + (size_t)b * n, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
      return STFEM_OK;
    }

    // The V-cycle between the two precision copies is a fixed launch sequence without host decisions: it is
    // captured once into a CUDA graph (after one eager pass has made every lazy allocation) and replayed.
    cudaGraphExec_t graph_exec = nullptr, graph_up = nullptr;
    long long       graph_launches = 0, graph_up_launches = 0;
    int             n_vcycles = 0;
    ~Multigrid()
    {
      if (graph_exec) cudaGraphExecDestroy(graph_exec);
      if (graph_up) cudaGraphExecDestroy(graph_up);
    }
    int cycle()
    {
      const int  top      = (int)L.size() - 1;
      static const bool no_graph = std::getenv("STFEM_NO_GRAPH") != nullptr;
      const bool part     = L[top].op->mesh->part.active;
      bool       timing   = false;
      for (auto &lv : L) timing = timing || lv.op->timing;
      static const bool no_part_graph = std::getenv("STFEM_NO_PART_GRAPH") != nullptr;
      if (no_graph || (part && no_part_graph) || timing || n_vcycles++ == 0) return v_step(top);
      if (opt.coarse_gmres_maxiter > 0)
        {
          // iterative coarse solver: its Hessenberg updates need the host, so the two halves of the cycle are captured
          // separately and the coarse GMRES runs between them
          if (top == 0) return coarse_solve();
          auto capture = [&](cudaGraphExec_t &exec, long long &launches, bool down) -> int {
            cudaGraph_t     graph = nullptr;
            const long long l0    = ctx->launches;
            STFEM_CUDA_CHECK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
            const int         rc = down ? v_down(top) : v_up(top);
            const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
            if (rc != STFEM_OK || ce != cudaSuccess)
              {
                if (graph) cudaGraphDestroy(graph);
                if (rc != STFEM_OK) return rc;
                STFEM_CUDA_CHECK(ce);
              }
            launches      = ctx->launches - l0;
            ctx->launches = l0;
            STFEM_CUDA_CHECK(cudaGraphInstantiate(&exec, graph, 0));
            cudaGraphDestroy(graph);
            return STFEM_OK;
          };
          if (!graph_exec) STFEM_FORWARD(capture(graph_exec, graph_launches, true));
          if (!graph_up) STFEM_FORWARD(capture(graph_up, graph_up_launches, false));
          STFEM_CUDA_CHECK(cudaGraphLaunch(graph_exec, ctx->stream));
          ctx->launches += graph_launches;
          STFEM_FORWARD(coarse_solve());
          STFEM_CUDA_CHECK(cudaGraphLaunch(graph_up, ctx->stream));
          ctx->launches += graph_up_launches;
          return STFEM_OK;
        }
      if (!graph_exec)
        {
          cudaGraph_t     graph = nullptr;
          const long long l0    = ctx->launches;
          STFEM_CUDA_CHECK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
          const int rc = v_step(top);
          const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
          if (rc != STFEM_OK || ce != cudaSuccess)
            {
              if (graph) cudaGraphDestroy(graph);
              if (rc != STFEM_OK) return rc;
              STFEM_CUDA_CHECK(ce);
            }
          graph_launches = ctx->launches - l0;
          ctx->launches  = l0;
          STFEM_CUDA_CHECK(cudaGraphInstantiate(&graph_exec, graph, 0));
          cudaGraphDestroy(graph);
        }
      STFEM_CUDA_CHECK(cudaGraphLaunch(graph_exec, ctx->stream));
      ctx->launches += graph_launches;
      return STFEM_OK;
    }

    // GMG::vmult (stmg.h:1331-1344): double in, one V-cycle in level precision, double out
    int vmult(void *const *dst, const void *const *src) override
    {
      MGLevel<T> &top = L.back();
      STFEM_FORWARD(stage_in(top.defect, src, top.op->nb_rows, top.op->N));
      STFEM_FORWARD(cycle());
      return stage_out(dst, top.sol, top.op->nb_rows, top.op->N);
    }

    // unit-test hooks; vectors in LEVEL precision, given as block pointers
    int level_apply(int l, int what, void *const *dst, const void *const *src) override
    {
      STFEM_REQUIRE(l >= 0 && l < (int)L.size(), "mg level %d out of range", l);
      MGLevel<T> &lv = L[l];
      auto load = [&](BlockVec<T> &v, const void *const *p) -> int {
        for (int b = 0; b < v.nb; ++b)
          STFEM_CUDA_CHECK(cudaMemcpyAsync(v.d + (size_t)b * v.n, p[b], sizeof(T) * v.n, cudaMemcpyDeviceToDevice, ctx->stream));
        return STFEM_OK;
      };
      auto store = [&](void *const *p, const BlockVec<T> &v) -> int {
        for (int b = 0; b < v.nb; ++b)
          STFEM_CUDA_CHECK(cudaMemcpyAsync(p[b], v.d + (size_t)b * v.n, sizeof(T) * v.n, cudaMemcpyDeviceToDevice, ctx->stream));
        return STFEM_OK;
      };
      switch (what)
        {
          case 0: // Vanka
            STFEM_REQUIRE(lv.vanka || lv.inv_diag.d, "level %d has no inner preconditioner", l);
            STFEM_FORWARD(load(lv.defect, src));
            STFEM_FORWARD(lv.sol.zero());
            STFEM_FORWARD(vanka_add(l, lv.sol, lv.defect, (T)1));
            return store(dst, lv.sol);
          case 1: // PreconditionSTMG::vmult
            STFEM_FORWARD(load(lv.defect, src));
            STFEM_FORWARD(smoother_vmult(l, lv.sol, lv.defect));
            return store(dst, lv.sol);
          case 2: // restrict level l -> l-1 (dst zeroed)
            STFEM_REQUIRE(l > 0, "no coarser level");
            STFEM_FORWARD(load(lv.d, src));
            STFEM_FORWARD(L[l - 1].defect.zero());
            STFEM_FORWARD(restrict_level(l, L[l - 1].defect, lv.d));
            return store(dst, L[l - 1].defect);
          case 3: // prolongate level l-1 -> l (dst zeroed)
            STFEM_REQUIRE(l > 0, "no coarser level");
            STFEM_FORWARD(load(L[l - 1].sol, src));
            STFEM_FORWARD(lv.sol.zero());
            STFEM_FORWARD(prolongate_level(l, lv.sol, L[l - 1].sol));
            return store(dst, lv.sol);
          case 4: // level operator
            STFEM_FORWARD(load(lv.defect, src));
            STFEM_FORWARD(A(l, lv.sol, lv.defect));
            return store(dst, lv.sol);
          case 5: // one V-cycle starting at this level (defect -> correction)
            STFEM_FORWARD(load(lv.defect, src));
            STFEM_FORWARD(v_step(l));
            return store(dst, lv.sol);
          default: set_error("mg level_apply: unknown operation %d", what); return STFEM_ERR_INVALID;
        }
    }

    int level_info(int l, double *out) override
    {
      STFEM_REQUIRE(l >= 0 && l < (int)L.size() && out, "mg level_info: bad arguments");
      MGLevel<T> &lv = L[l];
      out[0] = lv.smoother; out[1] = lv.lambda; out[2] = lv.omega; out[3] = lv.theta; out[4] = lv.delta; out[5] = lv.steps;
      out[6] = (double)lv.op->N; out[7] = lv.op->nb_rows; out[8] = lv.vanka ? (double)lv.vanka->n_mat : 0.0;
      out[9] = lv.vanka ? (double)lv.vanka->bytes : 0.0;
      return STFEM_OK;
    }
  };

  // ------------------------------------------------------------------ FGMRES (double)
  struct FgmresResult { int iterations = 0; double initial_residual = 0, final_residual = 0; bool converged = false; };

  struct Fgmres
  {
    stfem_ctx                     *ctx = nullptr;
    std::vector<BlockVec<double>>  V, Z;
    BlockVec<double>               x, b, w;
    DotScratch                     sc;

    int ensure(std::vector<BlockVec<double>> &a, size_t i, int nb, long long n)
    {
      while (a.size() <= i) a.emplace_back();
      if (!a[i].d || a[i].nb != nb || a[i].n != n) STFEM_FORWARD(a[i].alloc(ctx, nb, n));
      return STFEM_OK;
    }

    // right-preconditioned flexible GMRES(max_basis), ReductionControl(max_iter, abstol, reduce)
    int solve(stfem_op *A, MGBase *M, void *const *x_blocks, const void *const *b_blocks, int max_basis, int max_iter, double abstol,
              double reduce, FgmresResult &res)
    {
      ctx = A->mesh->ctx;
      const int       nb = A->nb_rows;
      const long long n  = A->N;
      STFEM_REQUIRE(A->number_type == STFEM_F64, "fgmres: the outer operator must be double precision");
      sc.set_partition(A->mesh->part, A->np, A->mesh->dim);
      if (!x.d || x.nb != nb || x.n != n)
        {
          STFEM_FORWARD(x.alloc(ctx, nb, n));
          STFEM_FORWARD(b.alloc(ctx, nb, n));
          STFEM_FORWARD(w.alloc(ctx, nb, n));
        }
      for (int k = 0; k < nb; ++k)
        {
          STFEM_CUDA_CHECK(cudaMemcpyAsync(x.d + (size_t)k * n, x_blocks[k], sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
          STFEM_CUDA_CHECK(cudaMemcpyAsync(b.d + (size_t)k * n, b_blocks[k], sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        }
      auto apply_A = [&](BlockVec<double> &dst, const BlockVec<double> &src) {
        return op_apply(A, dst.block_ptrs(), src.cblock_ptrs(), nb, nb, A->d_alpha, A->d_beta, true);
      };
      res = FgmresResult();
      double tol = 0;
      bool   first = true;
      std::vector<double> H((size_t)(max_basis + 1) * max_basis), g(max_basis + 1), cs(max_basis), sn(max_basis), h(max_basis + 2);
      auto Hm = [&](int i, int j) -> double & { return H[(size_t)i * max_basis + j]; };
      while (true)
        {
          // r = b - A x
          STFEM_FORWARD(ensure(V, 0, nb, n));
          STFEM_FORWARD(apply_A(V[0], x));
          v_sadd(V[0], -1.0, 1.0, b);
          double beta2 = 0;
          STFEM_FORWARD(v_dot(sc, V[0], V[0], &beta2));
          const double beta = std::sqrt(beta2);
          if (!std::isfinite(beta))
            {
              set_error("fgmres: residual is not finite");
              return STFEM_ERR_NO_CONVERGENCE;
            }
          if (first)
            {
              res.initial_residual = beta;
              tol                  = std::max(abstol, reduce * beta);
              first                = false;
            }
          // also after a restart: the recomputed residual may already meet the tolerance (or be exactly 0)
          res.final_residual = beta;
          if (beta <= tol)
            {
              res.converged = true;
              break;
            }
          v_scale(V[0], 1.0 / beta);
          std::fill(H.begin(), H.end(), 0.0);
          std::fill(g.begin(), g.end(), 0.0);
          g[0]       = beta;
          int  jdone = 0;
          bool stop  = false;
          for (int j = 0; j < max_basis; ++j)
            {
              STFEM_FORWARD(ensure(Z, j, nb, n));
              STFEM_FORWARD(ensure(V, j + 1, nb, n));
              if (M)
                STFEM_FORWARD(M->vmult(Z[j].block_ptrs(), V[j].cblock_ptrs()));
              else
                STFEM_FORWARD(v_copy(Z[j], V[j]));
              STFEM_FORWARD(apply_A(w, Z[j]));
              // Classical Gram-Schmidt with one re-orthogonalisation ("twice is enough"), in three passes over the basis instead
              // of five: (1) fused multi-dot: all h_ij; (2) w -= V h fused with the inner products of the UPDATED w with V
              // and with itself; (3) the second update, the normalisation and the copy into the basis in one pass - the norm
              // of the final vector follows from Pythagoras, ||w' - V h'||^2 = ||w'||^2 - ||h'||^2, which is safe here: after
              // the first pass h' is at rounding level, so nothing cancels.
              std::vector<const BlockVec<double> *> basis;
              for (int i = 0; i <= j; ++i) basis.push_back(&V[i]);
              double              hn = 0;
              std::vector<double> c(j + 1);
              STFEM_FORWARD(v_multi_dot(sc, w, basis, h.data()));
              for (int i = 0; i <= j; ++i)
                {
                  c[i] = -h[i];
                  Hm(i, j) += h[i];
                }
              if (j + 1 <= MAXK)
                {
                  STFEM_FORWARD(v_multi_axpy_dot(sc, w, basis, c.data(), h.data()));
                  double hh = 0;
                  for (int i = 0; i <= j; ++i)
                    {
                      c[i] = -h[i];
                      Hm(i, j) += h[i];
                      hh += h[i] * h[i];
                    }
                  hn = std::sqrt(std::max(h[j + 1] - hh, 0.0));
                  if (hn != 0.0) v_multi_axpy_scale_out(V[j + 1], w, basis, c.data(), 1.0 / hn);
                }
              else
                {
                  double wnorm2 = 0;
                  v_multi_axpy(w, basis, c.data());
                  STFEM_FORWARD(v_multi_dot(sc, w, basis, h.data()));
                  for (int i = 0; i <= j; ++i)
                    {
                      c[i] = -h[i];
                      Hm(i, j) += h[i];
                    }
                  STFEM_FORWARD(v_multi_axpy_norm(sc, w, basis, c.data(), &wnorm2));
                  hn = std::sqrt(std::max(wnorm2, 0.0));
                  if (hn != 0.0) v_scale_copy(V[j + 1], 1.0 / hn, w);
                }
              Hm(j + 1, j) = hn;
              for (int i = 0; i < j; ++i)
                {
                  const double t = cs[i] * Hm(i, j) + sn[i] * Hm(i + 1, j);
                  Hm(i + 1, j)   = -sn[i] * Hm(i, j) + cs[i] * Hm(i + 1, j);
                  Hm(i, j)       = t;
                }
              const double den = std::hypot(Hm(j, j), Hm(j + 1, j));
              cs[j] = Hm(j, j) / den;
              sn[j] = Hm(j + 1, j) / den;
              Hm(j, j)     = den;
              Hm(j + 1, j) = 0.0;
              g[j + 1]     = -sn[j] * g[j];
              g[j]         = cs[j] * g[j];
              res.iterations++;
              jdone              = j + 1;
              res.final_residual = std::fabs(g[j + 1]);
              if (res.final_residual <= tol)
                {
                  res.converged = true;
                  stop          = true;
                }
              if (res.iterations >= max_iter) stop = true;
              if (stop) break;
            }
          // back substitution and solution update x += sum_j y_j z_j (one fused multi-axpy)
          std::vector<double> y(jdone);
          for (int i = jdone - 1; i >= 0; --i)
            {
              double s = g[i];
              for (int k = i + 1; k < jdone; ++k) s -= Hm(i, k) * y[k];
              y[i] = s / Hm(i, i);
            }
          std::vector<const BlockVec<double> *> zs;
          for (int i = 0; i < jdone; ++i) zs.push_back(&Z[i]);
          v_multi_axpy(x, zs, y.data());
          if (res.converged || res.iterations >= max_iter) break;
        }
      for (int k = 0; k < nb; ++k)
        STFEM_CUDA_CHECK(cudaMemcpyAsync(x_blocks[k], x.d + (size_t)k * n, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
      STFEM_FORWARD(stream_sync_checked(ctx, "fgmres"));
      if (!res.converged)
        {
          set_error("fgmres: no convergence after %d iterations (residual %.3e, tolerance %.3e)", res.iterations, res.final_residual, tol);
          return STFEM_ERR_NO_CONVERGENCE;
        }
      return STFEM_OK;
    }
  };
} // namespace stfem

// handles of the C ABI
struct stfem_mg
{
  std::unique_ptr<stfem::MGBase> impl;
  int                            number_type = STFEM_F32;
};

struct stfem_solver
{
  stfem::Fgmres       fgmres;
  stfem::FgmresResult last;
};
