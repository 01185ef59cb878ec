// stfem_b200.hpp — header-only C++17 façade over the C ABI (stfem_b200.h).
//
// Mirrors the duck-typed interface deal.II's solver / multigrid templates instantiate in the reference
// (SURVEY.md §8b), same member names and argument meaning:
//   stfem::BlockVector<Number>   ~ BlockVectorT<Number>                 (reference include/types.h:20-23)
//   stfem::SystemMatrix<Number>  ~ SystemMatrix<dim,Number,...>         (include/operators.h:516-663)
//        vmult(dst, src), Tvmult(dst, src), vmult_slice_add(dst, src), vmult_slice(dst, src),
//        initialize_dof_vector(vec), m(), n(), get_matrix_diagonal()
//   stfem::GMG                   ~ GMG<dim,Number,LevelMatrixType>      (include/stmg.h:1047-1344)  vmult(dst, src)
//   stfem::SolverFGMRES          ~ dealii::SolverFGMRES + ReductionControl  (include/time_integrators.h:56-59)
//        solve(matrix, x, rhs, preconditioner)
// Error behaviour: the reference aborts through Assert/AssertThrow (include/time_integrators.h:317-320); here every
// failing C call throws stfem::Error carrying stfem_last_error().  Operators keep references to the mesh like the
// reference keeps `const &` to K, M, Alpha, Beta (include/operators.h:465-469): the caller keeps them alive.
#ifndef STFEM_B200_HPP
#define STFEM_B200_HPP

#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "stfem_b200.h"

namespace stfem
{
  struct Error : std::runtime_error
  {
    int code;
    Error(int c, const std::string &what) : std::runtime_error(what), code(c) {}
  };

  inline void check(int rc)
  {
    if (rc != STFEM_OK) throw Error(rc, std::string("stfem error ") + std::to_string(rc) + ": " + stfem_last_error());
  }

  template <typename Number>
  constexpr int number_type()
  {
    static_assert(std::is_same<Number, double>::value || std::is_same<Number, float>::value, "double or float");
    return std::is_same<Number, double>::value ? STFEM_F64 : STFEM_F32;
  }

  // one per GPU / rank (the reference: one MPI rank)
  class Context
  {
  public:
    explicit Context(int device = 0) { check(stfem_ctx_create(device, &h_)); }
    ~Context() { stfem_ctx_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    stfem_ctx_t handle() const { return h_; }
    void        synchronize() const { check(stfem_ctx_synchronize(h_)); }
    long long   launch_count() const { return stfem_ctx_launch_count(h_); }

  private:
    stfem_ctx_t h_ = nullptr;
  };

  // GridGenerator::subdivided_hyper_rectangle + refine_global (+ distort_random): tests/tp_01.cc:83-90
  class Mesh
  {
  public:
    Mesh(const Context &ctx, const std::vector<int> &n_cells, const std::vector<double> &lower = {},
         const std::vector<double> &upper = {}, const double *vertices = nullptr, unsigned dirichlet_faces = ~0u)
      : ctx_(ctx), dim_((int)n_cells.size())
    {
      std::vector<double> lo = lower.empty() ? std::vector<double>(dim_, 0.0) : lower;
      std::vector<double> up = upper.empty() ? std::vector<double>(dim_, 1.0) : upper;
      const unsigned      all = dim_ == 3 ? 0x3fu : 0xfu;
      check(stfem_mesh_create(ctx.handle(), dim_, n_cells.data(), lo.data(), up.data(), vertices,
                              dirichlet_faces == ~0u ? all : dirichlet_faces, &h_));
    }
    ~Mesh() { stfem_mesh_destroy(h_); }
    Mesh(const Mesh &) = delete;
    Mesh &operator=(const Mesh &) = delete;
    stfem_mesh_t   handle() const { return h_; }
    const Context &context() const { return ctx_; }
    int            dim() const { return dim_; }

  private:
    const Context &ctx_;
    int            dim_;
    stfem_mesh_t   h_ = nullptr;
  };

  // nb separate device arrays of N numbers
  template <typename Number>
  class BlockVector
  {
  public:
    BlockVector() = default;
    BlockVector(const Context &ctx, unsigned n_blocks, long long block_size) { reinit(ctx, n_blocks, block_size); }
    ~BlockVector() { clear(); }
    BlockVector(const BlockVector &) = delete;
    BlockVector &operator=(const BlockVector &) = delete;
    BlockVector(BlockVector &&o) noexcept { swap(o); }
    BlockVector &operator=(BlockVector &&o) noexcept
    {
      swap(o);
      return *this;
    }
    void swap(BlockVector &o)
    {
      std::swap(ctx_, o.ctx_);
      std::swap(n_, o.n_);
      ptrs_.swap(o.ptrs_);
    }
    void reinit(const Context &ctx, unsigned n_blocks, long long block_size)
    {
      clear();
      ctx_ = &ctx;
      n_   = block_size;
      ptrs_.assign(n_blocks, nullptr);
      for (auto &p : ptrs_) check(stfem_dev_alloc(ctx.handle(), sizeof(Number) * (size_t)block_size, &p));
      *this = Number(0);
    }
    void clear()
    {
      for (void *p : ptrs_)
        if (p) stfem_dev_free(ctx_->handle(), p);
      ptrs_.clear();
    }
    unsigned  n_blocks() const { return (unsigned)ptrs_.size(); }
    long long block_size() const { return n_; }
    long long size() const { return n_ * (long long)ptrs_.size(); }
    // only assignment of zero is provided (dealii: `vec = 0`)
    BlockVector &operator=(Number zero)
    {
      if (zero != Number(0)) throw Error(STFEM_ERR_INVALID, "BlockVector: only = 0 is supported");
      for (void *p : ptrs_) check(stfem_dev_memset(ctx_->handle(), p, 0, sizeof(Number) * (size_t)n_));
      return *this;
    }
    void  copy_from_host(unsigned b, const Number *src) { check(stfem_dev_upload(ctx_->handle(), ptrs_[b], src, sizeof(Number) * (size_t)n_)); }
    void  copy_to_host(unsigned b, Number *dst) const { check(stfem_dev_download(ctx_->handle(), dst, ptrs_[b], sizeof(Number) * (size_t)n_)); }
    void *block(unsigned b) const { return ptrs_[b]; }
    void *const       *data() { return ptrs_.data(); }
    const void *const *data() const { return const_cast<const void *const *>(ptrs_.data()); }

  private:
    const Context      *ctx_ = nullptr;
    long long           n_   = 0;
    std::vector<void *> ptrs_;
  };

  // dealii::DiagonalMatrix<BlockVectorType> as returned by get_matrix_diagonal: access to the vector of entries
  template <typename Number>
  class DiagonalMatrix
  {
  public:
    BlockVector<Number>       &get_vector() { return v_; }
    const BlockVector<Number> &get_vector() const { return v_; }
    long long                  m() const { return v_.size(); }

  private:
    BlockVector<Number> v_;
  };

  // A = Alpha (x) K + Beta (x) M  with K, M the matrix-free Laplace / mass operators of FE_Q(degree)
  template <typename Number>
  class SystemMatrix
  {
  public:
    using BlockVectorType = BlockVector<Number>;
    // Alpha, Beta: row-major nb_rows x nb_cols (FullMatrix layout)
    SystemMatrix(const Mesh &mesh, int degree, int nb_rows, int nb_cols, const double *Alpha, const double *Beta,
                 const double *laplace_coeff_cell = nullptr, const double *laplace_coeff_q = nullptr)
      : mesh_(mesh)
    {
      stfem_op_desc d{};
      d.degree             = degree;
      d.number_type        = number_type<Number>();
      d.nb_rows            = nb_rows;
      d.nb_cols            = nb_cols;
      d.Alpha              = Alpha;
      d.Beta               = Beta;
      d.laplace_coeff_cell = laplace_coeff_cell;
      d.laplace_coeff_q    = laplace_coeff_q;
      check(stfem_op_create(mesh.handle(), &d, &h_));
    }
    ~SystemMatrix() { stfem_op_destroy(h_); }
    SystemMatrix(const SystemMatrix &) = delete;
    SystemMatrix &operator=(const SystemMatrix &) = delete;

    void vmult(BlockVectorType &dst, const BlockVectorType &src) const { check(stfem_op_vmult(h_, dst.data(), src.data(), 0)); }
    void Tvmult(BlockVectorType &dst, const BlockVectorType &src) const { check(stfem_op_vmult(h_, dst.data(), src.data(), 1)); }
    // dst_j += Alpha(j,0) K src_0 + Beta(j,0) M src_0   (operators.h:586-611)
    void vmult_slice_add(BlockVectorType &dst, const BlockVectorType &src) const { check(stfem_op_vmult_slice_add(h_, dst.data(), src.block(0))); }
    void vmult_slice(BlockVectorType &dst, const BlockVectorType &src) const
    {
      dst = Number(0);
      vmult_slice_add(dst, src);
    }
    void initialize_dof_vector(BlockVectorType &vec) const { vec.reinit(mesh_.context(), (unsigned)stfem_op_n_blocks(h_), m()); }
    void initialize_dof_vector(BlockVectorType &vec, unsigned n_blocks) const { vec.reinit(mesh_.context(), n_blocks, m()); }
    long long  m() const { return stfem_op_n_dofs_per_block(h_); }
    long long  n() const { return m(); }
    // diag_i = Alpha(i,i) diag K + Beta(i,i) diag M   (operators.h:613-625)
    std::shared_ptr<DiagonalMatrix<Number>> get_matrix_diagonal() const
    {
      auto d = std::make_shared<DiagonalMatrix<Number>>();
      initialize_dof_vector(d->get_vector());
      check(stfem_op_diagonal(h_, d->get_vector().data()));
      return d;
    }
    stfem_op_t handle() const { return h_; }

  private:
    const Mesh &mesh_;
    stfem_op_t  h_ = nullptr;
  };

  // PreconditionerGMGAdditionalData (include/parameters.h:12-31)
  struct PreconditionerGMGAdditionalData
  {
    double       smoothing_range                  = 1;
    unsigned int smoothing_steps                  = 1;
    unsigned int smoothing_eig_cg_n_iterations    = 20;
    double       relaxation                       = 0.0;
    bool         restrict_is_transpose_prolongate = true;
    bool         variable                         = true;
    std::string  coarse_grid_smoother_type        = "Smoother"; // anything else: GMRES on the coarsest level (stmg.h:1240-1302)
    unsigned int coarse_grid_maxiter              = 10;
    double       coarse_grid_abstol               = 1e-20;
    int          inner_preconditioner             = 0; // 0 PreconditionVanka (reference), 1 point-Jacobi (not a reference option)
    int          vanka_storage                    = 0; // 0 level precision (reference), 1 FP16 patch inverses
  };

  // Space-time multigrid preconditioner; level matrices coarse -> fine, all of one precision
  template <typename Number>
  class GMG
  {
  public:
    GMG(const Context &ctx, const std::vector<const SystemMatrix<Number> *> &level_matrices, const std::string &mg_type_level,
        const std::vector<int> &smoother_types, int time_type, int n_timesteps_at_once, const std::vector<int> &poly_time_sequence,
        const PreconditionerGMGAdditionalData &data = PreconditionerGMGAdditionalData())
    {
      std::vector<stfem_op_t> ops;
      for (auto *m : level_matrices) ops.push_back(m->handle());
      stfem_mg_desc d{};
      d.n_levels                         = (int)ops.size();
      d.level_ops                        = ops.data();
      d.mg_type_level                    = mg_type_level.c_str();
      d.smoother_types                   = smoother_types.data();
      d.time_type                        = time_type;
      d.n_timesteps_at_once              = n_timesteps_at_once;
      d.poly_time_sequence               = poly_time_sequence.data();
      d.n_poly_time                      = (int)poly_time_sequence.size();
      d.smoothing_steps                  = (int)data.smoothing_steps;
      d.relaxation                       = data.relaxation;
      d.smoothing_range                  = data.smoothing_range;
      d.eig_n_iterations                 = (int)data.smoothing_eig_cg_n_iterations;
      d.variable                         = data.variable;
      d.restrict_is_transpose_prolongate = data.restrict_is_transpose_prolongate;
      d.inner_preconditioner             = data.inner_preconditioner;
      d.vanka_storage                    = data.vanka_storage;
      d.coarse_grid_maxiter              = data.coarse_grid_smoother_type == "Smoother" ? 0 : (int)data.coarse_grid_maxiter;
      d.coarse_grid_abstol               = data.coarse_grid_abstol;
      check(stfem_mg_create(ctx.handle(), &d, &h_));
    }
    ~GMG() { stfem_mg_destroy(h_); }
    GMG(const GMG &) = delete;
    GMG &operator=(const GMG &) = delete;
    // one V-cycle; dst, src in double like the outer solver's vectors (stmg.h:1331-1344)
    void       vmult(BlockVector<double> &dst, const BlockVector<double> &src) const { check(stfem_mg_vmult(h_, dst.data(), src.data())); }
    stfem_mg_t handle() const { return h_; }

  private:
    stfem_mg_t h_ = nullptr;
  };

  struct PreconditionIdentity
  {
  };

  // ReductionControl(max_steps, abs_tol, reduce) + SolverFGMRES::AdditionalData(max_basis_size)
  class SolverFGMRES
  {
  public:
    SolverFGMRES(unsigned max_steps = 200, double abs_tol = 1e-12, double reduce = 1e-12, unsigned max_basis_size = 100)
      : max_steps_(max_steps), abs_tol_(abs_tol), reduce_(reduce), max_basis_(max_basis_size)
    {
      check(stfem_solver_create(&h_));
    }
    ~SolverFGMRES() { stfem_solver_destroy(h_); }
    SolverFGMRES(const SolverFGMRES &) = delete;
    SolverFGMRES &operator=(const SolverFGMRES &) = delete;

    template <typename Number>
    void solve(const SystemMatrix<double> &A, BlockVector<double> &x, const BlockVector<double> &b, const GMG<Number> &preconditioner)
    {
      run(A.handle(), preconditioner.handle(), x, b);
    }
    void solve(const SystemMatrix<double> &A, BlockVector<double> &x, const BlockVector<double> &b, const PreconditionIdentity &)
    {
      run(A.handle(), nullptr, x, b);
    }
    unsigned last_step() const { return (unsigned)iterations_; }
    double   initial_value() const { return r0_; }
    double   last_value() const { return r1_; }

  private:
    void run(stfem_op_t A, stfem_mg_t M, BlockVector<double> &x, const BlockVector<double> &b)
    {
      // throws on SolverControl::NoConvergence like the reference's AssertThrow (time_integrators.h:317-320)
      check(stfem_fgmres_solve(h_, A, M, x.data(), b.data(), (int)max_basis_, (int)max_steps_, abs_tol_, reduce_, &iterations_, &r0_, &r1_));
    }
    stfem_solver_t h_ = nullptr;
    unsigned       max_steps_;
    double         abs_tol_, reduce_;
    unsigned       max_basis_;
    int            iterations_ = 0;
    double         r0_ = 0, r1_ = 0;
  };
} // namespace stfem

#endif /* STFEM_B200_HPP */
