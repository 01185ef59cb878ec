/* stfem_b200.h — C ABI of the B200-native space-time FEM hot path.
 *
 * The reference (immaaane/dealii-stfem) has no FFI: its boundary is the duck-typed C++
 * concept deal.II's solvers instantiate.  This header is the thin C layer
 * under the C++ façade (dealii-stfem_b200/include/stfem/*.h) that mirrors those classes;
 * every entry point names the reference interface it replaces.
 *
 * Conventions: opaque handles, int return codes (0 = STFEM_OK), no exceptions cross the
 * boundary, stfem_last_error() gives the message of the last failing call on this thread.
 * Block vectors keep the reference layout (include/types.h:20-23 of the reference): nb
 * separate contiguous arrays of N spatial DoFs, lexicographic DoF numbering of the
 * structured hex mesh (x fastest).  `number_type`: 0 = double, 1 = float.
 * Device pointers are plain `void *` / `double *`; no torch types anywhere.
 */
#ifndef STFEM_B200_H
#define STFEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STFEM_OK 0
#define STFEM_ERR_INVALID 1
#define STFEM_ERR_CUDA 2
#define STFEM_ERR_UNSUPPORTED 3
#define STFEM_ERR_NO_CONVERGENCE 4

#define STFEM_F64 0
#define STFEM_F32 1

#define STFEM_MAX_BLOCKS 16

typedef struct stfem_ctx *stfem_ctx_t;
typedef struct stfem_mesh *stfem_mesh_t;
typedef struct stfem_op *stfem_op_t;

const char *stfem_last_error(void);
const char *stfem_version(void);

/* ---- context: one per GPU / per rank (reference: one MPI rank, operators.h:28-40) ---- */
int stfem_ctx_create(int device, stfem_ctx_t *out);
int stfem_ctx_destroy(stfem_ctx_t ctx);
int stfem_ctx_synchronize(stfem_ctx_t ctx);
/* the cudaStream_t all work of this context is enqueued on (as void*) */
void *stfem_ctx_stream(stfem_ctx_t ctx);
/* number of kernels this library launched on the context so far (bench "gpu_launches") */
long long stfem_ctx_launch_count(stfem_ctx_t ctx);
/* CUDA-event stopwatch on the context stream: start records an event, stop records a second one,
 * synchronises on it and returns the elapsed device time in ms */
int stfem_ctx_timer_start(stfem_ctx_t ctx);
int stfem_ctx_timer_stop(stfem_ctx_t ctx, float *ms);

/* ---- device memory helpers (plain cudaMalloc/cudaMemcpyAsync on the context stream) ---- */
int stfem_dev_alloc(stfem_ctx_t ctx, size_t bytes, void **out);
int stfem_dev_free(stfem_ctx_t ctx, void *p);
int stfem_dev_upload(stfem_ctx_t ctx, void *dst_dev, const void *src_host, size_t bytes);
int stfem_dev_download(stfem_ctx_t ctx, void *dst_host, const void *src_dev, size_t bytes);
int stfem_dev_memset(stfem_ctx_t ctx, void *dst_dev, int value, size_t bytes);
int stfem_dev_copy(stfem_ctx_t ctx, void *dst_dev, const void *src_dev, size_t bytes);
int stfem_host_alloc_pinned(size_t bytes, void **out);
int stfem_host_free_pinned(void *p);

/* ---- mesh: GridGenerator::subdivided_hyper_rectangle + refine_global (+ distort_random),
 *      reference tests/tp_01.cc:83-90.  n_cells[d] cells per direction after refinement.
 *      vertices: NULL for a Cartesian box, else (n_cells[0]+1)*(n_cells[1]+1)*... points of
 *      `dim` doubles, lexicographic (x fastest) — MappingQ1 geometry.
 *      dirichlet_faces: bit (2*d+side) set = homogeneous Dirichlet on that face
 *      (DoFTools::make_zero_boundary_constraints, tp_01.cc:99); 0x3f / 0xf = all faces. ---- */
int stfem_mesh_create(stfem_ctx_t ctx, int dim, const int *n_cells, const double *lower,
                      const double *upper, const double *vertices, unsigned dirichlet_faces,
                      stfem_mesh_t *out);
int stfem_mesh_destroy(stfem_mesh_t mesh);

/* ---- space-time operator  A = Alpha (x) K + Beta (x) M
 *      replaces SystemMatrix<dim,Number,MatrixFreeOperatorScalar> (reference
 *      include/operators.h:516-663) together with its two MatrixFreeOperator members
 *      K (laplace_scaling=1) and M (mass_scaling=1) (operators.h:967-1191, tp_01.cc:114-119).
 *      Alpha, Beta: row-major nb_rows x nb_cols doubles (FullMatrix layout).  nb_cols == nb_rows
 *      for the system matrix; nb_cols == 1 for the "slice" right-hand-side matrices
 *      (operators.h:586-611).  laplace_coeff_cell: NULL or one coefficient per cell
 *      (Coefficient<dim>, operators.h:870-965, is piecewise constant per coarse cell). ---- */
typedef struct stfem_op_desc {
  int degree;      /* FE_Q(degree) in space, QGauss(degree+1) */
  int number_type; /* STFEM_F64 | STFEM_F32 */
  int nb_rows;
  int nb_cols;
  const double *Alpha;
  const double *Beta;
  const double *laplace_coeff_cell; /* host pointer, n_cells values, may be NULL */
  const double *laplace_coeff_q;    /* host pointer, n_cells * (degree+1)^dim values (cell-major, q-points
                                       lexicographic) = the reference's Table [cell][q]
                                       (operators.h:1060-1087, 1185-1186); may be NULL */
  int kernel_variant;               /* 0 = default: the brick kernel (st_vmult_brick.cuh) on 3D Cartesian constant-coefficient
                                       meshes with >= 2 time blocks, the per-cell kernels elsewhere; general geometry: stored
                                       metric where it fits, else geometry on the fly.  Tuning / cross-check selectors (all
                                       produce the same operator): 1 generic q-point kernel; 2 (level operators) dense Vanka
                                       patches instead of the Kronecker form; 3 per-cell Cartesian kernel of round 1 instead of
                                       the brick kernel; 5 / 6 general geometry with the stored metric / on the fly from the
                                       cell vertices; 11-19, 26-28 launch-bound configurations of the per-cell Cartesian kernel;
                                       17 largest-CTA rule; 18, 20-25 software-pipelined persistent kernel; 31-34 ablation
                                       experiments (WRONG results, timing only; rejected unless STFEM_ALLOW_ABLATION is set);
                                       40-42 cp.async.bulk + mbarrier gather; 51 two 8-byte exchange fields instead of 16-byte
                                       pairs; 60 fast-diagonalisation form of the per-cell kernel (st_vmult_cart_fd.cuh,
                                       verified on the GPU in round 2: tests/test_cart_fd_gpu.py); brick kernel: 70 plain loads
                                       instead of TMA, 72 no size threshold, 77 X and Y+Z phases on separate warps, 79 barrier
                                       pipeline instead of CTA barriers, 80-89 forced number of z chunks, 90 per-SM alternation
                                       of the X warps */
} stfem_op_desc;

int stfem_op_create(stfem_mesh_t mesh, const stfem_op_desc *desc, stfem_op_t *out);
int stfem_op_destroy(stfem_op_t op);
/* m(): operators.h:642-646 */
long long stfem_op_n_dofs_per_block(stfem_op_t op);
int stfem_op_n_blocks(stfem_op_t op);

/* vmult / Tvmult (operators.h:536-583): dst = A src (transpose != 0: Alpha^T, Beta^T).
 * dst, src: arrays of nb device pointers.  Constrained rows of dst are 0 (App. A.3). */
int stfem_op_vmult(stfem_op_t op, void *const *dst, const void *const *src, int transpose);
/* vmult_slice_add (operators.h:586-611): dst_j += Alpha(j,0) K src0 + Beta(j,0) M src0 */
int stfem_op_vmult_slice_add(stfem_op_t op, void *const *dst, const void *src0);
/* get_matrix_diagonal (operators.h:613-625): diag_i = Alpha(i,i) diag K + Beta(i,i) diag M */
int stfem_op_diagonal(stfem_op_t op, void *const *diag);

/* kernel_variant 60: the Cartesian operator in fast-diagonalisation form (csrc/st_vmult_cart_fd.cuh; parity-tested on the GPU,
 * not the default: 1.01 ms against 0.90 ms of the brick kernel on configs[1]).  This host-only helper returns the modes it is
 * built on: V (n1 x n1, row-major) and lam (n1) with  Mh = V^T V,  Kh = V^T diag(lam) V  for the 1D reference matrices of
 * FE_Q(degree) / QGauss(degree+1). */
int stfem_cart_fd_modes(int degree, double *V, double *lam);

/* Same call with HOST buffers (nb arrays of N numbers of the operator's number type):
 * copies src to the device, applies, copies dst back — the end-to-end path bench.py times. */
int stfem_op_vmult_host(stfem_op_t op, void *const *dst_host, const void *const *src_host,
                        int transpose);

/* Measurement helper: the copy floor of stfem_op_vmult_host - the operator's blocks uploaded and downloaded concurrently
 * on two copy streams without any kernel; host wall time per round in milliseconds (bench.py: e2e.copy_floor_ms). */
int stfem_op_host_copy_floor(stfem_op_t op, void *const *dst_host, const void *const *src_host, int reps, double *ms_per_round);

/* time of the last vmult kernel sequence on this operator, measured with CUDA events on the
 * context stream (ms); enabled with stfem_op_set_timing(op, 1) */
int stfem_op_set_timing(stfem_op_t op, int enable);
float stfem_op_last_kernel_ms(stfem_op_t op);

/* ---- multi-GPU: one process per GPU, box partition of the structured mesh; replaces the MPI domain
 *      decomposition deal.II provides to the reference (ghost update + compress(add) inside
 *      MatrixFree::cell_loop called at include/operators.h:1016, MPI_Allreduce of dot products).
 *      NCCL is opened at run time (dlopen): rank 0 calls stfem_comm_unique_id, the 128 bytes are
 *      distributed by the launcher (torch.distributed / MPI / files), every rank calls
 *      stfem_ctx_comm_init.  A mesh created from stfem_partition_brick's local extents and marked with
 *      stfem_mesh_set_partition makes every operator built on it exchange its interface DoFs. ---- */
int stfem_comm_unique_id(char *id128);
int stfem_ctx_comm_init(stfem_ctx_t ctx, int rank, int n_ranks, const char *id128);
int stfem_ctx_comm_destroy(stfem_ctx_t ctx);
int stfem_ctx_rank(stfem_ctx_t ctx);
/* host values reduced in place over the ranks of the context's communicator (the MPI_Allreduce / Utilities::MPI::sum|max
 * of the reference's drivers, e.g. the error norms of tests/tp_01.cc:409-432); op: 0 sum, 1 max, 2 min; n <= 256.
 * Without a communicator the values are returned unchanged. */
int stfem_ctx_allreduce(stfem_ctx_t ctx, double *values, int n, int op);
int stfem_ctx_n_ranks(stfem_ctx_t ctx);
/* brick of process `coords` (x fastest rank order) in `proc_grid`: local cells, global cell offset,
 * local bounding box, Dirichlet mask restricted to physical boundary faces.  Pure host logic. */
int stfem_partition_brick(int dim, const int *n_global, const double *lower, const double *upper, const int *proc_grid,
                          const int *coords, int *n_local, int *cell_offset, double *local_lower, double *local_upper,
                          unsigned *dirichlet_faces);
int stfem_mesh_set_partition(stfem_mesh_t mesh, const int *proc_grid, const int *coords);
/* Ghost layer of a partitioned mesh = the local brick extended by ONE cell layer across every face shared with another
 * rank (deal.II's ghost cells; the reference gets them from parallel::distributed::Triangulation, tests/tp_01.cc:83-90).
 * Needed only by the dense cell-patch smoother (PreconditionVanka, include/stmg.h:745-872: its patch matrices take the
 * contributions of the neighbouring cells on shared DoFs, include/compute_block_matrix.h:50-139).  The extended brick has
 * n_d + g_lo + g_hi cells per direction (g = 1 where stfem_mesh_set_partition found a neighbour), x fastest.
 *   vertices_ext   (n_ext_d + 1) points per direction, dim coordinates each: general (MappingQ1) meshes only
 *   coeff_cell_ext / coeff_q_ext   the operator's laplace_coeff_cell / laplace_coeff_q over the extended brick (NULL: none) */
int stfem_mesh_set_ghost_vertices(stfem_mesh_t mesh, const double *vertices_ext);
int stfem_op_set_ghost_coefficients(stfem_op_t op, const double *coeff_cell_ext, const double *coeff_q_ext);

/* host-only test hook: the interface exchange of a box partition among all its bricks inside one process, executing the
 * same element functions as the pack / unpack kernels (no GPU, no NCCL).  data: [n_ranks][nb][np0*np1*np2] doubles. */
int stfem_halo_emulate_host(int dim, const int *proc_grid, const int *np, int nb, double *data);
/* add the partial values of interface DoFs over the ranks sharing them (compress(add) + ghost update);
 * stfem_op_vmult & co. call this internally on partitioned meshes */
int stfem_op_halo_add(stfem_op_t op, void *const *blocks, int nb);

/* ---- space-time multigrid preconditioner: GMG<dim, Number, LevelMatrixType> (reference
 *      include/stmg.h:1047-1344) with PreconditionVanka smoothers (stmg.h:745-872), the transfers of
 *      build_stmg_transfers (stmg.h:538-617) and deal.II's Multigrid V-cycle / MGSmootherPrecondition /
 *      PreconditionRelaxation|Chebyshev.  Levels are ordered coarse -> fine like
 *      MGLevelObject; level operators are stfem_op_t objects of ONE precision (float for
 *      test<double,float>, tests/tp_01.cc:801-804) built by the caller on the level meshes
 *      (tests/tp_01.cc:250-321).  Parameter names follow PreconditionerGMGAdditionalData
 *      (include/parameters.h:12-31). ---- */
typedef struct stfem_mg *stfem_mg_t;
typedef struct stfem_solver *stfem_solver_t;

typedef struct stfem_mg_desc {
  int n_levels;
  const stfem_op_t *level_ops;   /* n_levels, coarse -> fine */
  const char *mg_type_level;     /* n_levels-1 chars of 'h','p','k','t' (get_mg_sequence) */
  const int *smoother_types;     /* n_levels: 0 Identity, 1 Relaxation, 2 Chebyshev (get_precondition_stmg_types) */
  int time_type;                 /* 1 CGP, 2 DG */
  int n_timesteps_at_once;       /* on the finest level */
  const int *poly_time_sequence; /* time degrees coarse -> fine (get_poly_mg_sequence) */
  int n_poly_time;
  int smoothing_steps;           /* default 1 */
  double relaxation;             /* 0 = estimate by power iteration */
  double smoothing_range;        /* default 1 */
  int eig_n_iterations;          /* default 20 */
  int variable;                  /* MGSmootherPrecondition variable smoothing, default 1 */
  int restrict_is_transpose_prolongate; /* default 1 */
  int inner_preconditioner;      /* 0 = PreconditionVanka (the reference, stmg.h:1055-1063); 1 = point-Jacobi: the inverse of
                                    SystemMatrix::get_matrix_diagonal (operators.h:613-625) inside Relaxation / Chebyshev.
                                    Not a reference configuration: the cheap smoother BASELINE.json's north_star names. */
  int coarse_grid_maxiter;       /* 0 = coarsest level solved by the smoother (MGCoarseGridApplySmoother, coarseGridSmootherType
                                    "Smoother", the reference's default); > 0 = MGCoarseGridIterativeSolver: left-preconditioned
                                    SolverGMRES with IterationNumberControl(coarse_grid_maxiter, coarse_grid_abstol) and the
                                    coarse smoother as preconditioner (stmg.h:1240-1302; parameters.h:25-26: 10, 1e-20) */
  double coarse_grid_abstol;
  int vanka_storage;             /* 0 = patch inverses in the level precision (the reference: float, stmg.h:901);
                                    1 = FP16, normalised per patch, FP32 accumulation: half the HBM traffic of the dense
                                    PreconditionVanka::vmult (stmg.h:832-872).  Levels in Kronecker form are unaffected. */
} stfem_mg_desc;

int stfem_mg_create(stfem_ctx_t ctx, const stfem_mg_desc *desc, stfem_mg_t *out);
int stfem_mg_destroy(stfem_mg_t mg);
int stfem_mg_n_levels(stfem_mg_t mg);
/* GMG::vmult (stmg.h:1331-1344): one V-cycle; dst, src = nb DOUBLE device arrays of the finest level */
int stfem_mg_vmult(stfem_mg_t mg, void *const *dst, const void *const *src);
/* Single pieces of the cycle, for parity tests; vectors in the LEVEL precision.
 * what: 0 PreconditionVanka::vmult, 1 PreconditionSTMG::vmult (smoother), 2 restrict level -> level-1,
 *       3 prolongate level-1 -> level, 4 level operator vmult, 5 V-cycle started on this level */
int stfem_mg_level_apply(stfem_mg_t mg, int level, int what, void *const *dst, const void *const *src);
/* out10: smoother type, lambda estimate, omega, theta, delta, smoothing steps, N, blocks,
 *        stored patch matrices, bytes of patch storage */
int stfem_mg_level_info(stfem_mg_t mg, int level, double *out10);

/* SolverFGMRES(ReductionControl(max_iterations, abs_tol, reduce), AdditionalData(max_basis_size))
 * (include/time_integrators.h:56-59): right-preconditioned flexible GMRES, x is the initial guess on
 * entry.  M may be NULL (unpreconditioned).  Returns STFEM_ERR_NO_CONVERGENCE like the reference's
 * AssertThrow on SolverControl::NoConvergence (time_integrators.h:317-320). */
int stfem_solver_create(stfem_solver_t *out);
int stfem_solver_destroy(stfem_solver_t s);
int stfem_fgmres_solve(stfem_solver_t s, stfem_op_t A, stfem_mg_t M, void *const *x, const void *const *b,
                       int max_basis_size, int max_iterations, double abs_tol, double reduce, int *iterations,
                       double *initial_residual, double *final_residual);

/* ---- time stepping: TimeIntegratorFO / TimeIntegratorWave (reference include/time_integrators.h:24-459).
 *      The right-hand-side function is one of the reference's analytic functions
 *      (include/exact_solution.h:27-197): 0 zero, 1 ExactSolution, 2 RHSFunction (heat),
 *      3 wave::ExactSolutionV, 4 wave::RHSFunction; `frequency` is their parameter f. ---- */
typedef struct stfem_time_integrator *stfem_ti_t;

typedef struct stfem_ti_desc {
  int time_type;            /* 1 CGP, 2 DG */
  int time_degree;          /* r */
  int n_timesteps_at_once;
  int problem;              /* 1 heat (TimeIntegratorFO), 2 wave (TimeIntegratorWave) */
  const double *Alpha_1;    /* single-step weights of get_fe_time_weights(type, r, tau, 1): nd x nd */
  const double *Beta_1;     /* nd x nd (wave only) */
  const double *Gamma_1;    /* nd x 1 */
  const double *Zeta_1;     /* nd x 1 (wave only) */
  stfem_op_t matrix;        /* system matrix, double precision */
  stfem_mg_t preconditioner; /* may be NULL */
  stfem_op_t rhs_matrix;    /* nb x 1 slice operator (tests/tp_01.cc:162-168) */
  stfem_op_t rhs_matrix_v;  /* wave only (tests/tp_01.cc:157-158) */
  int rhs_function_id;
  double frequency;
  int extrapolate;          /* initial guess = previous end value (time_integrators.h:180-190) */
  double gmres_tolerance;   /* ReductionControl reduce, default 1e-12 */
  double abs_tol;           /* default 1e-12 */
  int max_iterations;       /* default 200 */
  int max_basis_size;       /* default 100 */
} stfem_ti_desc;

int stfem_ti_create(const stfem_ti_desc *desc, stfem_ti_t *out);
int stfem_ti_destroy(stfem_ti_t ti);
/* TimeIntegratorFO::solve (time_integrators.h:300-321): rhs = rhs_matrix.vmult_slice(prev_x) + force,
 * x = extrapolated initial guess, FGMRES.  x, rhs: nb device arrays; prev_x: one spatial vector. */
int stfem_ti_solve_heat(stfem_ti_t ti, void *const *x, const void *prev_x, void *const *rhs, double time,
                        double time_step, int *iterations);
/* TimeIntegratorWave::solve (time_integrators.h:400-447) including the velocity recovery */
int stfem_ti_solve_wave(stfem_ti_t ti, void *const *u, void *const *v, void *const *rhs, const void *prev_u,
                        const void *prev_v, double time, double time_step, int *iterations);
int stfem_ti_last_residuals(stfem_ti_t ti, double *initial, double *final_);
/* VectorTools::interpolate / create_right_hand_side of an analytic function (tests/tp_01.cc:382-408);
 * integrate ADDS scale * (f, phi_i) into dst, constrained rows untouched */
int stfem_interpolate(stfem_mesh_t mesh, int degree, int function_id, double frequency, double time, void *dst);
int stfem_integrate_rhs(stfem_mesh_t mesh, int degree, int function_id, double frequency, double time, double scale,
                        void *dst);
/* ErrorCalculator::evaluate_error (include/exact_solution.h:533-633) over one solve interval against
 * ExactSolution: out3[0] += L2^2 part, out3[1] = max(out3[1], Linf), out3[2] += H1-seminorm^2 part */
int stfem_evaluate_error(stfem_mesh_t mesh, int degree, int time_type, int time_degree, int n_timesteps_at_once,
                         const void *const *x, const void *prev_x, double time, double time_step, double frequency,
                         int n_space_quad, double *out3);

/* Point-evaluation functionals (tests/tp_01.cc:455-481, 584-635: RemotePointEvaluation + FEPointEvaluation of every
 * time DoF at fixed points): out[b * n_points + p] = u_b(points[p]) for the nb DOUBLE device vectors x[b] of FE_Q(degree)
 * on `mesh`; points = n_points * dim doubles (host).  Points outside the mesh are an error. */
int stfem_point_evaluate(stfem_mesh_t mesh, int degree, int n_points, const double *points, int nb, const void *const *x,
                         double *out);

/* ---- host-side time algebra (no GPU): reference include/fe_time.h, include/fe_time.cc.
 *      type: 1 = CGP, 2 = DG (enum TimeStepType, fe_time.h:18-23).  All matrices row-major doubles. ---- */
int stfem_fe_time_n_blocks(int type, int r, int n_timesteps_at_once);
/* get_fe_time_weights (fe_time.h:351-409): Alpha*tau, Beta (nb x nb), Gamma, Zeta (nb x 1) */
int stfem_fe_time_weights(int type, int r, double tau, int n_timesteps_at_once, double *Alpha, double *Beta,
                          double *Gamma, double *Zeta);
/* get_fe_time_weights_wave (fe_time.h:157-305): inputs nd x nd / nd x 1 single-step matrices,
 * outputs (nd*nts)^2, (nd*nts)^2 and three (nd*nts) x 1 */
int stfem_fe_time_weights_wave(int type, int nd, const double *Alpha, const double *Beta, const double *Gamma,
                               const double *Zeta, int n_timesteps_at_once, double *lhs_uK, double *lhs_uM,
                               double *rhs_uK, double *rhs_uM, double *rhs_vM);
/* kind 0: get_time_projection_matrix(type, r_src=a, r_dst=b, nts)   (fe_time.h:749-805)
 * kind 1: get_time_prolongation_matrix(type, r=a, nts)              (fe_time.h:807-851)
 * kind 2: get_time_restriction_matrix(type, r=a, nts)               (fe_time.h:853-898)
 * out may be NULL to query rows/cols. */
int stfem_time_transfer_matrix(int kind, int type, int a, int b, int n_timesteps_at_once, double *out, int capacity,
                               int *rows, int *cols);
/* get_poly_mg_sequence (fe_time.cc:40-56); p_sequence: 0 bisect, 1 decrease_by_one, 2 go_to_one */
int stfem_poly_mg_sequence(int k_max, int k_min, int p_sequence, int *out, int capacity, int *count);
/* get_mg_sequence (fe_time.cc:58-127): level types 't','k','h','p' coarse -> fine, NUL-terminated.
 * coarsening_type: 0 space_or_time, 1 space_and_time (types.h:103-107) */
int stfem_mg_sequence(int n_sp_lvl, int n_k_seq, int n_p_seq, int n_timesteps_at_once, int n_timesteps_at_once_min,
                      char lower_lvl, int coarsening_type, int time_before_space, int use_p_multigrid_space,
                      int zip_from_back, char *out, int capacity);
/* get_precondition_stmg_types (fe_time.cc:129-150): out has strlen(seq)+1 entries */
int stfem_precondition_stmg_types(const char *seq, int coarsening_type, int time_before_space, int smoother, int *out);
/* 1D rules on [0,1]: kind 0 QGauss, 1 QGaussLobatto, 2 QGaussRadau(right) */
int stfem_quadrature_rule(int kind, int n, double *x, double *w);

/* ---- host-side problem data of the "practical" runs (spaceTimeConvergenceTest = false, tests/tp_01.cc:118-119,
 *      279-280, 374-381; tests/json/practical01.json).  Set-up work, no GPU.  Meshes are described like in
 *      stfem_mesh_create (n_cells, box, optional lexicographic vertices). ---- */
/* Coefficient<dim>'s per-coarse-cell factors (include/operators.h:905-921): prod(subdivisions) values
 * U(1-dc, 1+dc) from boost::mt19937(default_seed), one 32-bit draw each, in Table order (last index fastest) */
int stfem_coefficient_distortion(int dim, const int *subdivisions, double distort_coeff, double *table);
/* MatrixFreeOperator::evaluate_coefficient(Coefficient<dim>) (operators.h:1060-1087, 870-965): the coefficient at the
 * QGauss(degree+1) points of every cell, out[n_cells * (degree+1)^dim] in the layout stfem_op_desc::laplace_coeff_q
 * takes.  c123: NULL = {1, 9, 16}.  subdivisions / coeff_lower / coeff_upper: Parameters::subdivisions and
 * hyperrect corners (the table's grid, operators.h:922-925); ignored when distort_coeff == 0 */
int stfem_coefficient_at_qpoints(int dim, const int *n_cells, const double *lower, const double *upper,
                                 const double *vertices, int degree, const int *subdivisions, const double *coeff_lower,
                                 const double *coeff_upper, double distort_coeff, const double *c123, double *out);
/* VectorTools::interpolate of Functions::CutOffFunctionCinfty<dim>(radius, center, 1, invalid, integrate_to_one)
 * (tests/tp_01.cc:376-378, 393-400, 551) on the FE_Q(degree) support points: out[N], lexicographic */
int stfem_cutoff_cinfty_interpolate(int dim, const int *n_cells, const double *lower, const double *upper,
                                    const double *vertices, int degree, double radius, const double *center,
                                    int integrate_to_one, double *out);

#ifdef __cplusplus
}
#endif
#endif /* STFEM_B200_H */
