// C ABI: space-time multigrid preconditioner (GMG, reference include/stmg.h:1047-1344) and the FGMRES solve
// (SolverFGMRES as configured in include/time_integrators.h:56-59, called at :315 and :424).
#include "mg.cuh"

using namespace stfem;

extern "C" {

int stfem_mg_create(stfem_ctx_t ctx, const stfem_mg_desc *desc, stfem_mg_t *out)
{
  STFEM_REQUIRE(ctx && desc && out, "stfem_mg_create: null argument");
  STFEM_REQUIRE(desc->n_levels >= 1 && desc->level_ops && desc->smoother_types, "stfem_mg_create: bad level list");
  STFEM_REQUIRE(desc->n_levels == 1 || desc->mg_type_level, "stfem_mg_create: mg_type_level missing");
  STFEM_REQUIRE(desc->poly_time_sequence && desc->n_poly_time >= 1, "stfem_mg_create: poly_time_sequence missing");
  std::vector<stfem_op *> ops(desc->level_ops, desc->level_ops + desc->n_levels);
  for (int l = 0; l < desc->n_levels; ++l)
    {
      STFEM_REQUIRE(ops[l], "stfem_mg_create: level %d operator is null", l);
      STFEM_REQUIRE(ops[l]->number_type == ops[0]->number_type, "stfem_mg_create: mixed level precisions");
      STFEM_REQUIRE(ops[l]->mesh->ctx == ctx, "stfem_mg_create: level %d lives on another context", l);
    }
  std::string      types = desc->n_levels > 1 ? std::string(desc->mg_type_level) : std::string();
  std::vector<int> sm(desc->smoother_types, desc->smoother_types + desc->n_levels);
  std::vector<int> poly(desc->poly_time_sequence, desc->poly_time_sequence + desc->n_poly_time);
  MGOptions        o;
  o.smoothing_steps  = desc->smoothing_steps > 0 ? desc->smoothing_steps : 1;
  o.relaxation       = desc->relaxation;
  o.smoothing_range  = desc->smoothing_range > 0 ? desc->smoothing_range : 1.0;
  o.eig_n_iterations = desc->eig_n_iterations > 0 ? desc->eig_n_iterations : 20;
  o.variable         = desc->variable != 0;
  o.restrict_is_transpose_prolongate = desc->restrict_is_transpose_prolongate != 0;
  STFEM_REQUIRE(desc->inner_preconditioner == 0 || desc->inner_preconditioner == 1, "stfem_mg_create: inner_preconditioner must be 0 or 1");
  o.inner_preconditioner = desc->inner_preconditioner;
  STFEM_REQUIRE(desc->vanka_storage == 0 || desc->vanka_storage == 1, "stfem_mg_create: vanka_storage must be 0 or 1");
  o.vanka_storage = desc->vanka_storage;
  STFEM_REQUIRE(desc->coarse_grid_maxiter >= 0 && desc->coarse_grid_maxiter <= MAXK, "stfem_mg_create: coarse_grid_maxiter must be in 0..%d", MAXK);
  o.coarse_gmres_maxiter = desc->coarse_grid_maxiter;
  o.coarse_gmres_abstol  = desc->coarse_grid_abstol;
  STFEM_CUDA_CHECK(cudaSetDevice(ctx->device));
  auto mg         = std::make_unique<stfem_mg>();
  mg->number_type = ops[0]->number_type;
  int rc;
  if (mg->number_type == STFEM_F32)
    {
      auto impl = std::make_unique<Multigrid<float>>();
      rc        = impl->init(ctx, ops, types, sm, desc->time_type, desc->n_timesteps_at_once, poly, o);
      mg->impl  = std::move(impl);
    }
  else
    {
      auto impl = std::make_unique<Multigrid<double>>();
      rc        = impl->init(ctx, ops, types, sm, desc->time_type, desc->n_timesteps_at_once, poly, o);
      mg->impl  = std::move(impl);
    }
  if (rc != STFEM_OK) return rc;
  STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  *out = mg.release();
  return STFEM_OK;
}

int stfem_mg_destroy(stfem_mg_t mg)
{
  delete mg;
  return STFEM_OK;
}

int stfem_mg_vmult(stfem_mg_t mg, void *const *dst, const void *const *src)
{
  STFEM_REQUIRE(mg && dst && src, "stfem_mg_vmult: null argument");
  return mg->impl->vmult(dst, src);
}

int stfem_mg_n_levels(stfem_mg_t mg) { return mg ? mg->impl->n_levels() : 0; }

int stfem_mg_level_apply(stfem_mg_t mg, int level, int what, void *const *dst, const void *const *src)
{
  STFEM_REQUIRE(mg && dst && src, "stfem_mg_level_apply: null argument");
  return mg->impl->level_apply(level, what, dst, src);
}

int stfem_mg_level_info(stfem_mg_t mg, int level, double *out10)
{
  STFEM_REQUIRE(mg, "stfem_mg_level_info: null mg");
  return mg->impl->level_info(level, out10);
}

int stfem_solver_create(stfem_solver_t *out)
{
  STFEM_REQUIRE(out, "null out");
  *out = new stfem_solver();
  return STFEM_OK;
}

int stfem_solver_destroy(stfem_solver_t s)
{
  delete s;
  return STFEM_OK;
}

int stfem_fgmres_solve(stfem_solver_t s, stfem_op_t A, stfem_mg_t M, void *const *x, const void *const *b, int max_basis_size,
                       int max_iterations, double abs_tol, double reduce, int *iterations, double *initial_residual,
                       double *final_residual)
{
  STFEM_REQUIRE(s && A && x && b, "stfem_fgmres_solve: null argument");
  STFEM_REQUIRE(A->nb_rows == A->nb_cols, "stfem_fgmres_solve: operator not square");
  STFEM_REQUIRE(max_basis_size >= 1 && max_iterations >= 1, "stfem_fgmres_solve: bad iteration limits");
  const int rc = s->fgmres.solve(A, M ? M->impl.get() : nullptr, x, b, max_basis_size, max_iterations, abs_tol, reduce, s->last);
  if (iterations) *iterations = s->last.iterations;
  if (initial_residual) *initial_residual = s->last.initial_residual;
  if (final_residual) *final_residual = s->last.final_residual;
  return rc;
}

} // extern "C"
