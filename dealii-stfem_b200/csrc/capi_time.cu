// C ABI: time stepping around the solve — TimeIntegratorFO / TimeIntegratorWave of the reference
// (include/time_integrators.h:24-459), source integration / interpolation (tests/tp_01.cc:382-408) and the
// space-time error functional (include/exact_solution.h:503-649).
#include "assemble.cuh"
#include "mg.cuh"

using namespace stfem;

namespace
{
  struct Combine { const double *src[2 * STFEM_MAX_BLOCKS + 2]; double c[2 * STFEM_MAX_BLOCKS + 2]; int m; };

  __global__ void k_combine(long long n, Combine cb, double *__restrict__ dst, int add)
  {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      {
        double s = add ? dst[i] : 0.0;
        for (int k = 0; k < cb.m; ++k) s += cb.c[k] * cb.src[k][i];
        dst[i] = s;
      }
  }

  void combine(stfem_ctx *ctx, long long n, const std::vector<const double *> &src, const std::vector<double> &c, double *dst, bool add)
  {
    Combine cb;
    cb.m = 0;
    for (size_t k = 0; k < src.size(); ++k)
      if (c[k] != 0.0)
        {
          cb.src[cb.m] = src[k];
          cb.c[cb.m++] = c[k];
        }
    if (cb.m == 0 && add) return;
    k_combine<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(n, cb, dst, add ? 1 : 0);
    ctx->launches++;
  }
} // namespace

struct stfem_time_integrator
{
  int                 type = DG, r = 0, nts = 1, problem = 1, nd = 1, fid = 0;
  bool                extrapolate = true;
  double              freq = 1.0, abstol = 1e-12, reduce = 1e-12;
  int                 max_iter = 200, max_basis = 100;
  Mat                 A1, B1, G1, Z1, AixB, AixG, AixZ;
  stfem_op           *matrix = nullptr, *rhs_matrix = nullptr, *rhs_matrix_v = nullptr;
  stfem_mg           *mg = nullptr;
  std::vector<double> quad_time;
  Fgmres              solver;
  FgmresResult        last;
  AsmGeom             geom;
  BlockVec<double>    tmp; // one spatial vector
  BlockVec<double>    force_space; // Cartesian meshes: time-independent spatial part of the source term
  double             *d_sin_tab = nullptr;
  ~stfem_time_integrator()
  {
    if (d_sin_tab) cudaFree(d_sin_tab);
  }

  int assemble_force(void *const *rhs, double time, double tau)
  {
    // time_integrators.h:73-110
    stfem_ctx *ctx = matrix->mesh->ctx;
    if (fid == 0) return STFEM_OK;
    const long long total = geom.n_cells * (geom.dim == 3 ? geom.n1 * geom.n1 * geom.n1 : geom.n1 * geom.n1);
    // Cartesian mesh: f(x, t) = amp(t) * prod sin(2 pi f x_d)  ->  the spatial integral (f_space, phi_i) is computed ONCE
    // and reused for every quadrature point in time and every time step; each call reduces to rhs_b += coeff_b * F
    const bool          separable = geom.sin_tab[0] != nullptr;
    std::vector<double> coeff(nts * nd, 0.0);
    if (separable && !force_space.d)
      {
        STFEM_FORWARD(force_space.alloc(ctx, 1, matrix->N));
        k_integrate_function<<<grid_for(ctx, total, 128), 128, 0, ctx->stream>>>(geom, -1, 0.0, freq, 1.0, force_space.d);
        ctx->launches++;
      }
    auto add = [&](int block, double scale, double t) {
      if (scale == 0.0) return;
      if (separable)
        {
          coeff[block] += scale * analytic_amp(fid, geom.dim, t, freq);
          return;
        }
      k_integrate_function<<<grid_for(ctx, total, 128), 128, 0, ctx->stream>>>(geom, fid, t, freq, scale, (double *)rhs[block]);
      ctx->launches++;
    };
    for (int it = 0; it < nts; ++it)
      for (size_t j = 0; j < quad_time.size(); ++j)
        {
          const double t = time + tau * it + tau * quad_time[j];
          if (type == DG)
            add(it * nd + (int)j, A1((int)j, (int)j), t);
          else if (j == 0)
            for (int i = 0; i < nd; ++i) add(it * nd + i, -G1(i, 0), t);
          else
            add(it * nd + (int)j - 1, A1((int)j - 1, (int)j - 1), t);
        }
    if (separable)
      for (int b = 0; b < nts * nd; ++b)
        if (coeff[b] != 0.0)
          {
            k_axpy<double><<<grid_for(ctx, matrix->N, 256), 256, 0, ctx->stream>>>(matrix->N, coeff[b], force_space.d, (double *)rhs[b]);
            ctx->launches++;
          }
    STFEM_CUDA_CHECK(cudaGetLastError());
    return STFEM_OK;
  }

  int halo_add(void *const *blocks, int nb)
  {
    stfem_mesh *m = matrix->mesh;
    if (!m->part.active || fid == 0) return STFEM_OK;
    return halo_compress_add<double>(m->ctx, m->part, matrix->halo, blocks, nb, matrix->np, m->dim);
  }

  int do_extrapolate(void *const *x, const void *prev)
  {
    // time_integrators.h:180-190
    stfem_ctx   *ctx = matrix->mesh->ctx;
    const size_t bytes = sizeof(double) * (size_t)matrix->N;
    for (int b = 0; b < nts * nd; ++b)
      if (extrapolate)
        STFEM_CUDA_CHECK(cudaMemcpyAsync(x[b], prev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
      else
        STFEM_CUDA_CHECK(cudaMemsetAsync(x[b], 0, bytes, ctx->stream));
    return STFEM_OK;
  }
};

extern "C" {

int stfem_ti_create(const stfem_ti_desc *d, stfem_ti_t *out)
{
  STFEM_REQUIRE(d && out, "stfem_ti_create: null argument");
  STFEM_REQUIRE(d->time_type == CGP || d->time_type == DG, "stfem_ti_create: bad time type");
  STFEM_REQUIRE(d->matrix && d->rhs_matrix, "stfem_ti_create: matrix / rhs_matrix missing");
  STFEM_REQUIRE(d->problem == 1 || (d->problem == 2 && d->rhs_matrix_v), "stfem_ti_create: wave needs rhs_matrix_v");
  STFEM_REQUIRE(d->Alpha_1 && d->Gamma_1, "stfem_ti_create: Alpha_1 / Gamma_1 missing");
  auto ti      = std::make_unique<stfem_time_integrator>();
  ti->type     = d->time_type;
  ti->r        = d->time_degree;
  ti->nts      = d->n_timesteps_at_once;
  ti->problem  = d->problem;
  ti->nd       = ti->type == DG ? ti->r + 1 : ti->r;
  ti->fid      = d->rhs_function_id;
  ti->freq     = d->frequency;
  ti->extrapolate = d->extrapolate != 0;
  ti->abstol   = d->abs_tol > 0 ? d->abs_tol : 1e-12;
  ti->reduce   = d->gmres_tolerance > 0 ? d->gmres_tolerance : 1e-12;
  ti->max_iter = d->max_iterations > 0 ? d->max_iterations : 200;
  ti->max_basis = d->max_basis_size > 0 ? d->max_basis_size : 100;
  ti->matrix = d->matrix; ti->rhs_matrix = d->rhs_matrix; ti->rhs_matrix_v = d->rhs_matrix_v; ti->mg = d->preconditioner;
  STFEM_REQUIRE(ti->matrix->nb_rows == ti->nts * ti->nd, "stfem_ti_create: matrix has %d blocks, expected %d", ti->matrix->nb_rows, ti->nts * ti->nd);
  const int nd = ti->nd;
  ti->A1 = Mat(nd, nd); ti->B1 = Mat(nd, nd); ti->G1 = Mat(nd, 1); ti->Z1 = Mat(nd, 1);
  std::copy(d->Alpha_1, d->Alpha_1 + nd * nd, ti->A1.a.begin());
  std::copy(d->Gamma_1, d->Gamma_1 + nd, ti->G1.a.begin());
  if (d->Beta_1) std::copy(d->Beta_1, d->Beta_1 + nd * nd, ti->B1.a.begin());
  if (d->Zeta_1) std::copy(d->Zeta_1, d->Zeta_1 + nd, ti->Z1.a.begin());
  ti->quad_time = time_nodes(ti->type, ti->r);
  if (ti->problem == 2)
    {
      // time_integrators.h:385-398
      STFEM_REQUIRE(d->Beta_1, "stfem_ti_create: wave needs Beta_1");
      try
        {
          const Mat Ainv = inverse(ti->A1);
          ti->AixB = Ainv * ti->B1; ti->AixG = Ainv * ti->G1; ti->AixZ = Ainv * ti->Z1;
        }
      catch (const std::exception &e)
        {
          set_error("stfem_ti_create: %s", e.what());
          return STFEM_ERR_INVALID;
        }
      if (ti->type == DG) ti->AixG.scale(-1.0); else ti->AixZ.scale(-1.0);
    }
  fill_asm_geom(ti->geom, ti->matrix->mesh, ti->matrix->degree, ti->matrix->degree + 1);
  STFEM_FORWARD(ti->tmp.alloc(ti->matrix->mesh->ctx, 1, ti->matrix->N));
  if (ti->fid != 0 && !ti->geom.vertices)
    {
      // separable source term on a Cartesian mesh: 1D tables of sin(2 pi f x_q) instead of 3 sin() per quadrature point
      const AsmGeom      &g = ti->geom;
      std::vector<double> tab;
      size_t              off[3] = {0, 0, 0};
      for (int d = 0; d < g.dim; ++d)
        {
          off[d] = tab.size();
          for (int c = 0; c < g.n[d]; ++c)
            for (int q = 0; q < g.nq1; ++q) tab.push_back(std::sin(2 * ST_PI * ti->freq * (g.lower[d] + g.h[d] * (c + g.xq[q]))));
        }
      stfem_ctx *ctx = ti->matrix->mesh->ctx;
      STFEM_CUDA_CHECK(cudaMalloc(&ti->d_sin_tab, tab.size() * sizeof(double)));
      STFEM_CUDA_CHECK(cudaMemcpyAsync(ti->d_sin_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      for (int d = 0; d < g.dim; ++d) ti->geom.sin_tab[d] = ti->d_sin_tab + off[d];
    }
  *out = ti.release();
  return STFEM_OK;
}

int stfem_ti_destroy(stfem_ti_t ti)
{
  delete ti;
  return STFEM_OK;
}

// TimeIntegratorFO::solve (time_integrators.h:300-321)
int stfem_ti_solve_heat(stfem_ti_t ti, void *const *x, const void *prev_x, void *const *rhs, double time, double tau, int *iterations)
{
  STFEM_REQUIRE(ti && x && prev_x && rhs, "stfem_ti_solve_heat: null argument");
  stfem_ctx   *ctx = ti->matrix->mesh->ctx;
  const int    nb  = ti->nts * ti->nd;
  const size_t bytes = sizeof(double) * (size_t)ti->matrix->N;
  for (int b = 0; b < nb; ++b) STFEM_CUDA_CHECK(cudaMemsetAsync(rhs[b], 0, bytes, ctx->stream));
  const void *src[1] = {prev_x};
  // the source term is assembled cell-wise (partial sums on rank interfaces): sum it over the ranks before the
  // operator part, which does its own interface sum, is added
  STFEM_FORWARD(ti->assemble_force(rhs, time, tau));
  STFEM_FORWARD(ti->halo_add(rhs, nb));
  STFEM_FORWARD(op_apply(ti->rhs_matrix, rhs, src, 1, nb, ti->rhs_matrix->d_alpha, ti->rhs_matrix->d_beta, false));
  STFEM_FORWARD(ti->do_extrapolate(x, prev_x));
  const int rc = ti->solver.solve(ti->matrix, ti->mg ? ti->mg->impl.get() : nullptr, x, (const void *const *)rhs, ti->max_basis, ti->max_iter,
                                  ti->abstol, ti->reduce, ti->last);
  if (iterations) *iterations = ti->last.iterations;
  return rc;
}

// TimeIntegratorWave::solve (time_integrators.h:400-447)
int stfem_ti_solve_wave(stfem_ti_t ti, void *const *u, void *const *v, void *const *rhs, const void *prev_u, const void *prev_v,
                        double time, double tau, int *iterations)
{
  STFEM_REQUIRE(ti && u && v && rhs && prev_u && prev_v, "stfem_ti_solve_wave: null argument");
  STFEM_REQUIRE(ti->problem == 2, "stfem_ti_solve_wave: integrator was created for the heat equation");
  stfem_ctx      *ctx = ti->matrix->mesh->ctx;
  const int       nb = ti->nts * ti->nd, nd = ti->nd;
  const long long N  = ti->matrix->N;
  const size_t    bytes = sizeof(double) * (size_t)N;
  for (int b = 0; b < nb; ++b) STFEM_CUDA_CHECK(cudaMemsetAsync(rhs[b], 0, bytes, ctx->stream));
  const void *su[1] = {prev_u}, *sv[1] = {prev_v};
  STFEM_FORWARD(ti->assemble_force(rhs, time, tau));
  STFEM_FORWARD(ti->halo_add(rhs, nb));
  STFEM_FORWARD(op_apply(ti->rhs_matrix, rhs, su, 1, nb, ti->rhs_matrix->d_alpha, ti->rhs_matrix->d_beta, false));
  STFEM_FORWARD(ti->do_extrapolate(u, prev_u));
  STFEM_FORWARD(op_apply(ti->rhs_matrix_v, rhs, sv, 1, nb, ti->rhs_matrix_v->d_alpha, ti->rhs_matrix_v->d_beta, false));
  const int rc = ti->solver.solve(ti->matrix, ti->mg ? ti->mg->impl.get() : nullptr, u, (const void *const *)rhs, ti->max_basis, ti->max_iter,
                                  ti->abstol, ti->reduce, ti->last);
  if (iterations) *iterations = ti->last.iterations;
  if (rc != STFEM_OK) return rc;
  // velocity recovery: v = A^-1 B u (+ terms of the previous end values)
  for (int it = 0; it < ti->nts; ++it)
    {
      const double *pu = it == 0 ? (const double *)prev_u : (const double *)u[it * nd - 1];
      for (int i = 0; i < nd; ++i)
        {
          std::vector<const double *> src;
          std::vector<double>         c;
          for (int j = 0; j < nd; ++j)
            {
              src.push_back((const double *)u[it * nd + j]);
              c.push_back(ti->AixB(i, j));
            }
          if (ti->type == DG)
            {
              src.push_back(pu);
              c.push_back(ti->AixG(i, 0));
            }
          else
            {
              const double *pv = it == 0 ? (const double *)prev_v : (const double *)v[it * nd - 1];
              src.push_back(pv);
              c.push_back(ti->AixG(i, 0));
              src.push_back(pu);
              c.push_back(ti->AixZ(i, 0));
            }
          combine(ctx, N, src, c, (double *)v[it * nd + i], false);
        }
    }
  STFEM_CUDA_CHECK(cudaGetLastError());
  return STFEM_OK;
}

int stfem_ti_last_residuals(stfem_ti_t ti, double *initial, double *final_)
{
  STFEM_REQUIRE(ti, "null integrator");
  if (initial) *initial = ti->last.initial_residual;
  if (final_) *final_ = ti->last.final_residual;
  return STFEM_OK;
}

int stfem_interpolate(stfem_mesh_t mesh, int degree, int function_id, double frequency, double time, void *dst)
{
  STFEM_REQUIRE(mesh && dst && degree >= 1 && degree <= 6, "stfem_interpolate: bad arguments");
  AsmGeom g;
  fill_asm_geom(g, mesh, degree, degree + 1);
  stfem_ctx      *ctx = mesh->ctx;
  const long long N   = (long long)g.np[0] * g.np[1] * g.np[2];
  k_interpolate_function<<<grid_for(ctx, N, 256), 256, 0, ctx->stream>>>(g, function_id, time, frequency, (double *)dst);
  ctx->launches++;
  STFEM_CUDA_CHECK(cudaGetLastError());
  return STFEM_OK;
}

int stfem_integrate_rhs(stfem_mesh_t mesh, int degree, int function_id, double frequency, double time, double scale, void *dst)
{
  STFEM_REQUIRE(mesh && dst && degree >= 1 && degree <= 6, "stfem_integrate_rhs: bad arguments");
  AsmGeom g;
  fill_asm_geom(g, mesh, degree, degree + 1);
  stfem_ctx      *ctx   = mesh->ctx;
  const long long total = g.n_cells * (g.dim == 3 ? g.n1 * g.n1 * g.n1 : g.n1 * g.n1);
  k_integrate_function<<<grid_for(ctx, total, 128), 128, 0, ctx->stream>>>(g, function_id, time, frequency, scale, (double *)dst);
  ctx->launches++;
  STFEM_CUDA_CHECK(cudaGetLastError());
  return STFEM_OK;
}

// ErrorCalculator::evaluate_error (exact_solution.h:533-633) for one solve interval; adds tau*w_q*||.||^2 into
// out3[0] (L2^2) and out3[2] (H1-seminorm^2), max into out3[1] (Linf).  n_space_quad = points per direction of the
// spatial Gauss rule (the reference passes fe_degree+1 with fe_degree the TIME degree, tp_01.cc:492-498).
int stfem_evaluate_error(stfem_mesh_t mesh, int degree, int time_type, int time_degree, int n_timesteps_at_once, const void *const *x,
                         const void *prev_x, double time, double tau, double frequency, int n_space_quad, double *out3)
{
  STFEM_REQUIRE(mesh && x && prev_x && out3, "stfem_evaluate_error: null argument");
  STFEM_REQUIRE(n_space_quad >= 1 && n_space_quad <= 8, "stfem_evaluate_error: n_space_quad out of range");
  stfem_ctx *ctx = mesh->ctx;
  AsmGeom    g;
  fill_asm_geom(g, mesh, degree, n_space_quad);
  const long long N  = (long long)g.np[0] * g.np[1] * g.np[2];
  const int       nd = time_type == DG ? time_degree + 1 : time_degree;
  const Rule      tq = gauss(time_degree + 1);
  const auto      nodes = time_nodes(time_type, time_degree);
  double *d_u = nullptr, *d_out = nullptr;
  STFEM_CUDA_CHECK(cudaMalloc(&d_u, sizeof(double) * N));
  STFEM_CUDA_CHECK(cudaMalloc(&d_out, sizeof(double) * 3));
  const long long total = g.n_cells * (g.dim == 3 ? n_space_quad * n_space_quad * n_space_quad : n_space_quad * n_space_quad);
  for (int it = 0; it < n_timesteps_at_once; ++it)
    for (size_t q = 0; q < tq.x.size(); ++q)
      {
        const double  t  = time + tau * it + tq.x[q] * tau;
        const double *pv = it == 0 ? (const double *)prev_x : (const double *)x[nd * it - 1];
        std::vector<const double *> src;
        std::vector<double>         c;
        for (size_t i = 0; i < nodes.size(); ++i)
          {
            const double v = lagrange_value(nodes, (int)i, tq.x[q]);
            if (time_type == DG)
              src.push_back((const double *)x[nd * it + i]);
            else
              src.push_back(i == 0 ? pv : (const double *)x[nd * it + i - 1]);
            c.push_back(v);
          }
        combine(ctx, N, src, c, d_u, false);
        STFEM_CUDA_CHECK(cudaMemsetAsync(d_out, 0, sizeof(double) * 3, ctx->stream));
        k_error<<<grid_for(ctx, total, 128), 128, 0, ctx->stream>>>(g, d_u, t, frequency, d_out);
        ctx->launches++;
        double h[3];
        STFEM_CUDA_CHECK(cudaMemcpyAsync(h, d_out, sizeof(double) * 3, cudaMemcpyDeviceToHost, ctx->stream));
        STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        out3[0] += tau * tq.w[q] * h[0];
        out3[2] += tau * tq.w[q] * h[2];
        if (h[1] > out3[1]) out3[1] = h[1];
      }
  cudaFree(d_u);
  cudaFree(d_out);
  return STFEM_OK;
}

} // extern "C"

// ---- point evaluation functionals (tests/tp_01.cc:455-481, 584-635: Utilities::MPI::RemotePointEvaluation +
//      FEPointEvaluation of every time DoF at a few fixed points).  The host locates each point (cell + reference
//      coordinates, Newton inversion of the MappingQ1 map on perturbed meshes) and tabulates the (degree+1)^dim basis
//      values and DoF indices; one warp per (block, point) gathers and reduces on the device.
namespace
{
  struct PointPtrs { const double *p[STFEM_MAX_BLOCKS]; };

  __global__ void k_point_eval(int n_items, int n_points, int nc, PointPtrs blocks, const long long *__restrict__ idx,
                               const double *__restrict__ w, double *__restrict__ out)
  {
    const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (item >= n_items) return; // whole warps leave together
    const int     b = item / n_points, p = item - b * n_points;
    const double *u = blocks.p[b];
    double        s = 0.0;
    for (int i = lane; i < nc; i += 32) s += w[(long long)p * nc + i] * u[idx[(long long)p * nc + i]];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[item] = s;
  }

  // reference coordinates of x in cell c of a MappingQ1 mesh; false if Newton does not converge inside the cell
  bool invert_q1(const stfem_mesh *m, const int *c, const double *x, double *xi)
  {
    const int dim = m->dim;
    for (int a = 0; a < dim; ++a) xi[a] = 0.5;
    for (int it = 0; it < 30; ++it)
      {
        double f[3] = {0, 0, 0}, J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int v = 0; v < (1 << dim); ++v)
          {
            const int vb[3] = {v & 1, (v >> 1) & 1, (v >> 2) & 1};
            long long vid   = (long long)(c[0] + vb[0]) + (long long)(m->n[0] + 1) * (c[1] + vb[1]);
            if (dim == 3) vid += (long long)(m->n[0] + 1) * (m->n[1] + 1) * (c[2] + vb[2]);
            double N[3] = {1, 1, 1}, dN[3] = {0, 0, 0};
            for (int a = 0; a < dim; ++a)
              {
                N[a]  = vb[a] ? xi[a] : 1.0 - xi[a];
                dN[a] = vb[a] ? 1.0 : -1.0;
              }
            const double sh = N[0] * N[1] * N[2];
            for (int a = 0; a < dim; ++a) f[a] += m->h_vertices[vid * dim + a] * sh;
            for (int b = 0; b < dim; ++b)
              {
                double d = dN[b];
                for (int e = 0; e < dim; ++e)
                  if (e != b) d *= N[e];
                for (int a = 0; a < dim; ++a) J[a][b] += m->h_vertices[vid * dim + a] * d;
              }
          }
        double r[3] = {0, 0, 0}, nrm = 0;
        for (int a = 0; a < dim; ++a)
          {
            r[a] = x[a] - f[a];
            nrm += r[a] * r[a];
          }
        double dx[3] = {0, 0, 0};
        if (dim == 2)
          {
            const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
            dx[0]            = (J[1][1] * r[0] - J[0][1] * r[1]) / det;
            dx[1]            = (-J[1][0] * r[0] + J[0][0] * r[1]) / det;
          }
        else
          {
            const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
                         c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
            const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
            // Cramer's rule, column by column
            auto det3 = [](const double A[3][3]) {
              return A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                     A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
            };
            for (int col = 0; col < 3; ++col)
              {
                double A[3][3];
                for (int a = 0; a < 3; ++a)
                  for (int b = 0; b < 3; ++b) A[a][b] = b == col ? r[a] : J[a][b];
                dx[col] = det3(A) / det;
              }
          }
        for (int a = 0; a < dim; ++a) xi[a] += dx[a];
        if (std::sqrt(nrm) < 1e-14 * (1.0 + std::fabs(x[0]) + std::fabs(x[1]))) break;
      }
    for (int a = 0; a < dim; ++a)
      if (!(xi[a] >= -1e-10 && xi[a] <= 1.0 + 1e-10)) return false;
    return true;
  }
} // namespace

extern "C" {

int stfem_point_evaluate(stfem_mesh_t mesh, int degree, int n_points, const double *points, int nb, const void *const *x,
                         double *out)
{
  STFEM_REQUIRE(mesh && points && x && out, "stfem_point_evaluate: null argument");
  STFEM_REQUIRE(degree >= 1 && degree <= 6, "stfem_point_evaluate: degree %d out of range", degree);
  STFEM_REQUIRE(n_points >= 1 && nb >= 1 && nb <= STFEM_MAX_BLOCKS, "stfem_point_evaluate: bad counts");
  // partitioned mesh (collective call): every rank evaluates the points inside its own brick; the values are summed over
  // the ranks and divided by the number of ranks that found the point (a point on a rank interface has the same value
  // from either side)
  const bool         partitioned = mesh->part.active;
  std::vector<double> found_here(n_points, 1.0);
  stfem_ctx *ctx = mesh->ctx;
  const int  dim = mesh->dim, n1 = degree + 1;
  const int  nc  = dim == 3 ? n1 * n1 * n1 : n1 * n1;
  const Rule gl  = gauss_lobatto(n1);
  long long  np[3] = {1, 1, 1};
  for (int a = 0; a < dim; ++a) np[a] = (long long)degree * mesh->n[a] + 1;
  std::vector<long long> idx((size_t)n_points * nc);
  std::vector<double>    w((size_t)n_points * nc);
  for (int p = 0; p < n_points; ++p)
    {
      const double *xp = points + (size_t)p * dim;
      int           c[3] = {0, 0, 0};
      double        xi[3] = {0, 0, 0};
      bool          found = false;
      int           guess[3] = {0, 0, 0};
      for (int a = 0; a < dim; ++a)
        {
          const double h = (mesh->upper[a] - mesh->lower[a]) / mesh->n[a];
          int          g = (int)std::floor((xp[a] - mesh->lower[a]) / h);
          guess[a]       = g < 0 ? 0 : (g >= mesh->n[a] ? mesh->n[a] - 1 : g);
          xi[a]          = (xp[a] - mesh->lower[a]) / h - guess[a];
        }
      if (mesh->cartesian)
        {
          found = true;
          for (int a = 0; a < dim; ++a)
            {
              c[a] = guess[a];
              if (!(xi[a] >= -1e-12 && xi[a] <= 1.0 + 1e-12)) found = false;
            }
        }
      else
        {
          // vertices move by less than half a cell (GridTools::distort_random): the cell is the Cartesian guess or a neighbour;
          // candidates in lexicographic order, first hit wins (a point on a cell boundary has the same value from both sides)
          const int zlo = dim == 3 ? -1 : 0, zhi = dim == 3 ? 1 : 0;
          for (int dz = zlo; dz <= zhi && !found; ++dz)
            for (int dy = -1; dy <= 1 && !found; ++dy)
              for (int dx = -1; dx <= 1 && !found; ++dx)
                {
                  const int cc[3] = {guess[0] + dx, guess[1] + dy, guess[2] + dz};
                  bool      ok    = true;
                  for (int a = 0; a < dim; ++a)
                    if (cc[a] < 0 || cc[a] >= mesh->n[a]) ok = false;
                  if (!ok) continue;
                  double t[3];
                  if (invert_q1(mesh, cc, xp, t))
                    {
                      found = true;
                      for (int a = 0; a < 3; ++a) { c[a] = cc[a]; xi[a] = a < dim ? t[a] : 0.0; }
                    }
                }
        }
      if (!found && partitioned)
        {
          found_here[p] = 0.0;
          for (int o = 0; o < nc; ++o)
            {
              idx[(size_t)p * nc + o] = 0;
              w[(size_t)p * nc + o]   = 0.0;
            }
          continue;
        }
      STFEM_REQUIRE(found, "stfem_point_evaluate: point %d lies outside the mesh", p);
      double L[3][8];
      for (int a = 0; a < dim; ++a)
        for (int i = 0; i < n1; ++i) L[a][i] = lagrange_value(gl.x, i, xi[a]);
      const int n1z = dim == 3 ? n1 : 1;
      int       o   = 0;
      for (int iz = 0; iz < n1z; ++iz)
        for (int iy = 0; iy < n1; ++iy)
          for (int ix = 0; ix < n1; ++ix, ++o)
            {
              const long long gx = (long long)c[0] * degree + ix, gy = (long long)c[1] * degree + iy,
                              gz = dim == 3 ? (long long)c[2] * degree + iz : 0;
              idx[(size_t)p * nc + o] = gx + np[0] * (gy + np[1] * gz);
              w[(size_t)p * nc + o]   = L[0][ix] * L[1][iy] * (dim == 3 ? L[2][iz] : 1.0);
            }
    }
  long long *d_idx = nullptr;
  double    *d_w = nullptr, *d_out = nullptr;
  const int  n_items = nb * n_points;
  STFEM_CUDA_CHECK(cudaMalloc(&d_idx, sizeof(long long) * idx.size()));
  STFEM_CUDA_CHECK(cudaMalloc(&d_w, sizeof(double) * w.size()));
  STFEM_CUDA_CHECK(cudaMalloc(&d_out, sizeof(double) * n_items));
  STFEM_CUDA_CHECK(cudaMemcpyAsync(d_idx, idx.data(), sizeof(long long) * idx.size(), cudaMemcpyHostToDevice, ctx->stream));
  STFEM_CUDA_CHECK(cudaMemcpyAsync(d_w, w.data(), sizeof(double) * w.size(), cudaMemcpyHostToDevice, ctx->stream));
  PointPtrs bp;
  for (int b = 0; b < nb; ++b) bp.p[b] = (const double *)x[b];
  const int warps_per_cta = 4;
  k_point_eval<<<(n_items + warps_per_cta - 1) / warps_per_cta, 32 * warps_per_cta, 0, ctx->stream>>>(n_items, n_points, nc, bp, d_idx,
                                                                                                   d_w, d_out);
  ctx->launches++;
  STFEM_CUDA_CHECK(cudaGetLastError());
  STFEM_CUDA_CHECK(cudaMemcpyAsync(out, d_out, sizeof(double) * n_items, cudaMemcpyDeviceToHost, ctx->stream));
  STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_idx);
  cudaFree(d_w);
  cudaFree(d_out);
  if (partitioned)
    {
      std::vector<double> buf(out, out + n_items);
      buf.insert(buf.end(), found_here.begin(), found_here.end());
      for (size_t o = 0; o < buf.size(); o += 256)
        STFEM_FORWARD(stfem_ctx_allreduce(ctx, buf.data() + o, (int)std::min<size_t>(256, buf.size() - o), 0));
      for (int p = 0; p < n_points; ++p)
        {
          const double cnt = buf[(size_t)n_items + p];
          STFEM_REQUIRE(cnt >= 0.5, "stfem_point_evaluate: point %d lies outside the (global) mesh", p);
          for (int b = 0; b < nb; ++b) out[(size_t)b * n_points + p] = buf[(size_t)b * n_points + p] / cnt;
        }
    }
  return STFEM_OK;
}

} // extern "C"
