// Host C++ through the façade (include/stfem_b200.hpp): the calls a deal.II-style driver makes on the reference's
// SystemMatrix / GMG / SolverFGMRES (tests/tp_01.cc:121-168, 337-351; include/time_integrators.h:300-321).
//   facade_demo <out.bin> [n_cells_per_dir] [degree]
// 3D heat, FE_Q(degree) x DG(1), two h-levels.  Writes y = A x for x_b[i] = sin(0.1 i + b) (masked on the Dirichlet
// boundary by the operator) to <out.bin> as nb*N doubles; then solves A u = A x_true with GMG-preconditioned FGMRES
// and prints iterations and the error.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "stfem_b200.hpp"

int main(int argc, char **argv)
{
  const char *out_path = argc > 1 ? argv[1] : "facade_demo.bin";
  const int   n        = argc > 2 ? std::atoi(argv[2]) : 8;
  const int   degree   = argc > 3 ? std::atoi(argv[3]) : 2;
  try
    {
      stfem::Context ctx(0);
      const int      type_dg = 2, r = 1, nts = 1;
      const double   tau = 0.025;
      const int      nb  = stfem_fe_time_n_blocks(type_dg, r, nts);
      std::vector<double> Alpha(nb * nb), Beta(nb * nb), Gamma(nb), Zeta(nb);
      stfem::check(stfem_fe_time_weights(type_dg, r, tau, nts, Alpha.data(), Beta.data(), Gamma.data(), Zeta.data()));

      stfem::Mesh fine(ctx, {n, n, n}), coarse(ctx, {n / 2, n / 2, n / 2});
      stfem::SystemMatrix<double> A(fine, degree, nb, nb, Alpha.data(), Beta.data());
      stfem::SystemMatrix<float>  A1(fine, degree, nb, nb, Alpha.data(), Beta.data());
      stfem::SystemMatrix<float>  A0(coarse, degree, nb, nb, Alpha.data(), Beta.data());

      stfem::BlockVector<double> x, y, b, u;
      A.initialize_dof_vector(x);
      A.initialize_dof_vector(y);
      A.initialize_dof_vector(b);
      A.initialize_dof_vector(u);
      const long long     N = A.m();
      std::vector<double> h(N);
      for (int blk = 0; blk < nb; ++blk)
        {
          for (long long i = 0; i < N; ++i) h[i] = std::sin(0.1 * (double)i + blk);
          x.copy_from_host(blk, h.data());
        }
      A.vmult(y, x);
      FILE *f = std::fopen(out_path, "wb");
      if (!f) return 2;
      for (int blk = 0; blk < nb; ++blk)
        {
          y.copy_to_host(blk, h.data());
          std::fwrite(h.data(), sizeof(double), (size_t)N, f);
        }
      std::fclose(f);

      // GMG(levels coarse -> fine, "h", smoothers: relaxation on both) handed to FGMRES
      stfem::GMG<float>   gmg(ctx, {&A0, &A1}, "h", {1, 1}, type_dg, nts, {r});
      stfem::SolverFGMRES solver(200, 1e-12, 1e-12, 100);
      A.vmult(b, x); // rhs of a system whose solution is x on the unconstrained DoFs
      solver.solve(A, u, b, gmg);
      // ||A u - b||_inf / ||b||_inf
      A.vmult(y, u);
      std::vector<double> hb(N);
      double              res = 0, nrm = 0;
      for (int blk = 0; blk < nb; ++blk)
        {
          y.copy_to_host(blk, h.data());
          b.copy_to_host(blk, hb.data());
          for (long long i = 0; i < N; ++i)
            {
              res = std::fmax(res, std::fabs(h[i] - hb[i]));
              nrm = std::fmax(nrm, std::fabs(hb[i]));
            }
        }
      std::printf("N %lld nb %d iterations %u initial %.6e final %.6e rel_residual_inf %.3e launches %lld\n", N, nb, solver.last_step(),
                  solver.initial_value(), solver.last_value(), res / nrm, ctx.launch_count());
      // get_matrix_diagonal (operators.h:613-625): checksum of the entries for the Python side of the test
      {
        auto                diag = A.get_matrix_diagonal();
        std::vector<double> h((size_t)A.m());
        double              sum = 0;
        for (unsigned b = 0; b < diag->get_vector().n_blocks(); ++b)
          {
            diag->get_vector().copy_to_host(b, h.data());
            for (double v : h) sum += v;
          }
        std::printf("diagonal_sum %.15e\n", sum);
      }
      // error path: a non-square operator must be rejected by vmult like the reference's dimension Assert
      stfem::SystemMatrix<double> S(fine, degree, nb, 1, Gamma.data(), Zeta.data());
      bool                        threw = false;
      try
        {
          S.vmult(y, x);
        }
      catch (const stfem::Error &e)
        {
          threw = true;
        }
      std::printf("non-square vmult rejected: %s\n", threw ? "yes" : "no");
      return threw ? 0 : 3;
    }
  catch (const stfem::Error &e)
    {
      std::fprintf(stderr, "%s\n", e.what());
      return 1;
    }
}
