// C ABI: host-side time algebra (reference include/fe_time.h, fe_time.cc).  No GPU work here.
#include "common.hpp"
#include "fe_time.hpp"

using namespace stfem;

namespace
{
  void copy_out(const Mat &M, double *out)
  {
    if (out) std::copy(M.a.begin(), M.a.end(), out);
  }
  int n_time_dofs(int type, int r) { return type == DG ? r + 1 : r; }
} // namespace

extern "C" {

int stfem_fe_time_n_blocks(int type, int r, int n_timesteps_at_once) { return n_time_dofs(type, r) * n_timesteps_at_once; }

int stfem_fe_time_weights(int type, int r, double tau, int n_timesteps_at_once, double *Alpha, double *Beta,
                          double *Gamma, double *Zeta)
{
  STFEM_REQUIRE(type == CGP || type == DG, "stfem_fe_time_weights: type must be 1 (CGP) or 2 (DG)");
  STFEM_REQUIRE(r >= (type == CGP ? 1 : 0) && r <= 12, "stfem_fe_time_weights: degree %d out of range", r);
  STFEM_REQUIRE(n_timesteps_at_once >= 1, "stfem_fe_time_weights: n_timesteps_at_once < 1");
  try
    {
      auto w = fe_time_weights(type, r, tau, n_timesteps_at_once);
      copy_out(w[0], Alpha); copy_out(w[1], Beta); copy_out(w[2], Gamma); copy_out(w[3], Zeta);
    }
  catch (const std::exception &e)
    {
      set_error("stfem_fe_time_weights: %s", e.what());
      return STFEM_ERR_INVALID;
    }
  return STFEM_OK;
}

int stfem_fe_time_weights_wave(int type, int nd, const double *Alpha, const double *Beta, const double *Gamma,
                               const double *Zeta, int n_timesteps_at_once, double *lhs_uK, double *lhs_uM,
                               double *rhs_uK, double *rhs_uM, double *rhs_vM)
{
  STFEM_REQUIRE(type == CGP || type == DG, "stfem_fe_time_weights_wave: bad type");
  STFEM_REQUIRE(nd >= 1 && Alpha && Beta && Gamma, "stfem_fe_time_weights_wave: null/empty input");
  try
    {
      Mat A(nd, nd), B(nd, nd), G(nd, 1), Z(nd, 1);
      std::copy(Alpha, Alpha + nd * nd, A.a.begin());
      std::copy(Beta, Beta + nd * nd, B.a.begin());
      std::copy(Gamma, Gamma + nd, G.a.begin());
      if (Zeta) std::copy(Zeta, Zeta + nd, Z.a.begin());
      auto w = fe_time_weights_wave(type, A, B, G, Z, n_timesteps_at_once);
      copy_out(w[0], lhs_uK); copy_out(w[1], lhs_uM); copy_out(w[2], rhs_uK); copy_out(w[3], rhs_uM); copy_out(w[4], rhs_vM);
    }
  catch (const std::exception &e)
    {
      set_error("stfem_fe_time_weights_wave: %s", e.what());
      return STFEM_ERR_INVALID;
    }
  return STFEM_OK;
}

// kind: 0 projection (k-transfer, a=r_src, b=r_dst), 1 prolongation (tau, a=r), 2 restriction (tau, a=r)
int stfem_time_transfer_matrix(int kind, int type, int a, int b, int n_timesteps_at_once, double *out, int capacity,
                               int *rows, int *cols)
{
  STFEM_REQUIRE(type == CGP || type == DG, "stfem_time_transfer_matrix: bad type");
  STFEM_REQUIRE(rows && cols, "stfem_time_transfer_matrix: null rows/cols");
  try
    {
      Mat M = kind == 0 ? time_projection_matrix(type, a, b, n_timesteps_at_once) :
              kind == 1 ? time_prolongation_matrix(type, a, n_timesteps_at_once) :
                          time_restriction_matrix(type, a, n_timesteps_at_once);
      *rows = M.m;
      *cols = M.n;
      if (out)
        {
          STFEM_REQUIRE(capacity >= M.m * M.n, "stfem_time_transfer_matrix: capacity %d < %d", capacity, M.m * M.n);
          copy_out(M, out);
        }
    }
  catch (const std::exception &e)
    {
      set_error("stfem_time_transfer_matrix: %s", e.what());
      return STFEM_ERR_INVALID;
    }
  return STFEM_OK;
}

int stfem_poly_mg_sequence(int k_max, int k_min, int p_sequence, int *out, int capacity, int *count)
{
  STFEM_REQUIRE(count, "null count");
  auto d = poly_mg_sequence(k_max, k_min, p_sequence);
  *count = (int)d.size();
  if (out)
    {
      STFEM_REQUIRE(capacity >= (int)d.size(), "capacity too small");
      std::copy(d.begin(), d.end(), out);
    }
  return STFEM_OK;
}

int stfem_mg_sequence(int n_sp_lvl, int n_k_seq, int n_p_seq, int n_timesteps_at_once, int n_timesteps_at_once_min,
                      char lower_lvl, int coarsening_type, int time_before_space, int use_p_multigrid_space,
                      int zip_from_back, char *out, int capacity)
{
  STFEM_REQUIRE(out && capacity > 0, "stfem_mg_sequence: null out");
  STFEM_REQUIRE(lower_lvl == 'k' || lower_lvl == 't', "stfem_mg_sequence: lower_lvl must be 'k' or 't'");
  STFEM_REQUIRE(n_sp_lvl >= 1 && n_k_seq >= 1, "stfem_mg_sequence: need >= 1 space level and k entry");
  std::string s = mg_sequence(n_sp_lvl, n_k_seq, n_p_seq, n_timesteps_at_once, n_timesteps_at_once_min, lower_lvl,
                              coarsening_type, time_before_space != 0, use_p_multigrid_space != 0, zip_from_back != 0);
  STFEM_REQUIRE((int)s.size() + 1 <= capacity, "stfem_mg_sequence: capacity too small");
  std::copy(s.begin(), s.end(), out);
  out[s.size()] = 0;
  return STFEM_OK;
}

int stfem_precondition_stmg_types(const char *seq, int coarsening_type, int time_before_space, int smoother, int *out)
{
  STFEM_REQUIRE(seq && out, "null argument");
  auto r = precondition_stmg_types(seq, coarsening_type, time_before_space != 0, smoother);
  std::copy(r.begin(), r.end(), out);
  return STFEM_OK;
}

// 1D rules for tests: kind 0 Gauss, 1 Gauss-Lobatto, 2 Gauss-Radau(right); n points on [0,1]
int stfem_quadrature_rule(int kind, int n, double *x, double *w)
{
  STFEM_REQUIRE(n >= 1 && n <= 32 && x && w, "stfem_quadrature_rule: bad arguments");
  STFEM_REQUIRE(kind != 1 || n >= 2, "Gauss-Lobatto needs n >= 2");
  Rule r = kind == 0 ? gauss(n) : kind == 1 ? gauss_lobatto(n) : gauss_radau_right(n);
  std::copy(r.x.begin(), r.x.end(), x);
  std::copy(r.w.begin(), r.w.end(), w);
  return STFEM_OK;
}

} // extern "C"
