// tp_01 of the reference (tests/tp_01.cc:56-848) as a C++ host program on the C ABI (include/stfem_b200.h):
//
//     tp01_main --file tests/json/tf03.json --dim 2 [--precondition_double] [--max_steps n]
//
// reads a parameter file with the keys of include/parameters.h:92-144 (flat JSON object), runs the
// nDegCycles x nRefCycles loop of space-time convergence tests (the convergence_test lambda, tp_01.cc:56-725: level
// hierarchy :171-321, operators :114-168, time loop :646-702) and prints what tp_01 prints on rank 0: the per-run header
// (:101-104, 210, 703-707), one "Convergence table k=..." per degree (:712-760) and the "Iteration count table"
// (:761-764).  Everything numerical happens in libstfem_b200.so; this file is the host bookkeeping the reference's driver
// does.  Scope: spaceTimeConvergenceTest = true on the undistorted mesh (the practical set-up with coefficient table and
// point functionals is driven by dealii-stfem_b200/driver.py through the same C entry points).
#include <stfem_b200.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace
{
  void check(int rc)
  {
    if (rc != STFEM_OK) throw std::runtime_error(std::string("stfem: ") + stfem_last_error());
  }

  // ------------------------------------------------------------------ Parameters<dim> (include/parameters.h:12-176)
  using Json = std::map<std::string, std::string>;
  Json read_flat_json(const std::string &path)
  {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string s = ss.str();
    Json              out;
    size_t            i = 0;
    auto skip = [&]() { while (i < s.size() && (std::isspace((unsigned char)s[i]) || s[i] == ',' || s[i] == '{' || s[i] == '}')) ++i; };
    auto token = [&]() -> std::string {
      std::string t;
      if (s[i] == '"')
        {
          for (++i; i < s.size() && s[i] != '"'; ++i) t += s[i];
          ++i;
        }
      else
        while (i < s.size() && s[i] != ',' && s[i] != '}' && !std::isspace((unsigned char)s[i])) t += s[i++];
      return t;
    };
    for (skip(); i < s.size(); skip())
      {
        const std::string key = token();
        while (i < s.size() && (std::isspace((unsigned char)s[i]) || s[i] == ':')) ++i;
        if (i >= s.size()) break;
        out[key] = token();
      }
    return out;
  }

  struct Parameters
  {
    bool        space_time_mg = true, mg_time_before_space = false, space_time_level_first = true, use_pmg = false;
    std::string time_type = "CGP", problem = "wave", coarsening_type = "space_or_time", p_mg_type = "bisect", smoother = "relaxation";
    std::string coarse_grid_smoother_type = "Smoother";
    int         n_timesteps_at_once = 1, n_timesteps_at_once_min = -1, fe_degree = 1, fe_degree_min = -1, fe_degree_min_space = -1;
    int         n_deg_cycles = 1, n_ref_cycles = 1, refinement = 2, smoothing_steps = 1, eig_n_iterations = 20;
    double      frequency = 1.0, end_time = 1.0, smoothing_range = 1.0, relaxation = 0.0;
    bool        space_time_convergence_test = true, extrapolate = true, restrict_is_transpose_prolongate = true, variable = true;
    std::vector<double> lower, upper;
    std::vector<int>    subdivisions;

    Parameters(const Json &j, int dim) : lower(dim, 0.0), upper(dim, 1.0), subdivisions(dim, 1)
    {
      auto b = [&](const char *k, bool &v) { auto it = j.find(k); if (it != j.end()) v = it->second == "true" || it->second == "True"; };
      auto n = [&](const char *k, int &v) { auto it = j.find(k); if (it != j.end()) v = std::atoi(it->second.c_str()); };
      auto d = [&](const char *k, double &v) { auto it = j.find(k); if (it != j.end()) v = std::atof(it->second.c_str()); };
      auto s = [&](const char *k, std::string &v) { auto it = j.find(k); if (it != j.end()) v = it->second; };
      auto list = [&](const char *k, auto &v) {
        auto it = j.find(k);
        if (it == j.end()) return;
        std::stringstream ss(it->second);
        std::string       t;
        for (size_t a = 0; a < v.size() && std::getline(ss, t, ','); ++a) v[a] = (typename std::decay_t<decltype(v)>::value_type)std::atof(t.c_str());
      };
      b("spaceTimeMg", space_time_mg); b("mgTimeBeforeSpace", mg_time_before_space); b("spaceTimeLevelFirst", space_time_level_first);
      b("usePMg", use_pmg); s("timeType", time_type); s("problemType", problem); s("coarseningType", coarsening_type);
      s("pMgType", p_mg_type); s("smoother", smoother); s("coarseGridSmootherType", coarse_grid_smoother_type);
      n("nTimestepsAtOnce", n_timesteps_at_once); n("nTimestepsAtOnceMin", n_timesteps_at_once_min); n("feDegree", fe_degree);
      n("feDegreeMin", fe_degree_min); n("feDegreeMinSpace", fe_degree_min_space); n("nDegCycles", n_deg_cycles);
      n("nRefCycles", n_ref_cycles); n("refinement", refinement); n("smoothingSteps", smoothing_steps);
      n("smoothingEigCgNIterations", eig_n_iterations); d("frequency", frequency); d("endTime", end_time);
      d("smoothingRange", smoothing_range); d("relaxation", relaxation); b("spaceTimeConvergenceTest", space_time_convergence_test);
      b("extrapolate", extrapolate); b("restrictIsTransposeProlongate", restrict_is_transpose_prolongate); b("variable", variable);
      list("hyperRectLowerLeft", lower); list("hyperRectUpperRight", upper); list("subdivisions", subdivisions);
      for (auto &c : smoother) c = (char)std::tolower((unsigned char)c);
      // parameters.h:146-176
      const int nts = n_timesteps_at_once;
      if (n_timesteps_at_once_min == -1) n_timesteps_at_once_min = nts / 2;
      n_timesteps_at_once_min = std::min(std::max(n_timesteps_at_once_min, 1), nts);
      const int lowest = time_type == "DG" ? 0 : 1;
      if (fe_degree_min == -1) fe_degree_min = fe_degree - 1;
      fe_degree_min = std::min(std::max(fe_degree_min, lowest), fe_degree);
      if (fe_degree_min_space == -1) fe_degree_min_space = fe_degree_min;
    }
  };

  // ------------------------------------------------------------------ small RAII helpers over the C handles
  struct DeviceVectors
  {
    stfem_ctx_t         ctx;
    std::vector<void *> p;
    DeviceVectors(stfem_ctx_t c, int nb, long long n) : ctx(c), p(nb, nullptr)
    {
      for (auto &q : p)
        {
          check(stfem_dev_alloc(ctx, (size_t)n * 8, &q));
          check(stfem_dev_memset(ctx, q, 0, (size_t)n * 8));
        }
    }
    ~DeviceVectors() { for (void *q : p) stfem_dev_free(ctx, q); }
  };

  struct TimeWeights { std::vector<double> A, B, G, Z; int nb = 0; };
  TimeWeights fe_time_weights(int type, int r, double tau, int nts)
  {
    TimeWeights w;
    w.nb = stfem_fe_time_n_blocks(type, r, nts);
    w.A.assign((size_t)w.nb * w.nb, 0); w.B = w.A; w.G.assign(w.nb, 0); w.Z = w.G;
    check(stfem_fe_time_weights(type, r, tau, nts, w.A.data(), w.B.data(), w.G.data(), w.Z.data()));
    return w;
  }
  struct WaveWeights { std::vector<double> lhs_uK, lhs_uM, rhs_uK, rhs_uM, rhs_vM; int nt = 0; };
  WaveWeights fe_time_weights_wave(int type, const TimeWeights &w1, int nts)
  {
    WaveWeights w;
    w.nt = w1.nb * nts;
    w.lhs_uK.assign((size_t)w.nt * w.nt, 0); w.lhs_uM = w.lhs_uK;
    w.rhs_uK.assign(w.nt, 0); w.rhs_uM = w.rhs_uK; w.rhs_vM = w.rhs_uK;
    check(stfem_fe_time_weights_wave(type, w1.nb, w1.A.data(), w1.B.data(), w1.G.data(), w1.Z.data(), nts, w.lhs_uK.data(), w.lhs_uM.data(),
                                     w.rhs_uK.data(), w.rhs_uM.data(), w.rhs_vM.data()));
    return w;
  }
  stfem_op_t make_op(stfem_mesh_t mesh, int degree, int number_type, int nb_rows, int nb_cols, const double *Alpha, const double *Beta)
  {
    stfem_op_desc d;
    std::memset(&d, 0, sizeof(d));
    d.degree = degree; d.number_type = number_type; d.nb_rows = nb_rows; d.nb_cols = nb_cols; d.Alpha = Alpha; d.Beta = Beta;
    stfem_op_t op = nullptr;
    check(stfem_op_create(mesh, &d, &op));
    return op;
  }

  struct Row
  {
    long long cells, s_dofs;
    int       t_dofs, iterations, timesteps, n_levels;
    double    linf, l2, h1;
  };

  // ------------------------------------------------------------------ one (refinement, degree) run: tp_01.cc:56-725
  Row convergence_test(stfem_ctx_t ctx, const Parameters &p, int dim, int refinement, int fe_degree, int mg_number_type, int max_steps)
  {
    const int  type = p.time_type == "CGP" ? 1 : 2;
    const bool cgp = type == 1, wave = p.problem == "wave";
    const int  r = fe_degree, k = fe_degree + 1; // tp_01.cc:77
    const int  nts = p.n_timesteps_at_once, nd = cgp ? r : r + 1, nb = nd * nts;
    std::vector<int> n_cells(dim);
    double           diam2 = 0;
    for (int a = 0; a < dim; ++a)
      {
        n_cells[a] = p.subdivisions[a] << refinement;
        const double h = (p.upper[a] - p.lower[a]) / p.subdivisions[a];
        diam2 += h * h;
      }
    const double spc_step = std::sqrt(diam2) / std::sqrt((double)dim); // tp_01.cc:87
    const int    n_steps  = (int)((p.end_time - 0.0) / spc_step);
    const double tau      = p.end_time * std::pow(2.0, -(refinement + 1)) / n_steps; // tp_01.cc:106-109
    // ---- level hierarchy (tp_01.cc:171-214)
    const int pseq = p.p_mg_type == "bisect" ? 0 : (p.p_mg_type == "decrease_by_one" ? 1 : 2);
    auto poly_sequence = [&](int kmax, int kmin) {
      std::vector<int> out(64);
      int              cnt = 0;
      check(stfem_poly_mg_sequence(kmax, kmin, pseq, out.data(), 64, &cnt));
      out.resize(cnt);
      return out;
    };
    const std::vector<int> poly_time  = poly_sequence(fe_degree, p.space_time_mg ? p.fe_degree_min : fe_degree);
    const std::vector<int> poly_space = poly_sequence(fe_degree, p.fe_degree_min_space);
    const int              nts_min    = p.space_time_mg ? std::max(p.n_timesteps_at_once_min, 1) : nts;
    const int              coarsening = p.coarsening_type == "space_or_time" ? 0 : 1;
    char                   seq[256];
    check(stfem_mg_sequence(refinement + 1, (int)poly_time.size(), (int)poly_space.size(), nts, nts_min, 't', coarsening,
                            p.mg_time_before_space, p.use_pmg, p.space_time_level_first, seq, 256));
    const std::string mg_type_level(seq);
    const int         nl = (int)mg_type_level.size() + 1;
    std::vector<int>  level_ref(nl), level_degree(nl);
    {
      int rr = refinement;
      level_ref[nl - 1] = rr;
      for (int ii = nl - 2; ii >= 0; --ii)
        {
          if (mg_type_level[ii] == 'h') --rr;
          level_ref[ii] = rr;
        }
      int fi = p.use_pmg ? 0 : (int)poly_space.size() - 1;
      for (int l = 0; l < nl; ++l)
        {
          level_degree[l] = poly_space[fi] + 1;
          if (p.use_pmg && l < nl - 1 && mg_type_level[l] == 'p') ++fi;
        }
    }
    const int        smoother = p.smoother == "relaxation" ? 1 : (p.smoother == "chebyshev" ? 2 : 0);
    std::vector<int> ptypes(nl);
    check(stfem_precondition_stmg_types(mg_type_level.c_str(), coarsening, p.mg_time_before_space, smoother, ptypes.data()));
    // time weights per level (fe_time.h:411-474)
    std::vector<TimeWeights> lw(nl);
    std::vector<WaveWeights> lww(nl);
    {
      int    pi = (int)poly_time.size() - 1, n = nts;
      double t  = tau;
      lw[nl - 1] = fe_time_weights(type, poly_time[pi], t, n);
      for (int idx = nl - 2; idx >= 0; --idx)
        {
          const char mgt = mg_type_level[idx];
          if (mgt == 'k') --pi;
          else if (mgt == 't') { n /= 2; t *= 2; }
          lw[idx] = fe_time_weights(type, poly_time[pi], t, n);
        }
      if (wave)
        for (int l = 0; l < nl; ++l) lww[l] = fe_time_weights_wave(type, lw[l], 1);
    }
    // ---- meshes per refinement, level operators, multigrid
    std::map<int, stfem_mesh_t> meshes;
    for (int rf : level_ref)
      if (!meshes.count(rf))
        {
          std::vector<int> n(dim);
          for (int a = 0; a < dim; ++a) n[a] = p.subdivisions[a] << rf;
          stfem_mesh_t m = nullptr;
          check(stfem_mesh_create(ctx, dim, n.data(), p.lower.data(), p.upper.data(), nullptr, dim == 3 ? 0x3fu : 0xfu, &m));
          meshes[rf] = m;
        }
    std::vector<stfem_op_t> level_ops(nl);
    for (int l = 0; l < nl; ++l)
      {
        const int lnb = wave ? lww[l].nt : lw[l].nb;
        level_ops[l]  = make_op(meshes[level_ref[l]], level_degree[l], mg_number_type, lnb, lnb, wave ? lww[l].lhs_uK.data() : lw[l].A.data(),
                                wave ? lww[l].lhs_uM.data() : lw[l].B.data());
      }
    stfem_mg_desc md;
    std::memset(&md, 0, sizeof(md));
    md.n_levels = nl; md.level_ops = level_ops.data(); md.mg_type_level = mg_type_level.c_str(); md.smoother_types = ptypes.data();
    md.time_type = type; md.n_timesteps_at_once = nts; md.poly_time_sequence = poly_time.data(); md.n_poly_time = (int)poly_time.size();
    md.smoothing_steps = p.smoothing_steps; md.relaxation = p.relaxation; md.smoothing_range = p.smoothing_range;
    md.eig_n_iterations = p.eig_n_iterations; md.variable = p.variable; md.restrict_is_transpose_prolongate = p.restrict_is_transpose_prolongate;
    md.coarse_grid_maxiter = p.coarse_grid_smoother_type == "Smoother" ? 0 : 10; md.coarse_grid_abstol = 1e-20; // parameters.h:25-26
    stfem_mg_t mg = nullptr;
    check(stfem_mg_create(ctx, &md, &mg));
    // ---- fine operators (tp_01.cc:121-168)
    stfem_mesh_t        fmesh = meshes[refinement];
    const TimeWeights   w1 = fe_time_weights(type, fe_degree, tau, 1), w = fe_time_weights(type, fe_degree, tau, nts);
    std::vector<double> zero(nb, 0.0);
    stfem_op_t          matrix = nullptr, rhs_matrix = nullptr, rhs_matrix_v = nullptr;
    WaveWeights         ww;
    if (wave)
      {
        ww           = fe_time_weights_wave(type, w1, nts);
        matrix       = make_op(fmesh, k, STFEM_F64, nb, nb, ww.lhs_uK.data(), ww.lhs_uM.data());
        rhs_matrix   = make_op(fmesh, k, STFEM_F64, nb, 1, ww.rhs_uK.data(), ww.rhs_uM.data());
        rhs_matrix_v = make_op(fmesh, k, STFEM_F64, nb, 1, zero.data(), ww.rhs_vM.data());
      }
    else
      {
        matrix     = make_op(fmesh, k, STFEM_F64, nb, nb, w.A.data(), w.B.data());
        rhs_matrix = make_op(fmesh, k, STFEM_F64, nb, 1, cgp ? w.G.data() : zero.data(), cgp ? w.Z.data() : w.G.data());
      }
    stfem_ti_desc td;
    std::memset(&td, 0, sizeof(td));
    td.time_type = type; td.time_degree = fe_degree; td.n_timesteps_at_once = nts; td.problem = wave ? 2 : 1;
    td.Alpha_1 = w1.A.data(); td.Beta_1 = w1.B.data(); td.Gamma_1 = w1.G.data(); td.Zeta_1 = w1.Z.data();
    td.matrix = matrix; td.preconditioner = mg; td.rhs_matrix = rhs_matrix; td.rhs_matrix_v = rhs_matrix_v;
    td.rhs_function_id = wave ? 4 : 2; td.frequency = p.frequency; td.extrapolate = p.extrapolate;
    td.gmres_tolerance = 1e-12; td.abs_tol = 1e-12; td.max_iterations = 200; td.max_basis_size = 100; // time_integrators.h:56-59
    stfem_ti_t ti = nullptr;
    check(stfem_ti_create(&td, &ti));
    const long long N = stfem_op_n_dofs_per_block(matrix);
    Row             row{};
    {
      DeviceVectors x(ctx, nb, N), rhs(ctx, nb, N), v(ctx, wave ? nb : 0, N), prev_x(ctx, 1, N), prev_v(ctx, wave ? 1 : 0, N);
      check(stfem_interpolate(fmesh, k, 1, p.frequency, 0.0, x.p[nb - 1])); // tp_01.cc:551
      if (wave) check(stfem_interpolate(fmesh, k, 3, p.frequency, 0.0, v.p[nb - 1]));
      double time = 0.0, err[3] = {0.0, -1.0, 0.0};
      int    total = 0, solves = 0;
      while (time < p.end_time) // tp_01.cc:646-685
        {
          check(stfem_dev_copy(ctx, prev_x.p[0], x.p[nb - 1], (size_t)N * 8));
          int it = 0;
          if (!wave)
            check(stfem_ti_solve_heat(ti, x.p.data(), prev_x.p[0], rhs.p.data(), time, tau, &it));
          else
            {
              check(stfem_dev_copy(ctx, prev_v.p[0], v.p[nb - 1], (size_t)N * 8));
              check(stfem_ti_solve_wave(ti, x.p.data(), v.p.data(), rhs.p.data(), prev_x.p[0], prev_v.p[0], time, tau, &it));
            }
          total += it;
          ++solves;
          check(stfem_evaluate_error(fmesh, k, type, r, nts, x.p.data(), prev_x.p[0], time, tau, p.frequency, r + 1, err));
          time += nts * tau;
          if (max_steps > 0 && solves >= max_steps) break;
        }
      row.cells = 1;
      row.s_dofs = N;
      for (int a = 0; a < dim; ++a) row.cells *= n_cells[a];
      row.t_dofs = nb; row.iterations = total; row.timesteps = solves; row.n_levels = (int)mg_type_level.size();
      row.linf = err[1]; row.l2 = std::sqrt(err[0]); row.h1 = std::sqrt(err[2]);
    }
    stfem_ti_destroy(ti);
    stfem_mg_destroy(mg);
    for (stfem_op_t o : {matrix, rhs_matrix, rhs_matrix_v})
      if (o) stfem_op_destroy(o);
    for (stfem_op_t o : level_ops) stfem_op_destroy(o);
    for (auto &m : meshes) stfem_mesh_destroy(m.second);
    return row;
  }

  // ------------------------------------------------------------------ deal.II TableHandler / ConvergenceTable text output
  struct TextTable
  {
    struct Column { std::string key; std::vector<std::string> entries; std::vector<double> values; bool rate = false; };
    std::vector<Column> cols;
    Column &col(const std::string &key)
    {
      for (auto &c : cols)
        if (c.key == key) return c;
      cols.push_back(Column{key, {}, {}, false});
      return cols.back();
    }
    void add(const std::string &key, long long v) { col(key).entries.push_back(std::to_string(v)); }
    void add(const std::string &key, double v, int precision, bool scientific)
    {
      char buf[64];
      std::snprintf(buf, sizeof(buf), scientific ? "%.*e" : "%.*f", precision, v);
      Column &c = col(key);
      c.entries.push_back(std::isnan(v) ? "nan" : buf);
      c.values.push_back(v);
    }
    std::string write_text() const
    {
      std::string                                head;
      std::vector<std::vector<std::vector<std::string>>> sub;
      std::vector<std::vector<size_t>>                   widths;
      size_t                                             n_rows = 0;
      for (const Column &c : cols)
        {
          std::vector<std::vector<std::string>> s{c.entries};
          if (c.rate)
            {
              std::vector<std::string> rate{"-"};
              for (size_t i = 1; i < c.values.size(); ++i)
                {
                  const double q = c.values[i] != 0 ? c.values[i - 1] / c.values[i] : NAN;
                  char         buf[32];
                  std::snprintf(buf, sizeof(buf), "%.2f", std::log2(q));
                  rate.push_back((std::isnan(q) || q <= 0) ? "nan" : buf);
                }
              s.push_back(rate);
            }
          std::vector<size_t> w;
          size_t              total = 0;
          for (auto &colv : s)
            {
              size_t m = 0;
              for (auto &e : colv) m = std::max(m, e.size());
              w.push_back(m);
              total += m;
            }
          total += w.size() - 1;
          const size_t klen = c.key.size(); // bytes, as std::string counts them
          if (total < klen) { w[0] += klen - total; total = klen; }
          const size_t front = (total - klen) / 2;
          head += std::string(front, ' ') + c.key + std::string(total - klen - front, ' ') + " ";
          n_rows = std::max(n_rows, c.entries.size());
          sub.push_back(s);
          widths.push_back(w);
        }
      std::string out = head + "\n";
      for (size_t r = 0; r < n_rows; ++r)
        {
          for (size_t c = 0; c < sub.size(); ++c)
            for (size_t q = 0; q < sub[c].size(); ++q)
              {
                const std::string e = r < sub[c][q].size() ? sub[c][q][r] : "";
                out += std::string(widths[c][q] - e.size(), ' ') + e + " ";
              }
          out += "\n";
        }
      return out;
    }
  };
} // namespace

int main(int argc, char **argv)
{
  std::string file;
  int         dim = 2, max_steps = 0, number_type = STFEM_F32;
  for (int i = 1; i < argc; ++i)
    {
      const std::string a = argv[i];
      if ((a == "--file" || a == "-f") && i + 1 < argc) file = argv[++i];
      else if ((a == "--dim" || a == "-d") && i + 1 < argc) dim = std::atoi(argv[++i]);
      else if (a == "--max_steps" && i + 1 < argc) max_steps = std::atoi(argv[++i]);
      else if (a == "--precondition_double") number_type = STFEM_F64;
    }
  if (file.empty())
    {
      std::fprintf(stderr, "usage: tp01_main --file parameters.json [--dim 2|3] [--precondition_double] [--max_steps n]\n");
      return 2;
    }
  try
    {
      const Parameters p(read_flat_json(file), dim);
      if (!p.space_time_convergence_test) throw std::runtime_error("tp01_main drives spaceTimeConvergenceTest = true only (practical runs: driver.py)");
      stfem_ctx_t ctx = nullptr;
      check(stfem_ctx_create(0, &ctx));
      TextTable itable;
      for (int k = p.fe_degree; k < p.fe_degree + p.n_deg_cycles; ++k)
        {
          TextTable table;
          itable.add("k \\ r", (long long)k);
          for (int rf = p.refinement; rf < p.refinement + p.n_ref_cycles; ++rf)
            {
              const Row row = convergence_test(ctx, p, dim, rf, k, number_type, max_steps);
              const double avg = (double)row.iterations / row.timesteps;
              std::printf(":: Number of active cells: %lld\n:: Number of degrees of freedom: %lld\n:: Min Level 0  Max Level %d\n"
                          "Average GMRES iterations %g (%d gmres_iterations / %d timesteps)\n\n",
                          row.cells, row.s_dofs, row.n_levels, avg, row.iterations, row.timesteps);
              table.add("cells", row.cells);
              table.add("s-dofs", row.s_dofs);
              table.add("t-dofs", (long long)row.t_dofs);
              table.add("st-dofs", (long long)row.timesteps * row.s_dofs * row.t_dofs);
              table.add("work", row.s_dofs * row.t_dofs * row.iterations); // tp_01.cc:715
              table.add("L\xe2\x88\x9e-L\xe2\x88\x9e", row.linf, 5, true);
              table.add("L2-L2", row.l2, 5, true);
              table.add("L2-H1_semi", row.h1, 5, true);
              itable.add(std::to_string(rf), avg, 4, false);
            }
          for (const char *key : {"L\xe2\x88\x9e-L\xe2\x88\x9e", "L2-L2", "L2-H1_semi"}) table.col(key).rate = true;
          std::printf("Convergence table k=%d\n%s\n", k, table.write_text().c_str());
        }
      std::printf("Iteration count table\n%s\n", itable.write_text().c_str());
      check(stfem_ctx_destroy(ctx));
    }
  catch (const std::exception &e)
    {
      std::fprintf(stderr, "tp01_main: %s\n", e.what());
      return 1;
    }
  return 0;
}
