// Host-side time-discretisation algebra (product code, C++): the small dense matrices that become
// kernel constants.  Same outputs as the reference's include/fe_time.h / fe_time.cc:
//   get_cg_weights / get_dg_weights     fe_time.h:643-744
//   split_lhs_rhs                       fe_time.h:485-514
//   get_fe_time_weights (multi-step)    fe_time.h:351-409, per level :411-442
//   get_fe_time_weights_wave            fe_time.h:157-305, per level :444-474
//   time projection/prolongation/restriction matrices   fe_time.h:749-898
//   get_poly_mg_sequence / get_mg_sequence / get_precondition_stmg_types   fe_time.cc:16-150
//   BlockSlice indexing                 fe_time.h:901-1017 ; get_blk_indices  stmg.h:460-501
#pragma once
#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "basis_host.hpp"

namespace stfem
{
  struct Mat
  {
    int                 m = 0, n = 0;
    std::vector<double> a;
    Mat() = default;
    Mat(int m_, int n_) : m(m_), n(n_), a((size_t)m_ * n_, 0.0) {}
    double       &operator()(int i, int j) { return a[(size_t)i * n + j]; }
    const double &operator()(int i, int j) const { return a[(size_t)i * n + j]; }
    Mat operator*(const Mat &B) const
    {
      Mat C(m, B.n);
      for (int i = 0; i < m; ++i)
        for (int k = 0; k < n; ++k)
          for (int j = 0; j < B.n; ++j) C(i, j) += (*this)(i, k) * B(k, j);
      return C;
    }
    Mat transposed() const
    {
      Mat T(n, m);
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) T(j, i) = (*this)(i, j);
      return T;
    }
    void scale(double s)
    {
      for (auto &v : a) v *= s;
    }
    bool all_zero() const
    {
      for (auto v : a)
        if (v != 0.0) return false;
      return true;
    }
  };

  // inverse by Gauss-Jordan with partial pivoting (FullMatrix::gauss_jordan / invert)
  inline Mat inverse(const Mat &A)
  {
    const int n = A.m;
    Mat       W = A, I(n, n);
    for (int i = 0; i < n; ++i) I(i, i) = 1;
    for (int c = 0; c < n; ++c)
      {
        int p = c;
        for (int r = c + 1; r < n; ++r)
          if (std::fabs(W(r, c)) > std::fabs(W(p, c))) p = r;
        if (W(p, c) == 0.0) throw std::runtime_error("singular matrix");
        if (p != c)
          for (int j = 0; j < n; ++j)
            {
              std::swap(W(p, j), W(c, j));
              std::swap(I(p, j), I(c, j));
            }
        const double d = 1.0 / W(c, c);
        for (int j = 0; j < n; ++j)
          {
            W(c, j) *= d;
            I(c, j) *= d;
          }
        for (int r = 0; r < n; ++r)
          if (r != c && W(r, c) != 0.0)
            {
              const double f = W(r, c);
              for (int j = 0; j < n; ++j)
                {
                  W(r, j) -= f * W(c, j);
                  I(r, j) -= f * I(c, j);
                }
            }
      }
    return I;
  }

  // FullMatrix::fill(src, dst_offset_i, dst_offset_j, src_offset_i, src_offset_j)
  inline void fill(Mat &dst, const Mat &src, int di, int dj, int si, int sj)
  {
    const int rows = std::min(dst.m - di, src.m - si), cols = std::min(dst.n - dj, src.n - sj);
    for (int i = 0; i < rows; ++i)
      for (int j = 0; j < cols; ++j) dst(di + i, dj + j) = src(si + i, sj + j);
  }

  enum TimeType { CGP = 1, DG = 2 };

  inline std::vector<double> time_nodes(int type, int r)
  {
    return type == DG ? gauss_radau_right(r + 1).x : gauss_lobatto(r + 1).x;
  }

  // {matrix, matrix_der} (CGP, r x (r+1)) or {mass, der+jump, jump} (DG)
  inline std::vector<Mat> raw_weights(int type, int r)
  {
    const Rule q = gauss(r + 2);
    if (type == CGP)
      {
        const auto          trial = gauss_lobatto(r + 1).x;
        std::vector<double> test(trial.begin() + 1, trial.end());
        Mat                 M(r, r + 1), Md(r, r + 1);
        for (int i = 0; i < r; ++i)
          for (int j = 0; j < r + 1; ++j)
            for (size_t k = 0; k < q.x.size(); ++k)
              {
                const double ti = lagrange_value(test, i, q.x[k]);
                M(i, j) += q.w[k] * ti * lagrange_value(trial, j, q.x[k]);
                Md(i, j) += q.w[k] * ti * lagrange_deriv(trial, j, q.x[k]);
              }
        return {M, Md};
      }
    const auto nodes = gauss_radau_right(r + 1).x;
    Mat        M(r + 1, r + 1), Md(r + 1, r + 1), J(r + 1, 1);
    for (int i = 0; i < r + 1; ++i)
      {
        J(i, 0) = lagrange_value(nodes, i, 0.0);
        for (int j = 0; j < r + 1; ++j)
          {
            Md(i, j) += lagrange_value(nodes, i, 0.0) * lagrange_value(nodes, j, 0.0);
            for (size_t k = 0; k < q.x.size(); ++k)
              {
                const double vi = lagrange_value(nodes, i, q.x[k]);
                M(i, j) += q.w[k] * vi * lagrange_value(nodes, j, q.x[k]);
                Md(i, j) += q.w[k] * vi * lagrange_deriv(nodes, j, q.x[k]);
              }
          }
      }
    return {M, Md, J};
  }

  // [Alpha*tau, Beta, Gamma, Zeta] of the all-at-once system for n_timesteps_at_once steps
  inline std::vector<Mat> fe_time_weights(int type, int r, double tau, int nts)
  {
    std::vector<Mat> raw = raw_weights(type, r), tmp(4);
    if (type == CGP)
      {
        const int n = raw[0].n;
        tmp[0] = Mat(r, n - 1); tmp[1] = Mat(r, n - 1); tmp[2] = Mat(r, 1); tmp[3] = Mat(r, 1);
        fill(tmp[0], raw[0], 0, 0, 0, 1);
        fill(tmp[1], raw[1], 0, 0, 0, 1);
        fill(tmp[2], raw[0], 0, 0, 0, 0);
        fill(tmp[3], raw[1], 0, 0, 0, 0);
        tmp[2].scale(-tau);
        tmp[3].scale(-1.0);
      }
    else
      {
        tmp[0] = raw[0]; tmp[1] = raw[1];
        tmp[3] = raw[2];
        tmp[2] = Mat(raw[2].m, 1);
      }
    tmp[0].scale(tau);
    const int        nd = tmp[0].m, nt = nd * nts;
    std::vector<Mat> ret = {Mat(nt, nt), Mat(nt, nt), Mat(nt, 1), Mat(nt, 1)};
    for (int it = 0; it < nts; ++it)
      for (int i = 0; i < nd; ++i)
        {
          if (it < nts - 1 && i == nd - 1)
            for (int j = 0; j < nd; ++j)
              {
                ret[0](j + (it + 1) * nd, i + it * nd) = -tmp[2](j, 0);
                ret[1](j + (it + 1) * nd, i + it * nd) = -tmp[3](j, 0);
              }
          for (int j = 0; j < nd; ++j)
            {
              ret[0](i + it * nd, j + it * nd) = tmp[0](i, j);
              ret[1](i + it * nd, j + it * nd) = tmp[1](i, j);
            }
        }
    for (int i = 0; i < nd; ++i)
      {
        ret[2](i, 0) = tmp[type == CGP ? 2 : 3](i, 0);
        ret[3](i, 0) = tmp[type == CGP ? 3 : 2](i, 0);
      }
    return ret;
  }

  // [lhs_uK, lhs_uM, rhs_uK, rhs_uM, rhs_vM]
  inline std::vector<Mat> fe_time_weights_wave(int type, const Mat &Alpha, const Mat &Beta, const Mat &Gamma,
                                               const Mat &Zeta, int nts)
  {
    const Mat    Ainv = inverse(Alpha);
    const Mat    BAB = Beta * Ainv * Beta, BAG = Beta * Ainv * Gamma;
    const int    m = Gamma.m, nd = Alpha.m, nt = nd * nts;
    const double amm = Alpha(m - 1, m - 1), gxai = Gamma(m - 1, 0) / amm;
    Mat          GAG = Gamma;
    GAG.scale(gxai);
    Mat Brow(1, Beta.n);
    for (int j = 0; j < Beta.n; ++j) Brow(0, j) = Beta(m - 1, j);
    Mat GAB = Gamma * Brow;
    GAB.scale(1.0 / amm);
    std::vector<Mat> ret = {Mat(nt, nt), Mat(nt, nt), Mat(nt, 1), Mat(nt, 1), Mat(nt, 1)};
    if (type == CGP)
      {
        const Mat BAZ = Beta * Ainv * Zeta;
        Mat       ZmG = Zeta;
        for (int i = 0; i < m; ++i) ZmG(i, 0) -= BAG(i, 0);
        Mat ZmB = ZmG * Brow;
        ZmB.scale(1.0 / amm);
        const double zxai = Zeta(m - 1, 0) / amm;
        for (int it = 0; it < nts; ++it)
          for (int jt = 0; jt <= it; ++jt)
            for (int i = 0; i < nd; ++i)
              {
                if (it == 0 && jt == 0)
                  {
                    ret[2](i, 0) = Gamma(i, 0);
                    ret[3](i, 0) = BAZ(i, 0);
                    ret[4](i, 0) = ZmG(i, 0);
                  }
                else if (jt == 0)
                  {
                    ret[3](i + it * nd, 0) = -zxai * std::pow(gxai, it - 1) * ZmG(i, 0);
                    ret[4](i + it * nd, 0) = std::pow(gxai, it) * ZmG(i, 0);
                  }
                if (it == jt + 1)
                  {
                    ret[0](i + it * nd, nd - 1 + jt * nd) = -Gamma(i, 0);
                    ret[1](i + it * nd, nd - 1 + jt * nd) = -BAZ(i, 0);
                  }
                if (it == jt)
                  for (int j = 0; j < nd; ++j)
                    {
                      ret[0](i + it * nd, j + it * nd) = Alpha(i, j);
                      ret[1](i + it * nd, j + it * nd) = BAB(i, j);
                    }
                else
                  for (int j = 0; j < nd; ++j)
                    ret[1](i + it * nd, j + jt * nd) +=
                      -std::pow(gxai, it - jt - 1) * ZmB(i, j) +
                      ((it > 1 && it - 1 > jt && j == nd - 1) ? std::pow(gxai, it - jt - 2) * zxai * ZmG(i, 0) : 0.0);
              }
      }
    else
      {
        for (int it = 0; it < nts; ++it)
          for (int i = 0; i < nd; ++i)
            {
              if (it == 0)
                {
                  ret[3](i, 0) = BAG(i, 0);
                  ret[4](i, 0) = Gamma(i, 0);
                }
              if (it == 1) ret[3](nd + i, 0) = -GAG(i, 0);
              if (it < nts - 1)
                for (int j = 0; j < nd; ++j)
                  ret[1](j + (it + 1) * nd, i + it * nd) = -GAB(j, i) - (i == nd - 1 ? BAG(j, 0) : 0.0);
              if (it < nts - 2 && i == nd - 1)
                for (int j = 0; j < nd; ++j) ret[1](j + (it + 2) * nd, i + it * nd) = GAG(j, 0);
              for (int j = 0; j < nd; ++j)
                {
                  ret[0](i + it * nd, j + it * nd) = Alpha(i, j);
                  ret[1](i + it * nd, j + it * nd) = BAB(i, j);
                }
            }
      }
    return ret;
  }

  // ---- time transfer matrices
  inline Mat l2_projection(const std::vector<double> &src, const std::vector<double> &dst)
  {
    const Rule q = gauss((int)std::max(src.size(), dst.size()) + 1);
    const int  ns = (int)src.size(), nd = (int)dst.size();
    Mat        M(nd, nd), B(nd, ns);
    for (size_t k = 0; k < q.x.size(); ++k)
      for (int i = 0; i < nd; ++i)
        {
          const double vi = lagrange_value(dst, i, q.x[k]);
          for (int j = 0; j < nd; ++j) M(i, j) += q.w[k] * vi * lagrange_value(dst, j, q.x[k]);
          for (int j = 0; j < ns; ++j) B(i, j) += q.w[k] * vi * lagrange_value(src, j, q.x[k]);
        }
    return inverse(M) * B;
  }

  inline Mat time_projection_matrix(int type, int r_src, int r_dst, int nts)
  {
    const int nd_dst = type == DG ? r_dst + 1 : r_dst, nd_src = type == DG ? r_src + 1 : r_src;
    const int n_dst = type == DG ? nts * (r_dst + 1) : nts * r_dst + 1;
    const int n_src = type == DG ? nts * (r_src + 1) : nts * r_src + 1;
    const Mat proj = l2_projection(time_nodes(type, r_src), time_nodes(type, r_dst));
    Mat       pn(n_dst, n_src);
    for (int it = 0; it < nts; ++it) fill(pn, proj, it * nd_dst, it * nd_src, 0, 0);
    if (type == CGP)
      {
        Mat out(n_dst - 1, n_src - 1);
        fill(out, pn, 0, 0, 1, 1);
        return out;
      }
    return pn;
  }

  inline Mat child_embedding(const std::vector<double> &nodes, int child)
  {
    const int n = (int)nodes.size();
    Mat       P(n, n);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) P(i, j) = lagrange_value(nodes, j, 0.5 * (nodes[i] + child));
    return P;
  }

  inline Mat dg_child_restriction(const std::vector<double> &nodes, int child)
  {
    const int  n = (int)nodes.size();
    const Rule q = gauss(n + 1);
    Mat        M(n, n), B(n, n);
    for (size_t k = 0; k < q.x.size(); ++k)
      for (int i = 0; i < n; ++i)
        {
          const double vi = lagrange_value(nodes, i, q.x[k]);
          const double ci = lagrange_value(nodes, i, 0.5 * (q.x[k] + child));
          for (int j = 0; j < n; ++j)
            {
              M(i, j) += q.w[k] * vi * lagrange_value(nodes, j, q.x[k]);
              B(i, j) += 0.5 * q.w[k] * ci * lagrange_value(nodes, j, q.x[k]);
            }
        }
    return inverse(M) * B;
  }

  inline Mat q_child_restriction(const std::vector<double> &nodes, int child)
  {
    const int n = (int)nodes.size();
    Mat       R(n, n);
    for (int i = 0; i < n; ++i)
      {
        const double xc = 2.0 * nodes[i] - child;
        if (xc >= -1e-12 && xc <= 1.0 + 1e-12)
          for (int j = 0; j < n; ++j)
            {
              const double v = lagrange_value(nodes, j, xc);
              R(i, j)        = std::fabs(v) < 1e-14 ? 0.0 : v;
            }
      }
    return R;
  }

  inline Mat time_prolongation_matrix(int type, int r, int nts)
  {
    const auto nodes = time_nodes(type, r);
    const Mat  left = child_embedding(nodes, 0), right = child_embedding(nodes, 1);
    Mat        prol;
    int        nd;
    if (type == DG)
      {
        prol = Mat(2 * (r + 1), r + 1);
        fill(prol, left, 0, 0, 0, 0);
        fill(prol, right, r + 1, 0, 0, 0);
        nd = r + 1;
      }
    else
      {
        prol = Mat(2 * r, r);
        fill(prol, left, 0, 0, 1, 1);
        fill(prol, right, r, 0, 1, 1);
        nd = r;
      }
    Mat pn(nd * nts, nd * nts / 2);
    for (int it = 0; it < nts / 2; ++it) fill(pn, prol, it * 2 * nd, it * nd, 0, 0);
    return pn;
  }

  inline Mat time_restriction_matrix(int type, int r, int nts)
  {
    const auto nodes = time_nodes(type, r);
    Mat        rest;
    int        nd;
    if (type == DG)
      {
        rest = Mat(r + 1, 2 * (r + 1));
        fill(rest, dg_child_restriction(nodes, 0), 0, 0, 0, 0);
        fill(rest, dg_child_restriction(nodes, 1), 0, r + 1, 0, 0);
        nd = r + 1;
      }
    else
      {
        rest = Mat(r, 2 * r);
        fill(rest, q_child_restriction(nodes, 0), 0, 0, 1, 1);
        fill(rest, q_child_restriction(nodes, 1), 0, r, 1, 1);
        nd = r;
      }
    Mat rn(nd * nts / 2, nd * nts);
    for (int it = 0; it < nts / 2; ++it) fill(rn, rest, it * nd, it * 2 * nd, 0, 0);
    return rn;
  }

  // ---- multigrid level sequences (level types 't','k','h','p', coarse -> fine)
  inline int next_poly_degree(int prev, int p_sequence /*0 bisect, 1 decrease_by_one, 2 go_to_one*/, int k_min)
  {
    if (p_sequence == 0) return std::max(prev / 2, 0);
    if (p_sequence == 1) return std::max(prev - 1, 0);
    return k_min;
  }

  inline std::vector<int> poly_mg_sequence(int k_max, int k_min, int p_sequence)
  {
    std::vector<int> d{k_max};
    if (k_max == k_min) return d;
    while (d.back() > k_min) d.push_back(next_poly_degree(d.back(), p_sequence, k_min));
    std::reverse(d.begin(), d.end());
    return d;
  }

  inline std::string mg_sequence(int n_sp_lvl, int n_k_seq, int n_p_seq, int nts, int nts_min, char lower_lvl,
                                 int coarsening_type /*0 space_or_time, 1 space_and_time*/, bool time_before_space,
                                 bool use_pmg, bool zip_from_back)
  {
    const int  n_k_lvl = n_k_seq - 1;
    int        n_t_lvl = 0;
    for (int v = nts / std::max(nts_min, 1); v > 1; v /= 2) ++n_t_lvl;
    const char upper_lvl = lower_lvl == 'k' ? 't' : 'k';
    const char lower_s = lower_lvl == 'k' ? 'p' : 'h', upper_s = lower_lvl == 'k' ? 'h' : 'p';
    const int  n_ll = lower_lvl == 'k' ? n_k_lvl : n_t_lvl, n_ul = lower_lvl == 'k' ? n_t_lvl : n_k_lvl;
    const int  n_p_lvl = use_pmg ? n_p_seq - 1 : 0;
    const int  n_ll_s = lower_lvl == 'k' ? n_p_lvl : n_sp_lvl - 1, n_ul_s = lower_lvl == 'k' ? n_sp_lvl - 1 : n_p_lvl;
    std::string time_levels = std::string(n_ll, lower_lvl) + std::string(n_ul, upper_lvl);
    std::string space_levels = std::string(n_ll_s, lower_s) + std::string(n_ul_s, upper_s);
    const std::string &first = time_before_space ? time_levels : space_levels;
    const std::string &second = time_before_space ? space_levels : time_levels;
    std::string out;
    if (coarsening_type == 0)
      {
        if (zip_from_back)
          out = std::string(first.rbegin(), first.rend()) + std::string(second.rbegin(), second.rend());
        else
          out = first + second;
      }
    else
      {
        const size_t mx = std::max(first.size(), second.size());
        for (size_t i = 0; i < mx; ++i)
          {
            if (i < first.size()) out.push_back(zip_from_back ? first[first.size() - 1 - i] : first[i]);
            if (i < second.size()) out.push_back(zip_from_back ? second[second.size() - 1 - i] : second[i]);
          }
        if (zip_from_back) std::reverse(out.begin(), out.end());
      }
    return out;
  }

  inline bool is_space_lvl(char c) { return c == 'h' || c == 'p'; }
  inline bool is_time_lvl(char c) { return c == 't' || c == 'k'; }

  inline std::vector<int> precondition_stmg_types(const std::string &seq, int coarsening_type, bool time_before_space,
                                                  int smoother)
  {
    std::vector<int> ret(seq.size() + 1, smoother);
    if (coarsening_type == 0) return ret;
    for (size_t i = 0; i + 1 < seq.size(); ++i)
      {
        const bool hit = time_before_space ? (is_space_lvl(seq[i]) && is_time_lvl(seq[i + 1])) :
                                             (is_time_lvl(seq[i]) && is_space_lvl(seq[i + 1]));
        if (hit)
          {
            ret[i]     = smoother;
            ret[i + 1] = 0;
            ++i;
          }
      }
    return ret;
  }

  struct BlockSlice
  {
    int nt = 1, nv = 1, nd = 1;
    int n_blocks() const { return nt * nv * nd; }
    int index(int t, int v, int d) const { return t * (nv * nd) + v * nd + d; } // variable-major
  };
} // namespace stfem
