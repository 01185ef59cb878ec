// C ABI: host-side problem data of the reference's "practical" configurations (BASELINE configs[3]): the
// heterogeneous coefficient Coefficient<dim> (reference include/operators.h:870-965) sampled at the quadrature points
// of a level mesh (MatrixFreeOperator::evaluate_coefficient, operators.h:1060-1087) and the initial value
// Functions::CutOffFunctionCinfty around `sourcePoint` (tests/tp_01.cc:374-381, 551).  Set-up work done once per level:
// plain host C++, no GPU; the results are handed to stfem_op_create (laplace_coeff_q) / uploaded as a vector.
#include <cmath>
#include <random>

#include "basis_host.hpp"
#include "common.hpp"

using namespace stfem;

namespace
{
  // MappingQ1 image of the reference point xi of cell c (vertices lexicographic, x fastest; null = Cartesian box)
  inline void map_point(int dim, const int *n, const double *lower, const double *h, const double *vertices, const int *c,
                        const double *xi, double *x)
  {
    if (!vertices)
      {
        for (int a = 0; a < dim; ++a) x[a] = lower[a] + h[a] * (c[a] + xi[a]);
        return;
      }
    for (int a = 0; a < dim; ++a) x[a] = 0.0;
    for (int v = 0; v < (1 << dim); ++v)
      {
        const int vb[3] = {v & 1, (v >> 1) & 1, (v >> 2) & 1};
        long long vid   = (long long)(c[0] + vb[0]) + (long long)(n[0] + 1) * (c[1] + vb[1]);
        if (dim == 3) vid += (long long)(n[0] + 1) * (n[1] + 1) * (c[2] + vb[2]);
        double sh = 1.0;
        for (int a = 0; a < dim; ++a) sh *= vb[a] ? xi[a] : 1.0 - xi[a];
        for (int a = 0; a < dim; ++a) x[a] += vertices[vid * dim + a] * sh;
      }
  }

  int check_mesh(const char *who, int dim, const int *n_cells, const double *lower, const double *upper)
  {
    STFEM_REQUIRE(dim == 2 || dim == 3, "%s: dim must be 2 or 3", who);
    STFEM_REQUIRE(n_cells && lower && upper, "%s: null mesh description", who);
    for (int a = 0; a < dim; ++a)
      STFEM_REQUIRE(n_cells[a] >= 1 && upper[a] > lower[a], "%s: empty mesh in direction %d", who, a);
    return STFEM_OK;
  }
} // namespace

extern "C" {

int stfem_coefficient_distortion(int dim, const int *subdivisions, double distort_coeff, double *table)
{
  STFEM_REQUIRE(dim == 2 || dim == 3, "stfem_coefficient_distortion: dim must be 2 or 3");
  STFEM_REQUIRE(subdivisions && table, "stfem_coefficient_distortion: null argument");
  long long n = 1;
  for (int a = 0; a < dim; ++a)
    {
      STFEM_REQUIRE(subdivisions[a] >= 1, "stfem_coefficient_distortion: subdivisions[%d] < 1", a);
      n *= subdivisions[a];
    }
  // boost::random::mt19937(default_seed) is the standard MT19937 seeded with 5489; boost's
  // uniform_real_distribution<double> consumes ONE 32-bit draw per value: u / 2^32 * (b - a) + a
  std::mt19937 rng(5489u);
  const double a = 1.0 - distort_coeff, b = 1.0 + distort_coeff;
  for (long long i = 0; i < n; ++i) table[i] = (double)rng() / 4294967296.0 * (b - a) + a;
  return STFEM_OK;
}

int stfem_coefficient_at_qpoints(int dim, const int *n_cells, const double *lower, const double *upper, const double *vertices,
                                 int degree, const int *subdivisions, const double *coeff_lower, const double *coeff_upper,
                                 double distort_coeff, const double *c123, double *out)
{
  STFEM_FORWARD(check_mesh("stfem_coefficient_at_qpoints", dim, n_cells, lower, upper));
  STFEM_REQUIRE(degree >= 1 && degree <= 7, "stfem_coefficient_at_qpoints: degree %d out of range", degree);
  STFEM_REQUIRE(out, "stfem_coefficient_at_qpoints: null output");
  const double c1 = c123 ? c123[0] : 1.0, c2 = c123 ? c123[1] : 9.0, c3 = c123 ? c123[2] : 16.0;
  std::vector<double> table;
  double              step[3] = {1, 1, 1};
  int                 sub[3]  = {1, 1, 1};
  const bool          distorted = distort_coeff != 0.0;
  if (distorted)
    {
      STFEM_REQUIRE(subdivisions && coeff_lower && coeff_upper, "stfem_coefficient_at_qpoints: distortion needs subdivisions and the box");
      long long nt = 1;
      for (int a = 0; a < dim; ++a)
        {
          sub[a]  = subdivisions[a];
          step[a] = (coeff_upper[a] - coeff_lower[a]) / subdivisions[a];
          nt *= subdivisions[a];
        }
      table.resize(nt);
      STFEM_FORWARD(stfem_coefficient_distortion(dim, subdivisions, distort_coeff, table.data()));
    }
  const int  nq = degree + 1;
  const Rule g  = gauss(nq);
  double     h[3] = {1, 1, 1};
  int        n[3] = {1, 1, 1};
  for (int a = 0; a < dim; ++a)
    {
      n[a] = n_cells[a];
      h[a] = (upper[a] - lower[a]) / n_cells[a];
    }
  const int nqz = dim == 3 ? nq : 1;
  long long o   = 0;
  for (int cz = 0; cz < n[2]; ++cz)
    for (int cy = 0; cy < n[1]; ++cy)
      for (int cx = 0; cx < n[0]; ++cx)
        {
          const int c[3] = {cx, cy, cz};
          for (int qz = 0; qz < nqz; ++qz)
            for (int qy = 0; qy < nq; ++qy)
              for (int qx = 0; qx < nq; ++qx)
                {
                  const double xi[3] = {g.x[qx], g.x[qy], dim == 3 ? g.x[qz] : 0.0};
                  double       x[3];
                  map_point(dim, n, lower, h, vertices, c, xi, x);
                  double v = x[1] >= 0.2 ? (x[0] < 0.2 ? c2 : c3) : c1;
                  if (distorted)
                    {
                      // Table<dim,double>(s0, s1(, s2)) filled in C order: last index fastest (operators.h:905-921);
                      // static_cast<unsigned> of the non-negative quotient truncates (operators.h:946-961)
                      long long idx = 0;
                      for (int a = 0; a < dim; ++a)
                        {
                          long long i = (long long)((x[a] - coeff_lower[a]) / step[a]);
                          i           = i < 0 ? 0 : (i >= sub[a] ? sub[a] - 1 : i);
                          idx         = idx * sub[a] + i;
                        }
                      v *= table[idx];
                    }
                  out[o++] = v;
                }
        }
  return STFEM_OK;
}

int stfem_cutoff_cinfty_interpolate(int dim, const int *n_cells, const double *lower, const double *upper, const double *vertices,
                                    int degree, double radius, const double *center, int integrate_to_one, double *out)
{
  STFEM_FORWARD(check_mesh("stfem_cutoff_cinfty_interpolate", dim, n_cells, lower, upper));
  STFEM_REQUIRE(degree >= 1 && degree <= 7, "stfem_cutoff_cinfty_interpolate: degree %d out of range", degree);
  STFEM_REQUIRE(radius > 0 && center && out, "stfem_cutoff_cinfty_interpolate: bad radius / null argument");
  // unit-ball integrals of e exp(-1/(1-|x|^2)) in 1, 2, 3 dimensions (deal.II integral_Cinfty; re-derived by quadrature
  // in tests/test_problem_host.py)
  static const double unit_integral[3] = {1.20690032243787617533623799633, 1.26811216112759608094632335664,
                                          1.1990039070192139033798473858};
  const double rescaling = integrate_to_one ? 1.0 / (unit_integral[dim - 1] * std::pow(radius, dim)) : 1.0;
  const Rule   gl        = gauss_lobatto(degree + 1);
  double       h[3] = {1, 1, 1};
  int          n[3] = {1, 1, 1}, np[3] = {1, 1, 1};
  for (int a = 0; a < dim; ++a)
    {
      n[a]  = n_cells[a];
      np[a] = degree * n_cells[a] + 1;
      h[a]  = (upper[a] - lower[a]) / n_cells[a];
    }
  for (int iz = 0; iz < np[2]; ++iz)
    for (int iy = 0; iy < np[1]; ++iy)
      for (int ix = 0; ix < np[0]; ++ix)
        {
          // a support point shared by several cells has the same image from each of them: take the cell to its left
          const int i[3] = {ix, iy, iz};
          int       c[3] = {0, 0, 0};
          double    xi[3] = {0, 0, 0};
          for (int a = 0; a < dim; ++a)
            {
              c[a]        = i[a] / degree;
              int loc     = i[a] - c[a] * degree;
              if (c[a] == n[a]) { c[a] = n[a] - 1; loc = degree; }
              xi[a] = gl.x[loc];
            }
          double x[3];
          map_point(dim, n, lower, h, vertices, c, xi, x);
          double d2 = 0;
          for (int a = 0; a < dim; ++a) d2 += (x[a] - center[a]) * (x[a] - center[a]);
          const double d = std::sqrt(d2);
          double       v = 0.0;
          if (d < radius)
            {
              const double e = -radius * radius / (radius * radius - d * d);
              v              = e < -50 ? 0.0 : rescaling * 2.71828182845904523536 * std::exp(e);
            }
          out[(long long)ix + (long long)np[0] * (iy + (long long)np[1] * iz)] = v;
        }
  return STFEM_OK;
}

} // extern "C"
