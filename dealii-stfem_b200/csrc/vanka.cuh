// Cell-patch additive Schwarz smoother ("Vanka" / block Jacobi).
// Replaces PreconditionVanka of the reference (include/stmg.h:745-872) and the patch extraction
// restrict_to_full_matrices_ (include/compute_block_matrix.h:50-139):
//   per cell  B_c = Beta (x) M_c + Alpha (x) K_c,  where K_c, M_c are the ASSEMBLED matrices restricted to the
//   cell's DoFs (contributions of the neighbouring cells on shared DoFs included), rows scaled by the valence
//   of the row DoF; stored inverse; apply = gather -> dense matvec -> scatter-add.
// Here the patch matrices are built matrix-free on the device (no sparse matrix is ever assembled), the
// (nb*n_c)^2 inverses come from a batched Gauss-Jordan kernel in double precision, and on Cartesian meshes
// with constant coefficients only the <= 3^dim distinct patches are stored (the reference stores one per cell).
#pragma once
#include <cuda_fp16.h>
#include <cstdlib>
#include "op.hpp"
#include "vanka_fd.cuh"
#include "vec.cuh"

namespace stfem
{
  struct VankaGeom
  {
    int       dim, n1;
    int       n[3], np[3];
    long long n_cells;
    unsigned  dirichlet;
    double    h[3];
    const double *metric; // general: per cell per q (nsym+1) doubles (coefficient folded in), else null
    const double *coeff_cell;
    double        S[49], D[49], w[7]; // 1D shape values / derivatives at Gauss points [q*n1+i], weights
  };

  __device__ inline bool vk_constrained(const VankaGeom &g, int ix, int iy, int iz)
  {
    const unsigned d = g.dirichlet;
    if ((d & 1u) && ix == 0) return true;
    if ((d & 2u) && ix == g.np[0] - 1) return true;
    if ((d & 4u) && iy == 0) return true;
    if ((d & 8u) && iy == g.np[1] - 1) return true;
    if (g.dim == 3)
      {
        if ((d & 16u) && iz == 0) return true;
        if ((d & 32u) && iz == g.np[2] - 1) return true;
      }
    return false;
  }

  // K_c'(i,j), M_c'(i,j) of one cell by quadrature
  __device__ inline void vk_cell_entry(const VankaGeom &g, long long cell, const int *li, const int *lj, double &Kij, double &Mij)
  {
    const int n1 = g.n1, dim = g.dim;
    const int nq = dim == 3 ? n1 * n1 * n1 : n1 * n1;
    const int nsym = dim * (dim + 1) / 2;
    double    k = 0, m = 0;
    const double coef = g.coeff_cell ? g.coeff_cell[cell] : 1.0;
    const double vol  = g.h[0] * g.h[1] * (dim == 3 ? g.h[2] : 1.0);
    for (int q = 0; q < nq; ++q)
      {
        const int qx = q % n1, qy = (q / n1) % n1, qz = dim == 3 ? q / (n1 * n1) : 0;
        const double six = g.S[qx * n1 + li[0]], siy = g.S[qy * n1 + li[1]], siz = dim == 3 ? g.S[qz * n1 + li[2]] : 1.0;
        const double sjx = g.S[qx * n1 + lj[0]], sjy = g.S[qy * n1 + lj[1]], sjz = dim == 3 ? g.S[qz * n1 + lj[2]] : 1.0;
        const double dix = g.D[qx * n1 + li[0]], diy = g.D[qy * n1 + li[1]], diz = dim == 3 ? g.D[qz * n1 + li[2]] : 0.0;
        const double djx = g.D[qx * n1 + lj[0]], djy = g.D[qy * n1 + lj[1]], djz = dim == 3 ? g.D[qz * n1 + lj[2]] : 0.0;
        const double gi[3] = {dix * siy * siz, six * diy * siz, six * siy * diz};
        const double gj[3] = {djx * sjy * sjz, sjx * djy * sjz, sjx * sjy * djz};
        const double vi = six * siy * siz, vj = sjx * sjy * sjz;
        if (g.metric)
          {
            const double *mt = g.metric + ((size_t)cell * nq + q) * (nsym + 1);
            m += mt[nsym] * vi * vj;
            if (dim == 2)
              k += coef * (gi[0] * (mt[0] * gj[0] + mt[1] * gj[1]) + gi[1] * (mt[1] * gj[0] + mt[2] * gj[1]));
            else
              k += coef * (gi[0] * (mt[0] * gj[0] + mt[1] * gj[1] + mt[2] * gj[2]) + gi[1] * (mt[1] * gj[0] + mt[3] * gj[1] + mt[4] * gj[2]) +
                           gi[2] * (mt[2] * gj[0] + mt[4] * gj[1] + mt[5] * gj[2]));
          }
        else
          {
            const double wq = vol * g.w[qx] * g.w[qy] * (dim == 3 ? g.w[qz] : 1.0);
            m += wq * vi * vj;
            double s = 0;
            for (int d = 0; d < dim; ++d) s += gi[d] * gj[d] / (g.h[d] * g.h[d]);
            k += coef * wq * s;
          }
      }
    Kij = k; Mij = m;
  }

  // one thread per (patch, i, j): assembled patch entry incl. neighbour contributions, valence scaling,
  // constraint handling, then the nb x nb time blocks of B (row-major double, ld = nb*nc)
  static __global__ void k_vanka_build(VankaGeom g, const long long *__restrict__ cells, int n_patches, int nb, const double *__restrict__ alpha,
                                const double *__restrict__ beta, double *__restrict__ B)
  {
    const int n1 = g.n1, dim = g.dim, k = n1 - 1;
    const int nc = dim == 3 ? n1 * n1 * n1 : n1 * n1;
    const long long total = (long long)n_patches * nc * nc;
    const int       ld    = nb * nc;
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x)
      {
        const int       pidx = (int)(gid / ((long long)nc * nc));
        const int       ij   = (int)(gid % ((long long)nc * nc));
        const int       i = ij / nc, j = ij % nc;
        const long long cell = cells[pidx];
        const int c[3] = {(int)(cell % g.n[0]), (int)((cell / g.n[0]) % g.n[1]), dim == 3 ? (int)(cell / ((long long)g.n[0] * g.n[1])) : 0};
        const int li[3] = {i % n1, (i / n1) % n1, dim == 3 ? i / (n1 * n1) : 0};
        const int lj[3] = {j % n1, (j / n1) % n1, dim == 3 ? j / (n1 * n1) : 0};
        const int gi[3] = {c[0] * k + li[0], c[1] * k + li[1], c[2] * k + li[2]};
        const int gj[3] = {c[0] * k + lj[0], c[1] * k + lj[1], c[2] * k + lj[2]};
        const bool ci = vk_constrained(g, gi[0], gi[1], gi[2]), cj = vk_constrained(g, gj[0], gj[1], gj[2]);
        double Kp = 0, Mp = 0, valence = 1;
        // allowed neighbour offsets per direction
        int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        for (int d = 0; d < dim; ++d)
          {
            if (li[d] == 0 && lj[d] == 0 && c[d] > 0) lo[d] = -1;
            if (li[d] == k && lj[d] == k && c[d] < g.n[d] - 1) hi[d] = 1;
            if ((li[d] == 0 && c[d] > 0) || (li[d] == k && c[d] < g.n[d] - 1)) valence *= 2;
          }
        if ((ci || cj) && i != j)
          {
            Kp = 0; Mp = 0;
          }
        else
          {
            for (int oz = lo[2]; oz <= hi[2]; ++oz)
              for (int oy = lo[1]; oy <= hi[1]; ++oy)
                for (int ox = lo[0]; ox <= hi[0]; ++ox)
                  {
                    const int o[3] = {ox, oy, oz};
                    int       a[3], b[3];
                    for (int d = 0; d < 3; ++d)
                      {
                        a[d] = o[d] == 0 ? li[d] : (o[d] < 0 ? k : 0);
                        b[d] = o[d] == 0 ? lj[d] : (o[d] < 0 ? k : 0);
                      }
                    const long long nc_cell = (long long)(c[0] + ox) + (long long)g.n[0] * ((c[1] + oy) + (long long)g.n[1] * (c[2] + oz));
                    double kk, mm;
                    vk_cell_entry(g, nc_cell, a, b, kk, mm);
                    if (ci) { kk = fabs(kk); mm = fabs(mm); }
                    Kp += kk; Mp += mm;
                  }
          }
        double *Bp = B + (size_t)pidx * ld * ld;
        for (int a = 0; a < nb; ++a)
          for (int b = 0; b < nb; ++b)
            Bp[(size_t)(i + a * nc) * ld + (j + b * nc)] = valence * (beta[a * nb + b] * Mp + alpha[a * nb + b] * Kp);
      }
  }

  // in-place Gauss-Jordan inverse with partial pivoting, one CTA per matrix (FullMatrix::gauss_jordan)
  static __global__ void k_batched_inverse(double *__restrict__ A, int n, int *__restrict__ perm_ws, int *__restrict__ info)
  {
    double *M    = A + (size_t)blockIdx.x * n * n;
    int    *perm = perm_ws + (size_t)blockIdx.x * n;
    __shared__ double s_val[256];
    __shared__ int    s_idx[256];
    __shared__ double s_piv;
    extern __shared__ double s_col[];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int kcol = 0; kcol < n; ++kcol)
      {
        double best = -1;
        int    bi   = kcol;
        for (int r = kcol + tid; r < n; r += nt)
          {
            const double v = fabs(M[(size_t)r * n + kcol]);
            if (v > best) { best = v; bi = r; }
          }
        s_val[tid] = best; s_idx[tid] = bi;
        __syncthreads();
        for (int s = nt / 2; s > 0; s >>= 1)
          {
            if (tid < s && s_val[tid + s] > s_val[tid]) { s_val[tid] = s_val[tid + s]; s_idx[tid] = s_idx[tid + s]; }
            __syncthreads();
          }
        const int p = s_idx[0];
        if (tid == 0)
          {
            perm[kcol] = p;
            if (s_val[0] == 0.0) atomicExch(info, 1 + (int)blockIdx.x);
          }
        if (p != kcol)
          for (int cidx = tid; cidx < n; cidx += nt)
            {
              const double t = M[(size_t)kcol * n + cidx];
              M[(size_t)kcol * n + cidx] = M[(size_t)p * n + cidx];
              M[(size_t)p * n + cidx]    = t;
            }
        __syncthreads();
        if (tid == 0)
          {
            s_piv = M[(size_t)kcol * n + kcol];
            M[(size_t)kcol * n + kcol] = 1.0;
          }
        __syncthreads();
        const double ip = 1.0 / s_piv;
        for (int cidx = tid; cidx < n; cidx += nt) M[(size_t)kcol * n + cidx] *= ip;
        __syncthreads();
        // eliminate all other rows at once: the pivot column is staged in shared memory first
        for (int r = tid; r < n; r += nt) s_col[r] = M[(size_t)r * n + kcol];
        __syncthreads();
        for (long long e = tid; e < (long long)n * n; e += nt)
          {
            const int r = (int)(e / n), cidx = (int)(e % n);
            if (r == kcol) continue;
            const double f = s_col[r];
            if (f != 0.0) M[e] = (cidx == kcol ? 0.0 : M[e]) - f * M[(size_t)kcol * n + cidx];
          }
        __syncthreads();
      }
    // undo the row interchanges as column interchanges in reverse order
    for (int kcol = n - 1; kcol >= 0; --kcol)
      {
        const int p = perm[kcol];
        if (p != kcol)
          for (int r = tid; r < n; r += nt)
            {
              const double t = M[(size_t)r * n + kcol];
              M[(size_t)r * n + kcol] = M[(size_t)r * n + p];
              M[(size_t)r * n + p]    = t;
            }
        __syncthreads();
      }
  }

  // store the inverse transposed (column-major) in the level precision: invT[c*n + r] = inv[r][c]
  template <typename T>
  __global__ void k_store_inverse_T(const double *__restrict__ A, int n, long long n_mat, T *__restrict__ out)
  {
    const long long total = n_mat * n * n;
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x)
      {
        const long long mtx = gid / ((long long)n * n);
        const int       rc  = (int)(gid % ((long long)n * n));
        const int       cidx = rc / n, r = rc % n;
        out[gid] = (T)A[(size_t)mtx * n * n + (size_t)r * n + cidx];
      }
  }

  struct VankaApplyArgs
  {
    int       dim, n1, nb;
    int       n[3], np[3];
    long long n_cells, N;
    int       dedup;        // 1: matrix index from the position class of the cell (type_to_mat), 0: one per cell
    int       type_to_mat[27];
  };

  // dst += R_c^T B_c^-1 R_c src   for every cell; one CTA per cell, one thread per patch row
  template <typename T>
  __global__ void k_vanka_apply(VankaApplyArgs a, const T *__restrict__ invT, const T *__restrict__ src, T *__restrict__ dst, T scale)
  {
    extern __shared__ __align__(16) unsigned char vk_smem[];
    T            *x  = reinterpret_cast<T *>(vk_smem);
    const int     n1 = a.n1, k = n1 - 1, dim = a.dim;
    const int     nc = dim == 3 ? n1 * n1 * n1 : n1 * n1;
    const int     nrow = a.nb * nc;
    for (long long cell = blockIdx.x; cell < a.n_cells; cell += gridDim.x)
      {
        const int c[3] = {(int)(cell % a.n[0]), (int)((cell / a.n[0]) % a.n[1]), dim == 3 ? (int)(cell / ((long long)a.n[0] * a.n[1])) : 0};
        long long mat  = cell;
        if (a.dedup)
          {
            int t = 0, mul = 1;
            for (int d = 0; d < dim; ++d)
              {
                const int cls = a.n[d] == 1 ? 0 : (c[d] == 0 ? 0 : (c[d] == a.n[d] - 1 ? 2 : 1));
                t += cls * mul;
                mul *= 3;
              }
            mat = a.type_to_mat[t];
          }
        const T *Bm = invT + (size_t)mat * nrow * nrow;
        __syncthreads();
        for (int r = threadIdx.x; r < nrow; r += blockDim.x)
          {
            const int b = r / nc, l = r % nc;
            const int li[3] = {l % n1, (l / n1) % n1, dim == 3 ? l / (n1 * n1) : 0};
            const long long gi = (long long)(c[0] * k + li[0]) + (long long)a.np[0] * ((c[1] * k + li[1]) + (long long)a.np[1] * (c[2] * k + li[2]));
            x[r] = src[(size_t)b * a.N + gi];
          }
        __syncthreads();
        for (int r = threadIdx.x; r < nrow; r += blockDim.x)
          {
            T s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            int cc = 0;
            for (; cc + 3 < nrow; cc += 4)
              {
                s0 += Bm[(size_t)cc * nrow + r] * x[cc];
                s1 += Bm[(size_t)(cc + 1) * nrow + r] * x[cc + 1];
                s2 += Bm[(size_t)(cc + 2) * nrow + r] * x[cc + 2];
                s3 += Bm[(size_t)(cc + 3) * nrow + r] * x[cc + 3];
              }
            for (; cc < nrow; ++cc) s0 += Bm[(size_t)cc * nrow + r] * x[cc];
            const T   s = (s0 + s1) + (s2 + s3);
            const int b = r / nc, l = r % nc;
            const int li[3] = {l % n1, (l / n1) % n1, dim == 3 ? l / (n1 * n1) : 0};
            const long long gi = (long long)(c[0] * k + li[0]) + (long long)a.np[0] * ((c[1] * k + li[1]) + (long long)a.np[1] * (c[2] * k + li[2]));
            atomicAdd(dst + (size_t)b * a.N + gi, scale * s);
          }
      }
  }

  // ---- FP16 storage of the dense patch inverses (opt-in: stfem_mg_desc::vanka_storage = 1 or STFEM_VANKA_HALF=1).  k_vanka_apply streams (nb n_c)^2 numbers per cell and application — it is bound by HBM on
  //      the inverses — so halving their size is the lever.  Every matrix is normalised by its largest
  //      entry (patch inverses reach 1/(h^d tau), far outside the FP16 range) and stored [column][ld] with ld even, so that
  //      a thread owns two rows and reads them as one __half2; products are accumulated in FP32.  Entry-wise relative error
  //      2^-11: the smoother changes in the fourth digit, FGMRES iteration counts stay within +-1 (tests).
  template <typename T>
  __global__ void k_vanka_to_half(int nrow, int ld, const T *__restrict__ invT, __half *__restrict__ invH, float *__restrict__ mscale)
  {
    const T     *A = invT + (size_t)blockIdx.x * nrow * nrow;
    __half      *H = invH + (size_t)blockIdx.x * nrow * ld;
    __shared__ float red[32];
    float        m = 0.f;
    for (int i = threadIdx.x; i < nrow * nrow; i += blockDim.x) m = fmaxf(m, fabsf((float)A[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32)
      {
        float v = threadIdx.x < (blockDim.x + 31) / 32 ? red[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (threadIdx.x == 0) red[0] = v;
      }
    __syncthreads();
    m               = red[0];
    const float inv = m > 0.f ? 1.f / m : 0.f;
    for (int i = threadIdx.x; i < nrow * ld; i += blockDim.x)
      {
        const int cc = i / ld, r = i - cc * ld;
        H[i]         = __float2half_rn(r < nrow ? (float)A[(size_t)cc * nrow + r] * inv : 0.f);
      }
    if (threadIdx.x == 0) mscale[blockIdx.x] = m;
  }

  // dst += scale * R_c^T B_c^-1 R_c src with the FP16 inverses: one CTA per cell, one thread per PAIR of patch rows
  template <typename T>
  __global__ void k_vanka_apply_half(VankaApplyArgs a, int ld, const __half *__restrict__ invH, const float *__restrict__ mscale,
                                     const T *__restrict__ src, T *__restrict__ dst, T scale)
  {
    extern __shared__ __align__(16) unsigned char vk_smem[];
    float    *x  = reinterpret_cast<float *>(vk_smem);
    const int n1 = a.n1, k = n1 - 1, dim = a.dim;
    const int nc = dim == 3 ? n1 * n1 * n1 : n1 * n1;
    const int nrow = a.nb * nc, ld2 = ld >> 1;
    for (long long cell = blockIdx.x; cell < a.n_cells; cell += gridDim.x)
      {
        const int c[3] = {(int)(cell % a.n[0]), (int)((cell / a.n[0]) % a.n[1]), dim == 3 ? (int)(cell / ((long long)a.n[0] * a.n[1])) : 0};
        long long mat  = cell;
        if (a.dedup)
          {
            int t = 0, mul = 1;
            for (int d = 0; d < dim; ++d)
              {
                const int cls = a.n[d] == 1 ? 0 : (c[d] == 0 ? 0 : (c[d] == a.n[d] - 1 ? 2 : 1));
                t += cls * mul;
                mul *= 3;
              }
            mat = a.type_to_mat[t];
          }
        const __half2 *Bm = reinterpret_cast<const __half2 *>(invH + (size_t)mat * nrow * ld);
        const float    ms = mscale[mat] * (float)scale;
        __syncthreads();
        for (int r = threadIdx.x; r < nrow; r += blockDim.x)
          {
            const int b = r / nc, l = r % nc;
            const int li[3] = {l % n1, (l / n1) % n1, dim == 3 ? l / (n1 * n1) : 0};
            const long long gi = (long long)(c[0] * k + li[0]) + (long long)a.np[0] * ((c[1] * k + li[1]) + (long long)a.np[1] * (c[2] * k + li[2]));
            x[r] = (float)src[(size_t)b * a.N + gi];
          }
        __syncthreads();
        for (int t = threadIdx.x; 2 * t < nrow; t += blockDim.x)
          {
            float          s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            const __half2 *col = Bm + t;
            int            cc  = 0;
#pragma unroll 4
            for (; cc + 1 < nrow; cc += 2)
              {
                const float2 f0 = __half22float2(col[(size_t)cc * ld2]);
                const float2 f1 = __half22float2(col[(size_t)(cc + 1) * ld2]);
                const float  x0 = x[cc], x1 = x[cc + 1];
                s0 += f0.x * x0;
                s1 += f0.y * x0;
                s2 += f1.x * x1;
                s3 += f1.y * x1;
              }
            if (cc < nrow)
              {
                const float2 f0 = __half22float2(col[(size_t)cc * ld2]);
                s0 += f0.x * x[cc];
                s1 += f0.y * x[cc];
              }
            const float v[2] = {(s0 + s2) * ms, (s1 + s3) * ms};
            for (int e = 0; e < 2; ++e)
              {
                const int r = 2 * t + e;
                if (r >= nrow) break;
                const int b = r / nc, l = r % nc;
                const int li[3] = {l % n1, (l / n1) % n1, dim == 3 ? l / (n1 * n1) : 0};
                const long long gi = (long long)(c[0] * k + li[0]) + (long long)a.np[0] * ((c[1] * k + li[1]) + (long long)a.np[1] * (c[2] * k + li[2]));
                atomicAdd(dst + (size_t)b * a.N + gi, (T)v[e]);
              }
          }
      }
  }


  template <typename T>
  struct Vanka
  {
    stfem_op      *op = nullptr;
    T             *d_invT = nullptr;
    __half        *d_invH = nullptr;  // FP16 storage of the inverses (normalised per matrix), [mat][column][ld]
    float         *d_mscale = nullptr;
    int            ld_half = 0;
    long long      n_mat = 0;
    int            nrow = 0;
    VankaApplyArgs args;
    size_t         bytes = 0;
    // fast-diagonalisation form (vanka_fd.cuh): 3D Cartesian constant-coefficient levels
    bool                fd = false;
    T                  *d_modes = nullptr;
    std::vector<double> fdS, fdST; // [3][4][n1*n1]

    ~Vanka()
    {
      if (d_invT) cudaFree(d_invT);
      if (d_invH) cudaFree(d_invH);
      if (d_mscale) cudaFree(d_mscale);
      if (d_modes) cudaFree(d_modes);
    }

    int setup(stfem_op *op_, bool half_storage = false);
    int setup_fd();
    template <int N1, int NB>
    int launch_fd(const T *src, T *dst, T scale);
    int apply_fd(const T *src, T *dst, T scale);
    // dst += scale * sum_c R_c^T B_c^-1 R_c src
    int vmult_add(BlockVec<T> &dst, const BlockVec<T> &src, T scale)
    {
      stfem_ctx *ctx = op->mesh->ctx;
      if (fd) return apply_fd(src.d, dst.d, scale);
      const long long cap = (long long)ctx->sm_count * 16;
      const int grid = (int)(args.n_cells < cap ? args.n_cells : cap);
      if (d_invH)
        {
          const int pairs = (nrow + 1) / 2;
          const int th    = pairs < 32 ? 32 : (pairs > 256 ? 256 : ((pairs + 31) / 32) * 32);
          k_vanka_apply_half<T><<<grid, th, sizeof(float) * nrow, ctx->stream>>>(args, ld_half, d_invH, d_mscale, src.d, dst.d, scale);
          ctx->launches++;
          STFEM_CUDA_CHECK(cudaGetLastError());
          return STFEM_OK;
        }
      const int threads = nrow < 64 ? 64 : (nrow > 256 ? 256 : ((nrow + 31) / 32) * 32);
      k_vanka_apply<T><<<grid, threads, sizeof(T) * nrow, ctx->stream>>>(args, d_invT, src.d, dst.d, scale);
      ctx->launches++;
      STFEM_CUDA_CHECK(cudaGetLastError());
      return STFEM_OK;
    }
    // dst = sum_c R_c^T B_c^-1 R_c src   (stmg.h:832-872; dst zeroed first)
    int vmult(BlockVec<T> &dst, const BlockVec<T> &src)
    {
      STFEM_FORWARD(dst.zero());
      return vmult_add(dst, src, T(1));
    }
  };

  int metric_double(stfem_op *op, double **out); // capi_op.cu: double metric incl. per-q coefficient (caller frees)

  template <typename T>
  int Vanka<T>::setup(stfem_op *op_, bool half_storage)
  {
    op = op_;
    stfem_mesh *m   = op->mesh;
    stfem_ctx  *ctx = m->ctx;
    STFEM_REQUIRE(op->nb_rows == op->nb_cols, "Vanka: operator must be square in time");
    const int dim = m->dim, n1 = op->degree + 1, nb = op->nb_rows;
    const int nc = dim == 3 ? n1 * n1 * n1 : n1 * n1;
    nrow         = nb * nc;
    const bool general = op->d_metric != nullptr || !op->mesh->cartesian; // (operators with on-the-fly geometry keep no metric)
    const bool uniform = !general && op->h_coeff_cell.empty();
    const bool dedup   = uniform && !m->part.active; // (the position classes of a partitioned brick depend on its neighbours: one patch per cell there)
    args.dim = dim; args.n1 = n1; args.nb = nb; args.n_cells = m->n_cells; args.N = op->N;
    for (int d = 0; d < 3; ++d) { args.n[d] = m->n[d]; args.np[d] = op->np[d]; }
    // variant 2 on the level operator keeps the dense patch inverses (cross-check of the Kronecker form)
    if (uniform && dim == 3 && n1 >= 2 && n1 <= 6 && op->variant != 2 && (nb <= 4 || nb == 6 || nb == 8))
      return setup_fd();
    // Geometry the patch matrices are assembled on.  Partitioned meshes: the local brick extended by one ghost cell layer
    // across every face shared with another rank (the neighbours of an interface cell contribute to its patch, and the
    // valence of an interface DoF counts them), described by a temporary mesh / operator pair; the caller provides the
    // ghost vertices (general meshes) and ghost coefficients (stfem_mesh_set_ghost_vertices, stfem_op_set_ghost_coefficients).
    stfem_mesh *gm = m;
    stfem_op   *gop = op;
    std::unique_ptr<stfem_mesh> xm;
    std::unique_ptr<stfem_op>   xo;
    int                         glo[3] = {0, 0, 0};
    if (m->part.active)
      {
        xm            = std::make_unique<stfem_mesh>();
        xm->ctx       = ctx;
        xm->dim       = dim;
        xm->cartesian = m->cartesian;
        xm->dirichlet = m->dirichlet;
        xm->n_cells   = 1;
        size_t nv     = 1;
        for (int d = 0; d < dim; ++d)
          {
            glo[d]          = m->part.neighbor[d][0] >= 0 ? 1 : 0;
            const int    gh = m->part.neighbor[d][1] >= 0 ? 1 : 0;
            const double h  = (m->upper[d] - m->lower[d]) / m->n[d];
            xm->n[d]        = m->n[d] + glo[d] + gh;
            xm->lower[d]    = m->lower[d] - glo[d] * h;
            xm->upper[d]    = m->upper[d] + gh * h;
            xm->n_cells *= xm->n[d];
            nv *= (size_t)xm->n[d] + 1;
          }
        if (!m->cartesian)
          {
            STFEM_REQUIRE(m->h_vertices_ghost.size() == nv * dim,
                          "Vanka (dense patches) on a partitioned general mesh needs the ghost vertices (stfem_mesh_set_ghost_vertices)");
            STFEM_CUDA_CHECK(cudaMalloc(&xm->d_vertices, nv * dim * sizeof(double)));
            STFEM_CUDA_CHECK(cudaMemcpyAsync(xm->d_vertices, m->h_vertices_ghost.data(), nv * dim * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
          }
        xo         = std::make_unique<stfem_op>();
        xo->mesh   = xm.get();
        xo->degree = op->degree;
        xo->shape  = std::make_unique<ShapeHost>(op->degree);
        if (!op->h_coeff_cell.empty())
          {
            STFEM_REQUIRE((long long)op->h_coeff_cell_ghost.size() == xm->n_cells,
                          "Vanka (dense patches) on a partitioned mesh needs the per-cell coefficient of the ghost cells (stfem_op_set_ghost_coefficients)");
            xo->h_coeff_cell = op->h_coeff_cell_ghost;
          }
        if (!op->h_coeff_q.empty())
          {
            const long long nq = dim == 3 ? n1 * n1 * n1 : n1 * n1;
            STFEM_REQUIRE((long long)op->h_coeff_q_ghost.size() == xm->n_cells * nq,
                          "Vanka (dense patches) on a partitioned mesh needs the per-q coefficient of the ghost cells (stfem_op_set_ghost_coefficients)");
            xo->h_coeff_q = op->h_coeff_q_ghost;
          }
        gm  = xm.get();
        gop = xo.get();
      }
    VankaGeom g;
    g.dim = dim; g.n1 = n1; g.n_cells = gm->n_cells; g.dirichlet = gm->dirichlet;
    for (int d = 0; d < 3; ++d)
      {
        g.n[d]  = d < dim ? gm->n[d] : 1;
        g.np[d] = d < dim ? op->degree * gm->n[d] + 1 : 1;
        g.h[d]  = d < dim ? (gm->upper[d] - gm->lower[d]) / gm->n[d] : 1.0;
      }
    for (int q = 0; q < n1 * n1; ++q) { g.S[q] = op->shape->S[q]; g.D[q] = op->shape->D[q]; }
    for (int q = 0; q < n1; ++q) g.w[q] = op->shape->wq[q];
    double    *d_metric = nullptr, *d_coeff = nullptr;
    if (general) STFEM_FORWARD(metric_double(gop, &d_metric));
    g.metric = d_metric;
    if (!gop->h_coeff_cell.empty())
      {
        STFEM_CUDA_CHECK(cudaMalloc(&d_coeff, sizeof(double) * gm->n_cells));
        STFEM_CUDA_CHECK(cudaMemcpyAsync(d_coeff, gop->h_coeff_cell.data(), sizeof(double) * gm->n_cells, cudaMemcpyHostToDevice, ctx->stream));
      }
    g.coeff_cell = d_coeff;
    // patch list
    std::vector<long long> cells;
    args.dedup = dedup ? 1 : 0;
    for (int t = 0; t < 27; ++t) args.type_to_mat[t] = -1;
    if (dedup)
      {
        const int ntz = dim == 3 ? 3 : 1;
        for (int tz = 0; tz < ntz; ++tz)
          for (int ty = 0; ty < 3; ++ty)
            for (int tx = 0; tx < 3; ++tx)
              {
                const int t[3] = {tx, ty, tz};
                bool      ok   = true;
                long long c[3] = {0, 0, 0};
                for (int d = 0; d < dim; ++d)
                  {
                    const int nd = m->n[d];
                    if (t[d] == 0) c[d] = 0;
                    else if (t[d] == 2) { if (nd < 2) ok = false; c[d] = nd - 1; }
                    else { if (nd < 3) ok = false; c[d] = 1; }
                  }
                if (!ok) continue;
                args.type_to_mat[tx + 3 * (ty + 3 * tz)] = (int)cells.size();
                cells.push_back(c[0] + (long long)m->n[0] * (c[1] + (long long)m->n[1] * c[2]));
              }
      }
    else
      {
        // the local cells, numbered on the geometry's (possibly ghost-extended) brick
        cells.resize(m->n_cells);
        for (long long c = 0; c < m->n_cells; ++c)
          {
            const long long cx = c % m->n[0], cy = (c / m->n[0]) % m->n[1], cz = dim == 3 ? c / ((long long)m->n[0] * m->n[1]) : 0;
            cells[c] = (cx + glo[0]) + (long long)g.n[0] * ((cy + glo[1]) + (long long)g.n[1] * (cz + glo[2]));
          }
      }
    n_mat = (long long)cells.size();
    bytes = sizeof(T) * (size_t)n_mat * nrow * nrow;
    STFEM_CUDA_CHECK(cudaMalloc(&d_invT, bytes));
    double *d_alpha = nullptr, *d_beta = nullptr;
    STFEM_CUDA_CHECK(cudaMalloc(&d_alpha, sizeof(double) * nb * nb));
    STFEM_CUDA_CHECK(cudaMalloc(&d_beta, sizeof(double) * nb * nb));
    STFEM_CUDA_CHECK(cudaMemcpyAsync(d_alpha, op->Alpha.data(), sizeof(double) * nb * nb, cudaMemcpyHostToDevice, ctx->stream));
    STFEM_CUDA_CHECK(cudaMemcpyAsync(d_beta, op->Beta.data(), sizeof(double) * nb * nb, cudaMemcpyHostToDevice, ctx->stream));
    // chunks of patches so the double work matrices stay below ~1.5 GB
    const size_t per = sizeof(double) * (size_t)nrow * nrow;
    long long    chunk = (long long)((1500ull << 20) / per);
    if (chunk < 1) chunk = 1;
    if (chunk > n_mat) chunk = n_mat;
    double    *d_B = nullptr;
    long long *d_cells = nullptr;
    int       *d_perm = nullptr, *d_info = nullptr;
    STFEM_CUDA_CHECK(cudaMalloc(&d_B, per * chunk));
    STFEM_CUDA_CHECK(cudaMalloc(&d_cells, sizeof(long long) * chunk));
    STFEM_CUDA_CHECK(cudaMalloc(&d_perm, sizeof(int) * chunk * nrow));
    STFEM_CUDA_CHECK(cudaMalloc(&d_info, sizeof(int)));
    STFEM_CUDA_CHECK(cudaMemsetAsync(d_info, 0, sizeof(int), ctx->stream));
    for (long long c0 = 0; c0 < n_mat; c0 += chunk)
      {
        const long long cn = std::min(chunk, n_mat - c0);
        STFEM_CUDA_CHECK(cudaMemcpyAsync(d_cells, cells.data() + c0, sizeof(long long) * cn, cudaMemcpyHostToDevice, ctx->stream));
        const long long total = cn * nc * nc;
        k_vanka_build<<<grid_for(ctx, total, 128), 128, 0, ctx->stream>>>(g, d_cells, (int)cn, nb, d_alpha, d_beta, d_B);
        k_batched_inverse<<<(unsigned)cn, 256, sizeof(double) * nrow, ctx->stream>>>(d_B, nrow, d_perm, d_info);
        k_store_inverse_T<T><<<grid_for(ctx, cn * nrow * nrow, 256), 256, 0, ctx->stream>>>(d_B, nrow, cn, d_invT + (size_t)c0 * nrow * nrow);
        ctx->launches += 3;
        STFEM_CUDA_CHECK(cudaGetLastError());
      }
    int info = 0;
    STFEM_CUDA_CHECK(cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (void *p : {(void *)d_B, (void *)d_cells, (void *)d_perm, (void *)d_info, (void *)d_alpha, (void *)d_beta, (void *)d_metric, (void *)d_coeff})
      if (p) cudaFree(p);
    if (xm && xm->d_vertices) cudaFree(xm->d_vertices);
    STFEM_REQUIRE(info == 0, "Vanka: singular patch matrix (patch %d)", info - 1);
    args.dim = dim; args.n1 = n1; args.nb = nb; args.n_cells = m->n_cells; args.N = op->N;
    for (int d = 0; d < 3; ++d) { args.n[d] = m->n[d]; args.np[d] = op->np[d]; }
    static const bool env_half = std::getenv("STFEM_VANKA_HALF") != nullptr && std::getenv("STFEM_VANKA_HALF")[0] == '1';
    if (half_storage || env_half)
      {
        ld_half = (nrow + 1) & ~1;
        STFEM_CUDA_CHECK(cudaMalloc(&d_invH, sizeof(__half) * (size_t)n_mat * nrow * ld_half));
        STFEM_CUDA_CHECK(cudaMalloc(&d_mscale, sizeof(float) * (size_t)n_mat));
        k_vanka_to_half<T><<<(unsigned)n_mat, 256, 0, ctx->stream>>>(nrow, ld_half, d_invT, d_invH, d_mscale);
        ctx->launches++;
        STFEM_CUDA_CHECK(cudaGetLastError());
        STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_invT);
        d_invT = nullptr;
        bytes  = sizeof(__half) * (size_t)n_mat * nrow * ld_half + sizeof(float) * (size_t)n_mat;
      }
    return STFEM_OK;
  }

  // ------------------------------------------------------------------ fast-diagonalisation set-up (host, double)
  namespace fdhost
  {
    // symmetric Jacobi eigenvalue iteration: A (n x n, row-major, destroyed) -> eigenvalues on the diagonal, Q columns
    inline void jacobi_eig(std::vector<double> &A, std::vector<double> &Q, int n)
    {
      Q.assign((size_t)n * n, 0.0);
      for (int i = 0; i < n; ++i) Q[i * n + i] = 1.0;
      for (int sweep = 0; sweep < 100; ++sweep)
        {
          double off = 0, diag = 0;
          for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) (i == j ? diag : off) += A[i * n + j] * A[i * n + j];
          if (off <= 1e-32 * (diag + 1e-300)) break;
          for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q)
              {
                const double apq = A[p * n + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                const double t     = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s_ = t * c;
                for (int k = 0; k < n; ++k)
                  {
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s_ * akq;
                    A[k * n + q] = s_ * akp + c * akq;
                  }
                for (int k = 0; k < n; ++k)
                  {
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s_ * aqk;
                    A[q * n + k] = s_ * apk + c * aqk;
                  }
                for (int k = 0; k < n; ++k)
                  {
                    const double qkp = Q[k * n + p], qkq = Q[k * n + q];
                    Q[k * n + p] = c * qkp - s_ * qkq;
                    Q[k * n + q] = s_ * qkp + c * qkq;
                  }
              }
        }
    }

    // generalised symmetric problem K s = lambda M s, M SPD: S (columns) with S^T M S = I, S^T K S = diag(lambda)
    inline void gen_eig(const std::vector<double> &M, const std::vector<double> &K, int n, std::vector<double> &S, std::vector<double> &lam)
    {
      std::vector<double> Lc((size_t)n * n, 0.0);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j)
          {
            double s_ = M[i * n + j];
            for (int k = 0; k < j; ++k) s_ -= Lc[i * n + k] * Lc[j * n + k];
            Lc[i * n + j] = i == j ? std::sqrt(s_) : s_ / Lc[j * n + j];
          }
      // Li = L^-1
      std::vector<double> Li((size_t)n * n, 0.0);
      for (int c = 0; c < n; ++c)
        for (int i = 0; i < n; ++i)
          {
            double s_ = i == c ? 1.0 : 0.0;
            for (int k = 0; k < i; ++k) s_ -= Lc[i * n + k] * Li[k * n + c];
            Li[i * n + c] = s_ / Lc[i * n + i];
          }
      // A = Li K Li^T
      std::vector<double> T1((size_t)n * n, 0.0), A((size_t)n * n, 0.0), Q;
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
          for (int k = 0; k < n; ++k) T1[i * n + j] += Li[i * n + k] * K[k * n + j];
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
          for (int k = 0; k < n; ++k) A[i * n + j] += T1[i * n + k] * Li[j * n + k];
      for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) A[i * n + j] = A[j * n + i] = 0.5 * (A[i * n + j] + A[j * n + i]);
      jacobi_eig(A, Q, n);
      lam.resize(n);
      for (int i = 0; i < n; ++i) lam[i] = A[i * n + i];
      // S = Li^T Q
      S.assign((size_t)n * n, 0.0);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
          for (int k = 0; k < n; ++k) S[i * n + j] += Li[k * n + i] * Q[k * n + j];
    }

    // in-place inverse of a small dense matrix (Gauss-Jordan, partial pivoting); false if singular
    inline bool small_inverse(std::vector<double> &A, int n)
    {
      std::vector<double> I((size_t)n * n, 0.0);
      for (int i = 0; i < n; ++i) I[i * n + i] = 1.0;
      for (int c = 0; c < n; ++c)
        {
          int p = c;
          for (int r = c + 1; r < n; ++r)
            if (std::fabs(A[r * n + c]) > std::fabs(A[p * n + c])) p = r;
          if (A[p * n + c] == 0.0) return false;
          if (p != c)
            for (int k = 0; k < n; ++k)
              {
                std::swap(A[p * n + k], A[c * n + k]);
                std::swap(I[p * n + k], I[c * n + k]);
              }
          const double ip = 1.0 / A[c * n + c];
          for (int k = 0; k < n; ++k) { A[c * n + k] *= ip; I[c * n + k] *= ip; }
          for (int r = 0; r < n; ++r)
            if (r != c)
              {
                const double f = A[r * n + c];
                if (f != 0.0)
                  for (int k = 0; k < n; ++k) { A[r * n + k] -= f * A[c * n + k]; I[r * n + k] -= f * I[c * n + k]; }
              }
        }
      A = I;
      return true;
    }
  } // namespace fdhost

  template <typename T>
  int Vanka<T>::setup_fd()
  {
    stfem_mesh *m   = op->mesh;
    stfem_ctx  *ctx = m->ctx;
    const int   n1 = op->degree + 1, k = n1 - 1, nb = op->nb_rows;
    nrow           = nb * n1 * n1 * n1;
    fd             = true;
    fdS.assign((size_t)3 * 4 * n1 * n1, 0.0);
    fdST.assign((size_t)3 * 4 * n1 * n1, 0.0);
    std::vector<double> lam((size_t)3 * 4 * n1, 0.0);
    // reference-cell 1D matrices of the Gauss rule
    std::vector<double> Mh((size_t)n1 * n1, 0.0), Kh((size_t)n1 * n1, 0.0);
    for (int i = 0; i < n1; ++i)
      for (int j = 0; j < n1; ++j)
        {
          long double mm = 0, kk = 0;
          for (int q = 0; q < n1; ++q)
            {
              mm += (long double)op->shape->wq[q] * op->shape->S[q * n1 + i] * op->shape->S[q * n1 + j];
              kk += (long double)op->shape->wq[q] * op->shape->D[q * n1 + i] * op->shape->D[q * n1 + j];
            }
          Mh[i * n1 + j] = (double)mm;
          Kh[i * n1 + j] = (double)kk;
        }
    for (int d = 0; d < 3; ++d)
      {
        const double h = (m->upper[d] - m->lower[d]) / m->n[d];
        for (int cls = 0; cls < 4; ++cls)
          {
            const bool has_lo = cls == 1 || cls == 2, has_hi = cls == 0 || cls == 1;
            std::vector<double> M1((size_t)n1 * n1), K1((size_t)n1 * n1);
            for (int i = 0; i < n1 * n1; ++i) { M1[i] = h * Mh[i]; K1[i] = Kh[i] / h; }
            // neighbouring cell on the shared end node (assembled matrix restricted to the cell)
            if (has_lo) { M1[0] += h * Mh[k * n1 + k]; K1[0] += Kh[k * n1 + k] / h; }
            if (has_hi) { M1[k * n1 + k] += h * Mh[0]; K1[k * n1 + k] += Kh[0] / h; }
            // constrained end nodes: row/column decoupled, positive diagonal kept
            const bool con[2] = {!has_lo && ((m->dirichlet >> (2 * d)) & 1u), !has_hi && ((m->dirichlet >> (2 * d + 1)) & 1u)};
            for (int e = 0; e < 2; ++e)
              if (con[e])
                {
                  const int a_ = e == 0 ? 0 : k;
                  for (int j = 0; j < n1; ++j)
                    if (j != a_) M1[a_ * n1 + j] = M1[j * n1 + a_] = K1[a_ * n1 + j] = K1[j * n1 + a_] = 0.0;
                }
            std::vector<double> S1, l1;
            fdhost::gen_eig(M1, K1, n1, S1, l1);
            double *pS  = fdS.data() + ((size_t)d * 4 + cls) * n1 * n1;
            double *pST = fdST.data() + ((size_t)d * 4 + cls) * n1 * n1;
            for (int a_ = 0; a_ < n1; ++a_)
              for (int q = 0; q < n1; ++q)
                {
                  pS[a_ * n1 + q]  = S1[a_ * n1 + q];
                  pST[q * n1 + a_] = S1[a_ * n1 + q];
                }
            for (int q = 0; q < n1; ++q) lam[((size_t)d * 4 + cls) * n1 + q] = l1[q];
          }
      }
    // per (type, mode): (Beta + lambda Alpha)^-1
    const size_t   n_modes = (size_t)n1 * n1 * n1;
    std::vector<T> modes((size_t)64 * n_modes * nb * nb);
    std::vector<double> Bm((size_t)nb * nb);
    for (int cz = 0; cz < 4; ++cz)
      for (int cy = 0; cy < 4; ++cy)
        for (int cx = 0; cx < 4; ++cx)
          {
            const int type = cx + 4 * (cy + 4 * cz);
            for (int mz = 0; mz < n1; ++mz)
              for (int my = 0; my < n1; ++my)
                for (int mx = 0; mx < n1; ++mx)
                  {
                    const double l = lam[((size_t)0 * 4 + cx) * n1 + mx] + lam[((size_t)1 * 4 + cy) * n1 + my] + lam[((size_t)2 * 4 + cz) * n1 + mz];
                    for (int i = 0; i < nb * nb; ++i) Bm[i] = op->Beta[i] + l * op->Alpha[i];
                    STFEM_REQUIRE(fdhost::small_inverse(Bm, nb), "Vanka (Kronecker form): singular mode matrix");
                    T *out = modes.data() + (((size_t)type * n1 * n1 + (mz * n1 + my)) * n1 + mx) * nb * nb;
                    for (int i = 0; i < nb * nb; ++i) out[i] = (T)Bm[i];
                  }
          }
    bytes = modes.size() * sizeof(T);
    n_mat = 0;
    STFEM_CUDA_CHECK(cudaMalloc(&d_modes, bytes));
    STFEM_CUDA_CHECK(cudaMemcpyAsync(d_modes, modes.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return STFEM_OK;
  }

  template <typename T>
  template <int N1, int NB>
  int Vanka<T>::launch_fd(const T *src, T *dst, T scale)
  {
    stfem_mesh *m   = op->mesh;
    stfem_ctx  *ctx = m->ctx;
    VankaFdArgs<T, N1> a;
    for (int d = 0; d < 3; ++d)
      for (int cls = 0; cls < 4; ++cls)
        for (int i = 0; i < N1 * N1; ++i)
          {
            a.S[d][cls][i]  = (T)fdS[((size_t)d * 4 + cls) * N1 * N1 + i];
            a.ST[d][cls][i] = (T)fdST[((size_t)d * 4 + cls) * N1 * N1 + i];
          }
    for (int d = 0; d < 3; ++d) { a.n[d] = m->n[d]; a.np[d] = op->np[d]; }
    a.n_cells   = m->n_cells;
    a.nb        = NB;
    a.dirichlet = m->dirichlet;
    a.neighbor_mask = 0;
    for (int d = 0; d < 3; ++d)
      for (int sd = 0; sd < 2; ++sd)
        if (m->part.active && m->part.neighbor[d][sd] >= 0) a.neighbor_mask |= 1u << (2 * d + sd);
    a.src = src; a.dst = dst; a.N = op->N; a.modes = d_modes; a.scale = scale;
    const int tpc = NB * N1;
    int       best = 1;
    double    best_score = -1;
    for (int c = 1; c * tpc <= 256; ++c)
      {
        const int    thr = c * tpc;
        const double eff = (double)thr / (((thr + 31) / 32) * 32);
        const double score = eff >= 0.9 ? 2.0 - 1e-4 * thr : eff; // smallest CTA with well-filled warps (see launch_cart)
        if (score > best_score + 1e-9) { best_score = score; best = c; }
      }
    a.cells_per_cta = best;
    const size_t smem = (size_t)best * NB * ExchLayout<N1>::CBS * sizeof(T);
    auto kern = k_vanka_fd<N1, NB, T>;
    if (smem > 48 * 1024) STFEM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = (m->n_cells + best - 1) / best;
    kern<<<(unsigned)grid, best * tpc, smem, ctx->stream>>>(a);
    ctx->launches++;
    STFEM_CUDA_CHECK(cudaGetLastError());
    return STFEM_OK;
  }

  template <typename T>
  int Vanka<T>::apply_fd(const T *src, T *dst, T scale)
  {
    const int n1 = op->degree + 1, nb = op->nb_rows;
#define STFEM_FD_CASE(N1_, NB_) \
  if (n1 == N1_ && nb == NB_) return launch_fd<N1_, NB_>(src, dst, scale);
#define STFEM_FD_DEG(N1_) \
  STFEM_FD_CASE(N1_, 1) STFEM_FD_CASE(N1_, 2) STFEM_FD_CASE(N1_, 3) STFEM_FD_CASE(N1_, 4) STFEM_FD_CASE(N1_, 6) STFEM_FD_CASE(N1_, 8)
    STFEM_FD_DEG(2) STFEM_FD_DEG(3) STFEM_FD_DEG(4) STFEM_FD_DEG(5) STFEM_FD_DEG(6)
#undef STFEM_FD_DEG
#undef STFEM_FD_CASE
    set_error("Vanka (Kronecker form): degree %d with %d blocks not instantiated", n1 - 1, nb);
    return STFEM_ERR_UNSUPPORTED;
  }
} // namespace stfem
