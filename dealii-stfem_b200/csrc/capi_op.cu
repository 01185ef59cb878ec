// C ABI: space-time operator (SystemMatrix + MatrixFreeOperator of the reference,
// include/operators.h:516-663 and :967-1191).
#include <chrono>
#include <cstdlib>

#include "basis_host.hpp"
#include "common.hpp"
#include "op.hpp"
#include "st_vmult_cart.cuh"
#include "vec.cuh"
#include "st_vmult_generic.cuh"
#include "st_vmult_plane.cuh"
#include "st_vmult_cart_fd.cuh"
#include "assemble.cuh"

namespace stfem
{
  // ------------------------------------------------------------------ metric set-up kernel
  // MappingQ1: J = dx/dxi from the 2^dim cell vertices; stores, per cell and
  // quadrature point, the symmetric tensor J^-1 J^-T * det(J) * w_q and JxW = det(J) * w_q.
  template <int DIM, typename T>
  __global__ void metric_kernel(const double *__restrict__ vertices, int n0, int n1, int n2, int nq1,
                                const double *__restrict__ xq, const double *__restrict__ wq, const double *__restrict__ coeff_q,
                                T *__restrict__ metric)
  {
    constexpr int NSYM = DIM * (DIM + 1) / 2;
    const int     nq   = (DIM == 3) ? nq1 * nq1 * nq1 : nq1 * nq1;
    const long long n_cells = (long long)n0 * n1 * ((DIM == 3) ? n2 : 1);
    const long long gid     = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_cells * nq) return;
    const long long cell = gid / nq;
    const int       q    = (int)(gid % nq);
    const int       qx = q % nq1, qy = (q / nq1) % nq1, qz = (DIM == 3) ? q / (nq1 * nq1) : 0;
    const int       cx = (int)(cell % n0), cy = (int)((cell / n0) % n1), cz = (DIM == 3) ? (int)(cell / ((long long)n0 * n1)) : 0;
    const double    xi[3] = {xq[qx], xq[qy], (DIM == 3) ? xq[qz] : 0.0};
    double          J[DIM][DIM];
    for (int a = 0; a < DIM; ++a)
      for (int b = 0; b < DIM; ++b) J[a][b] = 0;
    for (int v = 0; v < (1 << DIM); ++v)
      {
        const int vx = v & 1, vy = (v >> 1) & 1, vz = (v >> 2) & 1;
        long long vid = (long long)(cx + vx) + (long long)(n0 + 1) * (cy + vy);
        if (DIM == 3) vid += (long long)(n0 + 1) * (n1 + 1) * (cz + vz);
        const double N[3]  = {vx ? xi[0] : 1 - xi[0], vy ? xi[1] : 1 - xi[1], vz ? xi[2] : 1 - xi[2]};
        const double dN[3] = {vx ? 1.0 : -1.0, vy ? 1.0 : -1.0, vz ? 1.0 : -1.0};
        for (int b = 0; b < DIM; ++b)
          {
            double s = dN[b];
            for (int c = 0; c < DIM; ++c)
              if (c != b) s *= N[c];
            for (int a = 0; a < DIM; ++a) J[a][b] += vertices[vid * DIM + a] * s;
          }
      }
    double det, inv[DIM][DIM];
    if (DIM == 2)
      {
        det       = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        inv[0][0] = J[1][1] / det;
        inv[0][1] = -J[0][1] / det;
        inv[1][0] = -J[1][0] / det;
        inv[1][1] = J[0][0] / det;
      }
    else
      {
        const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
        const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        det              = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
        inv[0][0]        = c00 / det;
        inv[1][0]        = c01 / det;
        inv[2][0]        = c02 / det;
        inv[0][1]        = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
        inv[1][1]        = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
        inv[2][1]        = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
        inv[0][2]        = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
        inv[1][2]        = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
        inv[2][2]        = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
      }
    const double w   = wq[qx] * wq[qy] * ((DIM == 3) ? wq[qz] : 1.0);
    const double jxw = det * w;
    const double cq  = coeff_q ? coeff_q[gid] : 1.0;
    T           *out = metric + (size_t)gid * (NSYM + 1);
    int          s   = 0;
    for (int a = 0; a < DIM; ++a)
      for (int b = a; b < DIM; ++b)
        {
          double g = 0;
          for (int c = 0; c < DIM; ++c) g += inv[a][c] * inv[b][c]; // (J^-1 J^-T)_ab, inv[a][c] = dxi_a/dx_c
          out[s++] = (T)(cq * g * jxw);
        }
    out[NSYM] = (T)jxw;
  }

  // uniform vertex grid for Cartesian meshes that need the general metric path (per-q coefficients)
  static int mesh_ensure_vertices(stfem_mesh *m)
  {
    if (m->d_vertices) return STFEM_OK;
    size_t nv = 1;
    for (int d = 0; d < m->dim; ++d) nv *= (size_t)m->n[d] + 1;
    m->h_vertices.resize(nv * m->dim);
    size_t v = 0;
    for (int iz = 0; iz <= (m->dim == 3 ? m->n[2] : 0); ++iz)
      for (int iy = 0; iy <= m->n[1]; ++iy)
        for (int ix = 0; ix <= m->n[0]; ++ix, ++v)
          {
            const int idx[3] = {ix, iy, iz};
            for (int d = 0; d < m->dim; ++d)
              m->h_vertices[v * m->dim + d] = m->lower[d] + (m->upper[d] - m->lower[d]) * idx[d] / m->n[d];
          }
    STFEM_CUDA_CHECK(cudaMalloc(&m->d_vertices, nv * m->dim * sizeof(double)));
    STFEM_CUDA_CHECK(cudaMemcpyAsync(m->d_vertices, m->h_vertices.data(), nv * m->dim * sizeof(double),
                                     cudaMemcpyHostToDevice, m->ctx->stream));
    STFEM_CUDA_CHECK(cudaStreamSynchronize(m->ctx->stream));
    return STFEM_OK;
  }

  // per cell / q-point metric of the operator's mesh in precision T (per-q coefficient folded in)
  template <typename T>
  static int compute_metric(stfem_op *op, void **out)
  {
    stfem_mesh *mesh = op->mesh;
    stfem_ctx  *ctx  = mesh->ctx;
    STFEM_FORWARD(mesh_ensure_vertices(mesh));
    const int       n1   = op->degree + 1;
    const int       nq   = mesh->dim == 3 ? n1 * n1 * n1 : n1 * n1;
    const int       nsym = mesh->dim * (mesh->dim + 1) / 2;
    const long long tot  = mesh->n_cells * nq;
    STFEM_CUDA_CHECK(cudaMalloc(out, (size_t)tot * (nsym + 1) * sizeof(T)));
    double *d_xq = nullptr, *d_wq = nullptr, *d_cq = nullptr;
    if (!op->h_coeff_q.empty())
      {
        STFEM_CUDA_CHECK(cudaMalloc(&d_cq, (size_t)tot * sizeof(double)));
        STFEM_CUDA_CHECK(cudaMemcpyAsync(d_cq, op->h_coeff_q.data(), (size_t)tot * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      }
    STFEM_CUDA_CHECK(cudaMalloc(&d_xq, n1 * sizeof(double)));
    STFEM_CUDA_CHECK(cudaMalloc(&d_wq, n1 * sizeof(double)));
    STFEM_CUDA_CHECK(cudaMemcpyAsync(d_xq, op->shape->xq.data(), n1 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    STFEM_CUDA_CHECK(cudaMemcpyAsync(d_wq, op->shape->wq.data(), n1 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const int       threads = 256;
    const long long blocks  = (tot + threads - 1) / threads;
    if (mesh->dim == 2)
      metric_kernel<2, T><<<(unsigned)blocks, threads, 0, ctx->stream>>>(mesh->d_vertices, mesh->n[0], mesh->n[1], 1, n1, d_xq, d_wq, d_cq, (T *)*out);
    else
      metric_kernel<3, T><<<(unsigned)blocks, threads, 0, ctx->stream>>>(mesh->d_vertices, mesh->n[0], mesh->n[1], mesh->n[2], n1, d_xq, d_wq, d_cq, (T *)*out);
    ctx->launches++;
    STFEM_CUDA_CHECK(cudaGetLastError());
    STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_xq);
    cudaFree(d_wq);
    if (d_cq) cudaFree(d_cq);
    return STFEM_OK;
  }

  int metric_double(stfem_op *op, double **out)
  {
    void *p  = nullptr;
    int   rc = compute_metric<double>(op, &p);
    *out     = (double *)p;
    return rc;
  }

  // ------------------------------------------------------------------ typed launchers
  template <int DIM, int N1, typename T>
  static int launch_generic(stfem_op *op, void *const *dst, const void *const *src, int nb_src, int nb_dst,
                            const void *alpha, const void *beta)
  {
    stfem_mesh     *m = op->mesh;
    VmultArgs<T, N1> a;
    for (int q = 0; q < N1 * N1; ++q)
      {
        a.sh.S[q]  = (T)op->shape->S[q];
        a.sh.Dc[q] = (T)op->shape->Dc[q];
      }
    for (int q = 0; q < N1; ++q) a.sh.w[q] = (T)op->shape->wq[q];
    for (int d = 0; d < 3; ++d)
      {
        a.n[d]  = m->n[d];
        a.np[d] = op->np[d];
        a.h[d]  = (T)((m->upper[d] - m->lower[d]) / m->n[d]);
      }
    a.n_cells = m->n_cells;
    a.nb_src  = nb_src;
    a.nb_dst  = nb_dst;
    for (int b = 0; b < STFEM_MAX_BLOCKS; ++b)
      {
        a.src[b] = b < nb_src ? (const T *)src[b] : nullptr;
        a.dst[b] = b < nb_dst ? (T *)dst[b] : nullptr;
      }
    a.alpha      = (const T *)alpha;
    a.beta       = (const T *)beta;
    STFEM_REQUIRE(m->cartesian || op->d_metric, "st_vmult (generic kernel): this operator keeps no stored metric (geometry on the fly); degree / block count not covered by the plane kernel");
    a.geom_mode  = op->d_metric ? 1 : 0;
    a.metric     = (const T *)op->d_metric;
    a.coeff_cell = (const T *)op->d_coeff;
    a.dirichlet  = m->dirichlet;
    a.nbmax      = nb_src > nb_dst ? nb_src : nb_dst;
    constexpr int NP = (DIM == 3) ? N1 * N1 : N1;
    constexpr int NC = (DIM == 3) ? N1 * N1 * N1 : N1 * N1;
    int           cpc = 192 / (NP * a.nbmax);
    if (cpc < 1) cpc = 1;
    size_t per_slot = (size_t)(DIM + 2) * a.nbmax * NC * sizeof(T);
    while (cpc > 1 && cpc * per_slot > 96 * 1024) --cpc;
    a.cells_per_cta = cpc;
    const size_t smem    = cpc * per_slot + 2 * (size_t)nb_src * nb_dst * sizeof(T);
    const int    threads = cpc * NP * a.nbmax;
    STFEM_REQUIRE(threads <= 1024, "st_vmult: %d threads per CTA exceed 1024 (degree %d, %d blocks)", threads,
                  N1 - 1, a.nbmax);
    STFEM_REQUIRE(smem <= 227 * 1024, "st_vmult: %zu bytes of shared memory exceed 227 KB", smem);
    auto kern = st_vmult_generic_kernel<DIM, N1, T>;
    if (smem > 48 * 1024)
      STFEM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = (m->n_cells + cpc - 1) / cpc;
    kern<<<(unsigned)grid, threads, smem, m->ctx->stream>>>(a);
    m->ctx->launches++;
    STFEM_CUDA_CHECK(cudaGetLastError());
    return STFEM_OK;
  }


  // Cartesian 3D fast path (st_vmult_cart.cuh)
  // PIPE = 0: one batch of cells per CTA; PIPE = NBS > 0: persistent software-pipelined kernel for NBS source blocks
  // (register budget 65536 / (MAXT * MINB), rounded down to a multiple of 8)
  template <int N1, typename T, int MAXT, int MINB, int PIPE = 0, int EXPERIMENT = 0, bool PACKED = false>
  static int launch_cart(stfem_op *op, void *const *dst, const void *const *src, int nb_src, int nb_dst, const void *alpha,
                         const void *beta)
  {
    stfem_mesh      *m = op->mesh;
    CartArgs<T, N1>  a;
    const ShapeHost &sh = *op->shape;
    double           h[3], vol = 1;
    for (int d = 0; d < 3; ++d)
      {
        h[d] = (m->upper[d] - m->lower[d]) / m->n[d];
        vol *= h[d];
        a.n[d]  = m->n[d];
        a.np[d] = op->np[d];
      }
    for (int i = 0; i < N1; ++i)
      for (int j = 0; j < N1; ++j)
        {
          long double mm = 0, kk = 0;
          for (int q = 0; q < N1; ++q)
            {
              mm += (long double)sh.wq[q] * sh.S[q * N1 + i] * sh.S[q * N1 + j];
              kk += (long double)sh.wq[q] * sh.D[q * N1 + i] * sh.D[q * N1 + j];
            }
          a.M[i * N1 + j]  = (T)mm;
          a.Ky[i * N1 + j] = (T)(kk / (h[1] * h[1]));
          a.Kz[i * N1 + j] = (T)(kk / (h[2] * h[2]));
          a.Mx[i * N1 + j] = (T)(mm * vol);
          a.Kx[i * N1 + j] = (T)(kk * vol / (h[0] * h[0]));
        }
    a.n_cells = 1;
    for (int d = 0; d < 3; ++d)
      {
        a.box_lo[d] = op->box_lo ? op->box_lo[d] : 0;
        a.box_n[d]  = op->box_n ? op->box_n[d] : m->n[d];
        a.n_cells *= a.box_n[d];
      }
    a.n_xbox = PIPE > 0 ? 0 : op->n_xbox;
    if (a.n_xbox > 0)
      {
        a.n_cells = 0;
        for (int b = 0; b < a.n_xbox; ++b)
          {
            a.xbox_start[b] = (int)a.n_cells;
            long long nb_ = 1;
            for (int d = 0; d < 3; ++d)
              {
                a.xbox_lo[b][d] = op->xbox_lo[b][d];
                a.xbox_n[b][d]  = op->xbox_n[b][d];
                nb_ *= op->xbox_n[b][d];
              }
            a.n_cells += nb_;
          }
        a.xbox_start[a.n_xbox] = (int)a.n_cells;
      }
    if (a.n_cells <= 0) return STFEM_OK;
    STFEM_REQUIRE(a.n_cells < (1ll << 31), "st_vmult: more than 2^31 cells per GPU are not supported");
    cudaStream_t stream = op->launch_stream ? op->launch_stream : m->ctx->stream;
    a.nb_src  = nb_src;
    a.nb_dst  = nb_dst;
    for (int b = 0; b < STFEM_MAX_BLOCKS; ++b)
      {
        a.src[b] = b < nb_src ? (const T *)src[b] : nullptr;
        a.dst[b] = b < nb_dst ? (T *)dst[b] : nullptr;
      }
    a.alpha      = (const T *)alpha;
    a.beta       = (const T *)beta;
    a.coeff_cell = (const T *)op->d_coeff;
    a.dirichlet  = m->dirichlet;
    // cells per CTA: fill warps (threads a multiple of 32 if possible), <= 256 threads, <= 72 KB of shared memory
    const int    tpc     = nb_dst * N1;
    const size_t per_cell = (size_t)2 * nb_dst * ExchLayout<N1>::CBS * sizeof(T);
    int          best = 1;
    double       best_score = -1;
    const size_t smem_cap = (size_t)(220 * 1024) / MINB;
    for (int c = 1; c * tpc <= MAXT; ++c)
      {
        if (c * per_cell > smem_cap) break;
        const int    thr = c * tpc;
        const double eff = (double)thr / (((thr + 31) / 32) * 32);
        // one-warp CTAs are fastest (barriers become warp-local, more independent CTAs per SM): take the smallest
        // CTA whose warps are >= 90 % full, else the fullest one; variant 17 keeps the old "largest CTA" rule
        const double score = op->variant == 17 ? eff + 1e-4 * thr : (eff >= 0.9 ? 2.0 - 1e-4 * thr : eff);
        if (score > best_score + 1e-9)
          {
            best_score = score;
            best       = c;
          }
      }
    STFEM_REQUIRE(tpc <= MAXT && per_cell <= smem_cap, "st_vmult: %d threads per cell / %zu bytes do not fit the CTA (degree %d, %d blocks)",
                  tpc, per_cell, N1 - 1, nb_dst);
    a.cells_per_cta      = best;
    const size_t smem    = best * per_cell;
    const int    threads = best * tpc;
    long long grid = (a.n_cells + best - 1) / best;
    if constexpr (PIPE > 0)
      {
        STFEM_REQUIRE(nb_src == PIPE, "st_vmult: pipelined kernel instantiated for %d source blocks, got %d", PIPE, nb_src);
        constexpr int budget = 65536 / (((MAXT + 31) / 32) * 32 * MINB);
        constexpr int maxreg = budget > 255 ? 255 : (budget / 8) * 8;
        auto kern = st_vmult_cart_pipe_kernel<N1, T, PIPE, maxreg, (sizeof(T) == 4)>;
        if (smem > 48 * 1024)
          STFEM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const long long resident = (long long)m->ctx->sm_count * MINB;
        if (grid > resident && op->variant != 24 && op->variant != 25) grid = resident; // 24/25: one batch per CTA (no prefetch)
        kern<<<(unsigned)grid, threads, smem, stream>>>(a);
      }
    else
      {
        auto kern = st_vmult_cart_kernel<N1, T, MAXT, MINB, EXPERIMENT, PACKED>;
        if (smem > 48 * 1024)
          STFEM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)grid, threads, smem, stream>>>(a);
      }
    m->ctx->launches++;
    STFEM_CUDA_CHECK(cudaGetLastError());
    return STFEM_OK;
  }

  // general-geometry 3D path (st_vmult_plane.cuh); OTF: geometry on the fly from the cell vertices
  template <int N1, typename T, int MAXT, int MINB, bool OTF = false>
  static int launch_plane(stfem_op *op, void *const *dst, const void *const *src, int nb_src, int nb_dst, const void *alpha,
                          const void *beta)
  {
    stfem_mesh       *m = op->mesh;
    PlaneArgs<T, N1>  a;
    for (int q = 0; q < N1 * N1; ++q)
      {
        a.S[q]  = (T)op->shape->S[q];
        a.Dc[q] = (T)op->shape->Dc[q];
      }
    a.n_cells = 1;
    for (int d = 0; d < 3; ++d)
      {
        a.n[d]      = m->n[d];
        a.np[d]     = op->np[d];
        a.box_lo[d] = op->box_lo ? op->box_lo[d] : 0;
        a.box_n[d]  = op->box_n ? op->box_n[d] : m->n[d];
        a.n_cells *= a.box_n[d];
      }
    if (a.n_cells <= 0) return STFEM_OK;
    STFEM_REQUIRE(a.n_cells < (1ll << 31), "st_vmult: more than 2^31 cells per GPU are not supported");
    cudaStream_t stream = op->launch_stream ? op->launch_stream : m->ctx->stream;
    a.nb_src = nb_src;
    a.nb_dst = nb_dst;
    for (int b = 0; b < STFEM_MAX_BLOCKS; ++b)
      {
        a.src[b] = b < nb_src ? (const T *)src[b] : nullptr;
        a.dst[b] = b < nb_dst ? (T *)dst[b] : nullptr;
      }
    a.alpha      = (const T *)alpha;
    a.beta       = (const T *)beta;
    a.coeff_cell = (const T *)op->d_coeff;
    a.dirichlet  = m->dirichlet;
    a.metric     = (const T *)op->d_metric_plane;
    a.vertices   = m->d_vertices;
    for (int q = 0; q < N1; ++q)
      {
        a.xq[q] = (T)op->shape->xq[q];
        a.wq[q] = (T)op->shape->wq[q];
      }
    if (OTF) STFEM_REQUIRE(m->d_vertices, "st_vmult (general, on-the-fly geometry): the mesh has no vertices");
    const int    tpc      = nb_dst * N1;
    const size_t per_cell = ((size_t)4 * nb_dst * ExchLayout<N1>::CBS + (OTF ? 3 * N1 * N1 * 3 + 36 : 0)) * sizeof(T);
    const size_t smem_cap = (size_t)(220 * 1024) / MINB;
    STFEM_REQUIRE(tpc <= MAXT && per_cell <= smem_cap, "st_vmult (general): %d threads per cell / %zu bytes do not fit the CTA", tpc, per_cell);
    int    best = 1;
    double best_score = -1;
    for (int c = 1; c * tpc <= MAXT; ++c)
      {
        if (c * per_cell > smem_cap) break;
        const int    thr   = c * tpc;
        const double eff   = (double)thr / (((thr + 31) / 32) * 32);
        // unlike the Cartesian kernel this one is faster with large CTAs (variant 17: smallest well-filled CTA)
        const double score = op->variant != 17 ? eff + 1e-4 * thr : (eff >= 0.9 ? 2.0 - 1e-4 * thr : eff);
        if (score > best_score + 1e-9)
          {
            best_score = score;
            best       = c;
          }
      }
    a.cells_per_cta   = best;
    const size_t smem = best * per_cell;
    auto         kern = st_vmult_plane_kernel<N1, T, MAXT, MINB, OTF>;
    if (smem > 48 * 1024) STFEM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = (a.n_cells + best - 1) / best;
    kern<<<(unsigned)grid, best * tpc, smem, stream>>>(a);
    m->ctx->launches++;
    STFEM_CUDA_CHECK(cudaGetLastError());
    return STFEM_OK;
  }

  // TMA variant of the Cartesian kernel (whole mesh rows, n_x divisible by CPC)
  template <int N1, typename T, int CPC, int MAXT, int MINB>
  static int launch_cart_tma(stfem_op *op, void *const *dst, const void *const *src, int nb_src, int nb_dst, const void *alpha,
                             const void *beta)
  {
    stfem_mesh      *m = op->mesh;
    CartArgs<T, N1>  a;
    const ShapeHost &sh = *op->shape;
    double           h[3], vol = 1;
    for (int d = 0; d < 3; ++d)
      {
        h[d] = (m->upper[d] - m->lower[d]) / m->n[d];
        vol *= h[d];
        a.n[d]      = m->n[d];
        a.np[d]     = op->np[d];
        a.box_lo[d] = 0;
        a.box_n[d]  = m->n[d];
      }
    for (int i = 0; i < N1; ++i)
      for (int j = 0; j < N1; ++j)
        {
          long double mm = 0, kk = 0;
          for (int q = 0; q < N1; ++q)
            {
              mm += (long double)sh.wq[q] * sh.S[q * N1 + i] * sh.S[q * N1 + j];
              kk += (long double)sh.wq[q] * sh.D[q * N1 + i] * sh.D[q * N1 + j];
            }
          a.M[i * N1 + j]  = (T)mm;
          a.Ky[i * N1 + j] = (T)(kk / (h[1] * h[1]));
          a.Kz[i * N1 + j] = (T)(kk / (h[2] * h[2]));
          a.Mx[i * N1 + j] = (T)(mm * vol);
          a.Kx[i * N1 + j] = (T)(kk * vol / (h[0] * h[0]));
        }
    a.n_cells = m->n_cells;
    a.n_xbox  = 0;
    a.nb_src  = nb_src;
    a.nb_dst  = nb_dst;
    for (int b = 0; b < STFEM_MAX_BLOCKS; ++b)
      {
        a.src[b] = b < nb_src ? (const T *)src[b] : nullptr;
        a.dst[b] = b < nb_dst ? (T *)dst[b] : nullptr;
      }
    a.alpha         = (const T *)alpha;
    a.beta          = (const T *)beta;
    a.coeff_cell    = (const T *)op->d_coeff;
    a.dirichlet     = m->dirichlet;
    a.cells_per_cta = CPC;
    constexpr int EPV    = 16 / (int)sizeof(T);
    constexpr int ROWLEN = CPC * (N1 - 1) + 1;
    constexpr int COPY   = ((ROWLEN + EPV - 1 + EPV - 1) / EPV) * EPV;
    const int     tpc    = nb_dst * N1;
    const size_t  smem   = 16 + (size_t)N1 * N1 * nb_src * COPY * sizeof(T) + (size_t)2 * CPC * nb_dst * ExchLayout<N1>::CBS * sizeof(T);
    const int     threads = ((CPC * tpc + 31) / 32) * 32;
    STFEM_REQUIRE(threads <= MAXT && m->n[0] % CPC == 0 && smem <= (size_t)(220 * 1024) / MINB, "st_vmult (TMA): configuration does not fit");
    auto kern = st_vmult_cart_tma_kernel<N1, T, CPC, MAXT, MINB>;
    if (smem > 48 * 1024) STFEM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long       grid     = m->n_cells / CPC;
    const long long resident = (long long)m->ctx->sm_count * MINB;
    if (grid > resident) grid = resident;
    cudaStream_t stream = op->launch_stream ? op->launch_stream : m->ctx->stream;
    kern<<<(unsigned)grid, threads, smem, stream>>>(a);
    m->ctx->launches++;
    STFEM_CUDA_CHECK(cudaGetLastError());
    return STFEM_OK;
  }

  // true if op_apply runs the Cartesian kernel (which supports cell sub-boxes / streams)
  static bool uses_cart(const stfem_op *op)
  {
    return op->mesh->dim == 3 && op->variant != 1 && op->mesh->cartesian && !op->d_metric && op->degree <= 5;
  }

  int ctx_ensure_aux(stfem_ctx *ctx)
  {
    if (ctx->aux[0]) return STFEM_OK;
    for (int i = 0; i < 2; ++i) STFEM_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking));
    STFEM_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    STFEM_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    return STFEM_OK;
  }

  // EXPERIMENTAL (kernel_variant 60): fast-diagonalisation form of the Cartesian operator, st_vmult_cart_fd.cuh.
  // Whole-mesh launches with square time matrices only; everything else keeps the default kernel.
  static void cart_fd_reference_matrices(const ShapeHost &sh, int n1, std::vector<double> &V, std::vector<double> &lam)
  {
    std::vector<double> Mh((size_t)n1 * n1), Kh((size_t)n1 * n1);
    for (int i = 0; i < n1; ++i)
      for (int j = 0; j < n1; ++j)
        {
          long double mm = 0, kk = 0;
          for (int q = 0; q < n1; ++q)
            {
              mm += (long double)sh.wq[q] * sh.S[q * n1 + i] * sh.S[q * n1 + j];
              kk += (long double)sh.wq[q] * sh.D[q * n1 + i] * sh.D[q * n1 + j];
            }
          Mh[i * n1 + j] = (double)mm;
          Kh[i * n1 + j] = (double)kk;
        }
    V.assign((size_t)n1 * n1, 0.0);
    lam.assign(n1, 0.0);
    cartfd_host::pencil_modes(Mh.data(), Kh.data(), n1, V.data(), lam.data());
  }

  template <int N1, int NB, typename T>
  static int launch_cart_fd(stfem_op *op, void *const *dst, const void *const *src, const void *alpha, const void *beta)
  {
    stfem_mesh *m = op->mesh;
    if (op->fd_V.empty()) cart_fd_reference_matrices(*op->shape, N1, op->fd_V, op->fd_lam);
    CartFdArgs<T, N1> a;
    double            h[3], vol = 1;
    for (int d = 0; d < 3; ++d)
      {
        h[d] = (m->upper[d] - m->lower[d]) / m->n[d];
        vol *= h[d];
        a.n[d]  = m->n[d];
        a.np[d] = op->np[d];
      }
    for (int q = 0; q < N1; ++q)
      for (int i = 0; i < N1; ++i)
        {
          const double v    = op->fd_V[(size_t)q * N1 + i];
          a.V[q * N1 + i]   = (T)v;
          a.Vt[i * N1 + q]  = (T)v;
          a.Vtx[i * N1 + q] = (T)(v * vol);
        }
    for (int d = 0; d < 3; ++d)
      for (int q = 0; q < N1; ++q) a.lam[d][q] = (T)(op->fd_lam[q] / (h[d] * h[d]));
    a.n_cells = m->n_cells;
    if (a.n_cells <= 0) return STFEM_OK;
    STFEM_REQUIRE(a.n_cells < (1ll << 31), "st_vmult: more than 2^31 cells per GPU are not supported");
    a.dirichlet = m->dirichlet;
    for (int b = 0; b < STFEM_MAX_BLOCKS; ++b)
      {
        a.src[b] = b < NB ? (const T *)src[b] : nullptr;
        a.dst[b] = b < NB ? (T *)dst[b] : nullptr;
      }
    a.alpha      = (const T *)alpha;
    a.beta       = (const T *)beta;
    a.coeff_cell = (const T *)op->d_coeff;
    const int tpc = NB * N1;
    int       best = 1;
    double    best_score = -1;
    for (int c = 1; c * tpc <= 128; ++c)
      {
        const int    thr = c * tpc;
        const double eff = (double)thr / (((thr + 31) / 32) * 32);
        const double score = eff >= 0.9 ? 2.0 - 1e-4 * thr : eff; // smallest CTA with well-filled warps (see launch_cart)
        if (score > best_score + 1e-9) { best_score = score; best = c; }
      }
    a.cells_per_cta = best;
    const size_t smem = (size_t)best * NB * ExchLayout<N1>::CBS * sizeof(T);
    auto         kern = k_st_vmult_cart_fd<N1, NB, T>;
    if (smem > 48 * 1024) STFEM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaStream_t    stream = op->launch_stream ? op->launch_stream : m->ctx->stream;
    const long long grid   = (a.n_cells + best - 1) / best;
    kern<<<(unsigned)grid, best * tpc, smem, stream>>>(a);
    m->ctx->launches++;
    STFEM_CUDA_CHECK(cudaGetLastError());
    return STFEM_OK;
  }

  // STFEM_ERR_UNSUPPORTED = this (degree, blocks) pair is not instantiated: the caller falls back to the default kernel
  template <typename T>
  static int launch_cart_fd_any(stfem_op *op, void *const *dst, const void *const *src, int nb, const void *alpha, const void *beta)
  {
    const int n1 = op->degree + 1;
#define STFEM_CFD_CASE(N1_, NB_) \
  if (n1 == N1_ && nb == NB_) return launch_cart_fd<N1_, NB_, T>(op, dst, src, alpha, beta);
    STFEM_CFD_CASE(3, 2) STFEM_CFD_CASE(3, 3) STFEM_CFD_CASE(4, 2) STFEM_CFD_CASE(4, 3) STFEM_CFD_CASE(5, 2) STFEM_CFD_CASE(5, 3)
#undef STFEM_CFD_CASE
    return STFEM_ERR_UNSUPPORTED;
  }

  template <int DIM, typename T>
  static int dispatch_degree(stfem_op *op, void *const *dst, const void *const *src, int nb_src, int nb_dst,
                             const void *alpha, const void *beta)
  {
    if (DIM == 3 && op->variant != 1 && op->mesh->cartesian && !op->d_metric)
      {
        if (op->variant == 60 && nb_src == nb_dst && !op->box_lo && op->n_xbox == 0)
          {
            const int rc = launch_cart_fd_any<T>(op, dst, src, nb_dst, alpha, beta);
            if (rc != STFEM_ERR_UNSUPPORTED) return rc;
          }
        const int nbd = nb_dst;
        switch (op->degree)
          {
            case 1: return launch_cart<2, T, 256, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
            case 2:
              if (sizeof(T) == 8 && op->variant != 51) return launch_cart<3, T, 256, 2, 0, 0, true>(op, dst, src, nb_src, nb_dst, alpha, beta);
              return launch_cart<3, T, 256, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
            case 3:
              // variants 10.. = tuning configurations (launch bounds) of the same kernel
              if (op->variant == 11 || nbd * 4 > 128) return launch_cart<4, T, 256, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
              if (op->variant == 12) return launch_cart<4, T, 256, 1>(op, dst, src, nb_src, nb_dst, alpha, beta);
              return launch_cart<4, T, 128, 3>(op, dst, src, nb_src, nb_dst, alpha, beta);
            case 4:
              if (op->variant >= 40 && op->variant <= 42 && !op->box_lo && op->n_xbox == 0) // whole-mesh launches only
                {
                  if (nbd * 5 * 12 <= 128 && op->mesh->n[0] % 12 == 0)
                    {
                      if (op->variant == 40) return launch_cart_tma<5, T, 12, 128, 3>(op, dst, src, nb_src, nb_dst, alpha, beta);
                      if (op->variant == 41) return launch_cart_tma<5, T, 12, 128, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
                    }
                  if (nbd * 5 * 8 <= 128 && op->mesh->n[0] % 8 == 0)
                    return launch_cart_tma<5, T, 8, 128, 3>(op, dst, src, nb_src, nb_dst, alpha, beta);
                }
              if (sizeof(T) == 8 && op->variant >= 31 && op->variant <= 34 && nbd * 5 <= 128)
                {
                  if (op->variant == 31) return launch_cart<5, T, 128, 3, 0, 1>(op, dst, src, nb_src, nb_dst, alpha, beta);
                  if (op->variant == 32) return launch_cart<5, T, 128, 3, 0, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
                  if (op->variant == 33) return launch_cart<5, T, 128, 3, 0, 3>(op, dst, src, nb_src, nb_dst, alpha, beta);
                  return launch_cart<5, T, 128, 3, 0, 4>(op, dst, src, nb_src, nb_dst, alpha, beta);
                }
              if (op->variant == 24 && nb_src == 2 && nbd * 5 <= 128) return launch_cart<5, T, 128, 3, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
              if (op->variant == 25 && nb_src == 2 && nbd * 5 <= 128) return launch_cart<5, T, 128, 2, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
              if (op->variant == 20 && nb_src == 2 && nbd * 5 <= 128) return launch_cart<5, T, 128, 3, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
              if (op->variant == 21 && nb_src == 2 && nbd * 5 <= 160) return launch_cart<5, T, 160, 2, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
              if (op->variant == 22 && nb_src == 2 && nbd * 5 <= 128) return launch_cart<5, T, 128, 2, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
              if (op->variant == 23 && nb_src == 2 && nbd * 5 <= 96) return launch_cart<5, T, 96, 3, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
              if (op->variant == 11 || nbd * 5 > 128) return launch_cart<5, T, 256, 1>(op, dst, src, nb_src, nb_dst, alpha, beta);
              if (op->variant == 12) return launch_cart<5, T, 160, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
              if (op->variant == 13) return launch_cart<5, T, 192, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
              if (op->variant == 14) return launch_cart<5, T, 96, 4>(op, dst, src, nb_src, nb_dst, alpha, beta);
              if (op->variant == 27 && nbd * 5 <= 32) return launch_cart<5, T, 32, 13>(op, dst, src, nb_src, nb_dst, alpha, beta);
              if (op->variant == 28 && nbd * 5 <= 32) return launch_cart<5, T, 32, 20>(op, dst, src, nb_src, nb_dst, alpha, beta);
              if (op->variant == 19 && nbd * 5 <= 32) return launch_cart<5, T, 32, 14>(op, dst, src, nb_src, nb_dst, alpha, beta);
              if (op->variant == 26 && nbd * 5 <= 32) return launch_cart<5, T, 32, 16>(op, dst, src, nb_src, nb_dst, alpha, beta);
              if (op->variant == 15 && nbd * 5 <= 64) return launch_cart<5, T, 64, 6>(op, dst, src, nb_src, nb_dst, alpha, beta);
              if (op->variant == 16 && nbd * 5 <= 32) return launch_cart<5, T, 32, 12>(op, dst, src, nb_src, nb_dst, alpha, beta);
              // variant 18: FP32 persistent kernel with the next batch's gather prefetched before the x sweep
              if (sizeof(T) == 4 && op->variant == 18 && nb_src == 2 && nbd * 5 <= 128)
                return launch_cart<5, T, 128, 3, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
              // FP32: 128 registers suffice without spills -> 16 one-warp CTAs per SM
              if (sizeof(T) == 4 && op->variant == 0 && nbd * 5 <= 32) return launch_cart<5, T, 32, 16>(op, dst, src, nb_src, nb_dst, alpha, beta);
              // FP64: P/Q exchanged as 16-byte pairs (conflict-free 128-bit accesses); variant 51 = two 8-byte fields
              if (sizeof(T) == 8 && op->variant != 51) return launch_cart<5, T, 128, 3, 0, 0, true>(op, dst, src, nb_src, nb_dst, alpha, beta);
              return launch_cart<5, T, 128, 3>(op, dst, src, nb_src, nb_dst, alpha, beta);
            case 5: return launch_cart<6, T, 256, 1>(op, dst, src, nb_src, nb_dst, alpha, beta);
            default: break;
          }
      }
    // general geometry, 3D: geometry on the fly from the cell vertices (kernel_variant 6, or chosen at op_create when the
    // stored metric would not fit; see stfem_op::geom_otf) - otherwise the precomputed metric below, which is faster
    if (DIM == 3 && op->variant != 1 && op->geom_otf && nb_dst * (op->degree + 1) <= 128)
      switch (op->degree)
        {
          case 1: return launch_plane<2, T, 128, 2, true>(op, dst, src, nb_src, nb_dst, alpha, beta);
          case 2: return launch_plane<3, T, 128, 2, true>(op, dst, src, nb_src, nb_dst, alpha, beta);
          case 3:
            if (op->variant == 11) return launch_plane<4, T, 128, 3, true>(op, dst, src, nb_src, nb_dst, alpha, beta);
            return launch_plane<4, T, 128, 2, true>(op, dst, src, nb_src, nb_dst, alpha, beta);
          case 4:
            if (op->variant == 11) return launch_plane<5, T, 128, 2, true>(op, dst, src, nb_src, nb_dst, alpha, beta);
            return launch_plane<5, T, 128, 3, true>(op, dst, src, nb_src, nb_dst, alpha, beta);
          default: break;
        }
    if (DIM == 3 && op->variant != 1 && op->d_metric_plane && nb_dst * (op->degree + 1) <= 128)
      switch (op->degree)
        {
          case 1: return launch_plane<2, T, 128, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
          case 2: return launch_plane<3, T, 128, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
          case 3:
            if (op->variant == 11) return launch_plane<4, T, 128, 3>(op, dst, src, nb_src, nb_dst, alpha, beta);
            return launch_plane<4, T, 128, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
          case 4:
            if (op->variant == 11) return launch_plane<5, T, 128, 2>(op, dst, src, nb_src, nb_dst, alpha, beta);
            return launch_plane<5, T, 128, 3>(op, dst, src, nb_src, nb_dst, alpha, beta);
          default: break;
        }
    switch (op->degree)
      {
        case 1: return launch_generic<DIM, 2, T>(op, dst, src, nb_src, nb_dst, alpha, beta);
        case 2: return launch_generic<DIM, 3, T>(op, dst, src, nb_src, nb_dst, alpha, beta);
        case 3: return launch_generic<DIM, 4, T>(op, dst, src, nb_src, nb_dst, alpha, beta);
        case 4: return launch_generic<DIM, 5, T>(op, dst, src, nb_src, nb_dst, alpha, beta);
        case 5: return launch_generic<DIM, 6, T>(op, dst, src, nb_src, nb_dst, alpha, beta);
        case 6: return launch_generic<DIM, 7, T>(op, dst, src, nb_src, nb_dst, alpha, beta);
        default: set_error("st_vmult: degree %d unsupported (1..6)", op->degree); return STFEM_ERR_UNSUPPORTED;
      }
  }

  int op_apply(stfem_op *op, void *const *dst, const void *const *src, int nb_src, int nb_dst, const void *alpha,
               const void *beta, bool zero_dst, const void *const *rhs)
  {
    stfem_ctx *ctx = op->mesh->ctx;
    STFEM_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (op->timing) STFEM_CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
    const size_t bytes = (size_t)op->N * (op->number_type == STFEM_F64 ? 8 : 4);
    bool         direct_partial = false; // partitioned mesh: dst already holds rhs / multiplicity, accumulate into it, then exchange
    if (rhs)
      {
        // dst = rhs + A src: one pass of the brick kernel on unpartitioned meshes, else copy + accumulate
        STFEM_REQUIRE(zero_dst, "op_apply: rhs needs zero_dst");
        const PartitionInfo &prt = op->mesh->part;
        auto halo_dst = [&]() -> int {
          if (!prt.active) return STFEM_OK;
          if (op->number_type == STFEM_F64) return halo_compress_add<double>(ctx, prt, op->halo, dst, nb_dst, op->np, op->mesh->dim, ctx->stream);
          return halo_compress_add<float>(ctx, prt, op->halo, dst, nb_dst, op->np, op->mesh->dim, ctx->stream);
        };
        if (brick_eligible(op, nb_src, nb_dst, alpha, beta))
          {
            // partitioned meshes: every rank adds rhs / multiplicity at its interface nodes (the kernel knows the shared faces),
            // so the one exchange that sums the partial A src also restores rhs: no scratch vector, no extra pass
            STFEM_FORWARD(brick_launch(op, dst, src, nb_dst, alpha, beta, 2, rhs, false));
            STFEM_FORWARD(halo_dst());
            if (op->timing)
              {
                STFEM_CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
                STFEM_CUDA_CHECK(cudaEventSynchronize(ctx->ev1));
                STFEM_CUDA_CHECK(cudaEventElapsedTime(&op->last_ms, ctx->ev0, ctx->ev1));
              }
            return STFEM_OK;
          }
        for (int b = 0; b < nb_dst; ++b) STFEM_CUDA_CHECK(cudaMemcpyAsync(dst[b], rhs[b], bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        zero_dst = false;
        if (prt.active)
          {
            // the same with the per-cell kernels: dst = rhs / multiplicity, the cell loop reduces straight into it, one exchange
            if (op->number_type == STFEM_F64)
              STFEM_FORWARD(halo_scale_interfaces<double>(ctx, prt, op->halo, dst, nb_dst, op->np, op->mesh->dim));
            else
              STFEM_FORWARD(halo_scale_interfaces<float>(ctx, prt, op->halo, dst, nb_dst, op->np, op->mesh->dim));
            direct_partial = true;
          }
      }
    // partitioned mesh + accumulate: the increment goes to a scratch vector first, so that only the increment is
    // summed over the ranks sharing an interface DoF
    void *const *target = dst;
    const bool   via_scratch = op->mesh->part.active && !zero_dst && !direct_partial;
    if (via_scratch)
      {
        while ((int)op->d_part_scratch.size() < nb_dst)
          {
            void *p = nullptr;
            STFEM_CUDA_CHECK(cudaMalloc(&p, bytes + 16));
            op->d_part_scratch.push_back(p);
          }
        target = op->d_part_scratch.data();
      }
    const PartitionInfo &part = op->mesh->part;
    static const bool no_overlap = std::getenv("STFEM_NO_OVERLAP") != nullptr;
    const int *mn = op->mesh->n;
    // the brick kernel writes every DoF exactly once: no zero fill, accumulation (if any) happens in its store.  On
    // partitioned meshes it runs over the whole brick and ONE grouped exchange follows (latency ~ one send/receive; the
    // shell / interior overlap below belongs to the per-cell kernel, whose three exchange rounds it was built to hide)
    static const bool part_no_brick = std::getenv("STFEM_PART_NO_BRICK") != nullptr;
    const bool brick   = brick_eligible(op, nb_src, nb_dst, alpha, beta) && !(part.active && part_no_brick);
    const bool overlap = !brick && part.active && uses_cart(op) && !no_overlap && mn[0] >= 8 && mn[1] >= 8 && mn[2] >= 8;
    if ((zero_dst || via_scratch) && !brick)
      for (int b = 0; b < nb_dst; ++b) STFEM_CUDA_CHECK(cudaMemsetAsync(target[b], 0, bytes, ctx->stream));
    auto dispatch = [&]() -> int {
      if (brick)
        return brick_launch(op, target, src, nb_dst, alpha, beta, (zero_dst || via_scratch) ? 0 : 1, nullptr, false);
      if (op->mesh->dim == 2)
        return op->number_type == STFEM_F64 ? dispatch_degree<2, double>(op, target, src, nb_src, nb_dst, alpha, beta) :
                                              dispatch_degree<2, float>(op, target, src, nb_src, nb_dst, alpha, beta);
      return op->number_type == STFEM_F64 ? dispatch_degree<3, double>(op, target, src, nb_src, nb_dst, alpha, beta) :
                                            dispatch_degree<3, float>(op, target, src, nb_src, nb_dst, alpha, beta);
    };
    auto halo = [&](cudaStream_t stream) -> int {
      if (op->number_type == STFEM_F64)
        return halo_compress_add<double>(ctx, op->mesh->part, op->halo, target, nb_dst, op->np, op->mesh->dim, stream);
      return halo_compress_add<float>(ctx, op->mesh->part, op->halo, target, nb_dst, op->np, op->mesh->dim, stream);
    };
    // multi-GPU: interface DoFs hold partial sums -> add over the ranks sharing them (cell_loop's compress(add)).
    // Large Cartesian bricks: the shell of cells touching a rank interface runs first, the exchange then overlaps
    // the interior cells on a second stream.
    if (overlap)
      {
        STFEM_FORWARD(ctx_ensure_aux(ctx));
        int ilo[3], ihi[3];
        for (int d = 0; d < 3; ++d)
          {
            ilo[d] = part.neighbor[d][0] >= 0 ? 1 : 0;
            ihi[d] = part.neighbor[d][1] >= 0 ? mn[d] - 1 : mn[d];
          }
        // shell: z slabs (full x, y), y slabs (z interior), x slabs (y, z interior) - all in ONE launch
        int rlo[3] = {0, 0, 0}, rhi[3] = {mn[0], mn[1], mn[2]};
        op->n_xbox = 0;
        for (int d = 2; d >= 0; --d)
          {
            for (int sd = 0; sd < 2; ++sd)
              {
                if (part.neighbor[d][sd] < 0) continue;
                int *lo = op->xbox_lo[op->n_xbox], *nn = op->xbox_n[op->n_xbox];
                for (int e = 0; e < 3; ++e)
                  {
                    lo[e] = rlo[e];
                    nn[e] = rhi[e] - rlo[e];
                  }
                lo[d] = sd == 0 ? 0 : mn[d] - 1;
                nn[d] = 1;
                op->n_xbox++;
              }
            rlo[d] = ilo[d];
            rhi[d] = ihi[d];
          }
        if (op->n_xbox > 0)
          {
            const int rc = dispatch();
            op->n_xbox   = 0;
            if (rc != STFEM_OK) return rc;
          }
        STFEM_CUDA_CHECK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        STFEM_CUDA_CHECK(cudaStreamWaitEvent(ctx->aux[0], ctx->ev_fork, 0));
        STFEM_FORWARD(halo(ctx->aux[0]));
        STFEM_CUDA_CHECK(cudaEventRecord(ctx->ev_join, ctx->aux[0]));
        {
          int nn[3] = {ihi[0] - ilo[0], ihi[1] - ilo[1], ihi[2] - ilo[2]};
          op->box_lo = ilo;
          op->box_n  = nn;
          const int rc = dispatch();
          op->box_lo = op->box_n = nullptr;
          if (rc != STFEM_OK) return rc;
        }
        STFEM_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
      }
    else
      {
        STFEM_FORWARD(dispatch());
        if (part.active) STFEM_FORWARD(halo(ctx->stream));
      }
    if (part.active)
      {
        if (via_scratch)
          for (int b = 0; b < nb_dst; ++b)
            {
              if (op->number_type == STFEM_F64)
                k_axpy<double><<<grid_for(ctx, op->N, 256), 256, 0, ctx->stream>>>(op->N, 1.0, (const double *)target[b], (double *)dst[b]);
              else
                k_axpy<float><<<grid_for(ctx, op->N, 256), 256, 0, ctx->stream>>>(op->N, 1.0f, (const float *)target[b], (float *)dst[b]);
              ctx->launches++;
            }
      }
    if (op->timing)
      {
        STFEM_CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
        STFEM_CUDA_CHECK(cudaEventSynchronize(ctx->ev1));
        STFEM_CUDA_CHECK(cudaEventElapsedTime(&op->last_ms, ctx->ev0, ctx->ev1));
      }
    return STFEM_OK;
  }

  template <typename T>
  static int upload_matrix(stfem_ctx *ctx, const std::vector<double> &M, void **d)
  {
    std::vector<T> tmp(M.begin(), M.end());
    STFEM_CUDA_CHECK(cudaMalloc(d, tmp.size() * sizeof(T) + 16));
    STFEM_CUDA_CHECK(cudaMemcpyAsync(*d, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return STFEM_OK;
  }
} // namespace stfem

using namespace stfem;

namespace stfem
{
  // diag K and diag M of the operator's spatial parts (double, N entries each, cudaMalloc'ed: the caller frees),
  // constrained rows 0 (include/operators.h:1092-1110)
  int op_spatial_diagonals(stfem_op *op, double **dK_out, double **dM_out)
  {
    stfem_mesh *m   = op->mesh;
    stfem_ctx  *ctx = m->ctx;
    STFEM_REQUIRE(op->degree >= 1 && op->degree <= 6, "diagonal: degree out of range");
    AsmGeom g;
    fill_asm_geom(g, m, op->degree, op->degree + 1);
    const int nq = m->dim == 3 ? g.nq1 * g.nq1 * g.nq1 : g.nq1 * g.nq1;
    double   *dK = nullptr, *dM = nullptr, *d_cc = nullptr, *d_cq = nullptr;
    STFEM_CUDA_CHECK(cudaMalloc(&dK, sizeof(double) * op->N));
    STFEM_CUDA_CHECK(cudaMalloc(&dM, sizeof(double) * op->N));
    STFEM_CUDA_CHECK(cudaMemsetAsync(dK, 0, sizeof(double) * op->N, ctx->stream));
    STFEM_CUDA_CHECK(cudaMemsetAsync(dM, 0, sizeof(double) * op->N, ctx->stream));
    if (!op->h_coeff_cell.empty())
      {
        STFEM_CUDA_CHECK(cudaMalloc(&d_cc, sizeof(double) * m->n_cells));
        STFEM_CUDA_CHECK(cudaMemcpyAsync(d_cc, op->h_coeff_cell.data(), sizeof(double) * m->n_cells, cudaMemcpyHostToDevice, ctx->stream));
      }
    if (!op->h_coeff_q.empty())
      {
        STFEM_CUDA_CHECK(cudaMalloc(&d_cq, sizeof(double) * m->n_cells * nq));
        STFEM_CUDA_CHECK(cudaMemcpyAsync(d_cq, op->h_coeff_q.data(), sizeof(double) * m->n_cells * nq, cudaMemcpyHostToDevice, ctx->stream));
      }
    const long long cap  = (long long)ctx->sm_count * 8;
    const int       grid = (int)(m->n_cells < cap ? m->n_cells : cap);
    k_diagonal<<<grid, 128, 0, ctx->stream>>>(g, d_cc, d_cq, dK, dM);
    ctx->launches++;
    STFEM_CUDA_CHECK(cudaGetLastError());
    if (m->part.active)
      {
        // partitioned mesh: the cells of the neighbouring ranks contribute to the interface rows (compress(add) of the
        // reference's diagonal vector, operators.h:1092-1110); a one-off exchange with its own buffers
        HaloBuffers hb;
        hb.no_p2p         = true;
        void *blocks[2]   = {dK, dM};
        STFEM_FORWARD(halo_compress_add<double>(ctx, m->part, hb, blocks, 2, op->np, m->dim, ctx->stream));
        STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      }
    STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (d_cc) cudaFree(d_cc);
    if (d_cq) cudaFree(d_cq);
    *dK_out = dK;
    *dM_out = dM;
    return STFEM_OK;
  }
} // namespace stfem

extern "C" {

int stfem_op_create(stfem_mesh_t mesh, const stfem_op_desc *desc, stfem_op_t *out)
{
  STFEM_REQUIRE(mesh && desc && out, "stfem_op_create: null argument");
  STFEM_REQUIRE(desc->degree >= 1 && desc->degree <= 6, "stfem_op_create: degree %d not in 1..6", desc->degree);
  STFEM_REQUIRE(desc->number_type == STFEM_F64 || desc->number_type == STFEM_F32, "stfem_op_create: bad number_type");
  STFEM_REQUIRE(desc->nb_rows >= 1 && desc->nb_rows <= STFEM_MAX_BLOCKS && desc->nb_cols >= 1 &&
                  desc->nb_cols <= STFEM_MAX_BLOCKS,
                "stfem_op_create: block counts %d x %d not in 1..%d", desc->nb_rows, desc->nb_cols, STFEM_MAX_BLOCKS);
  STFEM_REQUIRE(desc->Alpha && desc->Beta, "stfem_op_create: Alpha/Beta null");
  stfem_ctx *ctx = mesh->ctx;
  STFEM_CUDA_CHECK(cudaSetDevice(ctx->device));
  auto op         = std::make_unique<stfem_op>();
  op->mesh        = mesh;
  op->degree      = desc->degree;
  op->number_type = desc->number_type;
  op->nb_rows     = desc->nb_rows;
  op->nb_cols     = desc->nb_cols;
  // 31-34 are timing-only ablation builds of the Cartesian kernel (parts of the algorithm removed: WRONG results).
  // They stay out of reach of normal callers: an explicit opt-in through the environment is required.
  STFEM_REQUIRE(desc->kernel_variant < 31 || desc->kernel_variant > 34 || std::getenv("STFEM_ALLOW_ABLATION") != nullptr,
                "stfem_op_create: kernel_variant %d is a timing-only ablation build (wrong results); set STFEM_ALLOW_ABLATION=1 to use it",
                desc->kernel_variant);
  op->variant     = desc->kernel_variant;
  op->shape       = std::make_unique<ShapeHost>(desc->degree);
  op->N           = 1;
  for (int d = 0; d < 3; ++d)
    {
      op->np[d] = d < mesh->dim ? desc->degree * mesh->n[d] + 1 : 1;
      op->N *= op->np[d];
    }
  const int nr = desc->nb_rows, nc = desc->nb_cols;
  op->Alpha.assign(desc->Alpha, desc->Alpha + nr * nc);
  op->Beta.assign(desc->Beta, desc->Beta + nr * nc);
  std::vector<double> AT(nr * nc), BT(nr * nc);
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j)
      {
        AT[j * nr + i] = op->Alpha[i * nc + j];
        BT[j * nr + i] = op->Beta[i * nc + j];
      }
  std::vector<double> AN(op->Alpha), BN(op->Beta);
  for (auto &v : AN) v = -v;
  for (auto &v : BN) v = -v;
  op->AlphaT = AT;
  op->BetaT = BT;
  op->AlphaNeg = AN;
  op->BetaNeg = BN;
  const bool f64 = desc->number_type == STFEM_F64;
  if (f64)
    {
      STFEM_FORWARD(upload_matrix<double>(ctx, AN, &op->d_alpha_neg));
      STFEM_FORWARD(upload_matrix<double>(ctx, BN, &op->d_beta_neg));
      STFEM_FORWARD(upload_matrix<double>(ctx, op->Alpha, &op->d_alpha));
      STFEM_FORWARD(upload_matrix<double>(ctx, op->Beta, &op->d_beta));
      STFEM_FORWARD(upload_matrix<double>(ctx, AT, &op->d_alphaT));
      STFEM_FORWARD(upload_matrix<double>(ctx, BT, &op->d_betaT));
    }
  else
    {
      STFEM_FORWARD(upload_matrix<float>(ctx, AN, &op->d_alpha_neg));
      STFEM_FORWARD(upload_matrix<float>(ctx, BN, &op->d_beta_neg));
      STFEM_FORWARD(upload_matrix<float>(ctx, op->Alpha, &op->d_alpha));
      STFEM_FORWARD(upload_matrix<float>(ctx, op->Beta, &op->d_beta));
      STFEM_FORWARD(upload_matrix<float>(ctx, AT, &op->d_alphaT));
      STFEM_FORWARD(upload_matrix<float>(ctx, BT, &op->d_betaT));
    }
  if (desc->laplace_coeff_cell)
    {
      op->h_coeff_cell.assign(desc->laplace_coeff_cell, desc->laplace_coeff_cell + mesh->n_cells);
      if (f64)
        STFEM_FORWARD(upload_matrix<double>(ctx, op->h_coeff_cell, &op->d_coeff));
      else
        STFEM_FORWARD(upload_matrix<float>(ctx, op->h_coeff_cell, &op->d_coeff));
    }
  if (!mesh->cartesian || desc->laplace_coeff_q)
    {
      const int n1 = op->degree + 1;
      const int nq = mesh->dim == 3 ? n1 * n1 * n1 : n1 * n1;
      if (desc->laplace_coeff_q) op->h_coeff_q.assign(desc->laplace_coeff_q, desc->laplace_coeff_q + mesh->n_cells * nq);
      // Geometry on the fly or stored?  Stored (8 numbers per quadrature point, streamed by the plane kernel) is faster
      // where it fits (measured on B200, Q4, 96^3 cells: 4.2 ms against 5.9 ms in FP64); it costs (degree+1)^3 * 8 numbers per
      // cell, i.e. 7 GB for that mesh and 95 GB for the 228^3-cell bricks of configs[4].  On the fly (cell vertices only,
      // 192 B per cell) is taken when the stored copy would need more than a quarter of the free device memory, or on
      // request (kernel_variant 6; 5 forces the stored metric).
      if (mesh->dim == 3 && op->degree <= 4 && !mesh->cartesian && !desc->laplace_coeff_q && desc->kernel_variant != 5 && desc->kernel_variant != 1)
        {
          size_t free_b = 0, total_b = 0;
          cudaMemGetInfo(&free_b, &total_b);
          const size_t need = (size_t)mesh->n_cells * nq * 8 * (f64 ? 8 : 4);
          op->geom_otf      = desc->kernel_variant == 6 || need > free_b / 4;
        }
      if (op->geom_otf)
        STFEM_FORWARD(mesh_ensure_vertices(mesh));
      else if (f64)
        STFEM_FORWARD(compute_metric<double>(op.get(), &op->d_metric));
      else
        STFEM_FORWARD(compute_metric<float>(op.get(), &op->d_metric));
      // the re-ordered copy the precomputed-metric plane kernel streams: only where that kernel runs
      if (mesh->dim == 3 && op->degree <= 4 && !op->geom_otf)
        {
          const size_t nvals = (size_t)mesh->n_cells * nq * 8;
          STFEM_CUDA_CHECK(cudaMalloc(&op->d_metric_plane, nvals * (f64 ? 8 : 4) + 16));
          const long long tot = mesh->n_cells * nq;
          if (f64)
            k_metric_to_plane_layout<double><<<grid_for(ctx, tot, 256), 256, 0, ctx->stream>>>((const double *)op->d_metric, mesh->n_cells, n1, (double *)op->d_metric_plane);
          else
            k_metric_to_plane_layout<float><<<grid_for(ctx, tot, 256), 256, 0, ctx->stream>>>((const float *)op->d_metric, mesh->n_cells, n1, (float *)op->d_metric_plane);
          ctx->launches++;
          STFEM_CUDA_CHECK(cudaGetLastError());
          STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        }
    }
  *out = op.release();
  return STFEM_OK;
}

int stfem_op_destroy(stfem_op_t op)
{
  if (!op) return STFEM_OK;
  cudaSetDevice(op->mesh->ctx->device);
  cudaStreamSynchronize(op->mesh->ctx->stream);
  for (void *p : {op->d_alpha, op->d_beta, op->d_alphaT, op->d_betaT, op->d_alpha_neg, op->d_beta_neg, op->d_metric, op->d_metric_plane, op->d_coeff})
    if (p) cudaFree(p);
  for (void *p : op->d_scratch)
    if (p) cudaFree(p);
  for (void *p : op->d_part_scratch)
    if (p) cudaFree(p);
  if (op->d_xface) cudaFree(op->d_xface);
  if (op->h_xface) cudaFreeHost(op->h_xface);
  delete op;
  return STFEM_OK;
}

long long stfem_op_n_dofs_per_block(stfem_op_t op) { return op ? op->N : 0; }
int stfem_op_n_blocks(stfem_op_t op) { return op ? op->nb_rows : 0; }

int stfem_op_vmult(stfem_op_t op, void *const *dst, const void *const *src, int transpose)
{
  STFEM_REQUIRE(op && dst && src, "stfem_op_vmult: null argument");
  STFEM_REQUIRE(op->nb_rows == op->nb_cols, "stfem_op_vmult: operator is %d x %d blocks, not square (use vmult_slice_add)",
                op->nb_rows, op->nb_cols);
  for (int b = 0; b < op->nb_rows; ++b)
    STFEM_REQUIRE(dst[b] && src[b] && dst[b] != src[b], "stfem_op_vmult: block %d null or aliased", b);
  return op_apply(op, dst, src, op->nb_cols, op->nb_rows, transpose ? op->d_alphaT : op->d_alpha,
                  transpose ? op->d_betaT : op->d_beta, true);
}

int stfem_op_vmult_slice_add(stfem_op_t op, void *const *dst, const void *src0)
{
  STFEM_REQUIRE(op && dst && src0, "stfem_op_vmult_slice_add: null argument");
  STFEM_REQUIRE(op->nb_cols == 1, "stfem_op_vmult_slice_add: operator has %d block columns, expected 1", op->nb_cols);
  const void *src[1] = {src0};
  return op_apply(op, dst, src, 1, op->nb_rows, op->d_alpha, op->d_beta, false);
}

int stfem_op_diagonal(stfem_op_t op, void *const *diag)
{
  STFEM_REQUIRE(op && diag, "stfem_op_diagonal: null argument");
  STFEM_REQUIRE(op->nb_rows == op->nb_cols, "stfem_op_diagonal: operator not square in time");
  stfem_ctx *ctx = op->mesh->ctx;
  double    *dK = nullptr, *dM = nullptr;
  STFEM_FORWARD(stfem::op_spatial_diagonals(op, &dK, &dM));
  const int nb = op->nb_rows;
  for (int b = 0; b < nb; ++b)
    {
      const double a = op->Alpha[(size_t)b * nb + b], be = op->Beta[(size_t)b * nb + b];
      if (op->number_type == STFEM_F64)
        k_diag_combine<double><<<grid_for(ctx, op->N, 256), 256, 0, ctx->stream>>>(op->N, a, be, dK, dM, (double *)diag[b]);
      else
        k_diag_combine<float><<<grid_for(ctx, op->N, 256), 256, 0, ctx->stream>>>(op->N, a, be, dK, dM, (float *)diag[b]);
      ctx->launches++;
    }
  STFEM_CUDA_CHECK(cudaGetLastError());
  STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  cudaFree(dK);
  cudaFree(dM);
  return STFEM_OK;
}

} // extern "C"

namespace stfem
{
  // the node column x = ix of every block, packed: out[(b * n_rows + r)] = blocks[b][r * np0 + ix]
  template <typename T>
  __global__ void k_pack_xface(BlockPtrs blocks, int nb, long long n_rows, int np0, int ix, T *__restrict__ out)
  {
    const long long total = n_rows * nb;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
      {
        const int       b = (int)(e / n_rows);
        const long long r = e - (long long)b * n_rows;
        out[e]            = static_cast<const T *>(blocks.p[b])[r * np0 + ix];
      }
  }
} // namespace stfem

extern "C" {

int stfem_op_vmult_host(stfem_op_t op, void *const *dst_host, const void *const *src_host, int transpose)
{
  STFEM_REQUIRE(op && dst_host && src_host, "stfem_op_vmult_host: null argument");
  STFEM_REQUIRE(op->nb_rows == op->nb_cols, "stfem_op_vmult_host: operator not square");
  stfem_ctx   *ctx   = op->mesh->ctx;
  const int    nb    = op->nb_rows;
  const size_t esz   = op->number_type == STFEM_F64 ? 8 : 4;
  const size_t bytes = (size_t)op->N * esz;
  STFEM_CUDA_CHECK(cudaSetDevice(ctx->device));
  if (op->d_scratch.size() < (size_t)2 * nb)
    {
      for (void *p : op->d_scratch) cudaFree(p);
      op->d_scratch.assign(2 * nb, nullptr);
      for (auto &p : op->d_scratch) STFEM_CUDA_CHECK(cudaMalloc(&p, bytes + 16));
    }
  std::vector<void *>       d(nb);
  std::vector<const void *> s(nb);
  for (int b = 0; b < nb; ++b)
    {
      d[b] = op->d_scratch[b];
      s[b] = op->d_scratch[nb + b];
    }
  const void *alpha = transpose ? op->d_alphaT : op->d_alpha, *beta = transpose ? op->d_betaT : op->d_beta;
  const int  *mn = op->mesh->n;
  static const bool no_pipeline = std::getenv("STFEM_NO_PIPELINE") != nullptr;
  // partitioned meshes: pipelined only with the brick kernel (whole x-y cross-sections per slab, no overlap scheme needed)
  bool part_pipeline = false;
  if (op->mesh->part.active && uses_cart(op) && mn[2] >= 4 && !no_pipeline)
    {
      int lo0[3] = {0, 0, 0}, nn0[3] = {mn[0], mn[1], 1};
      op->box_lo = lo0;
      op->box_n  = nn0;
      part_pipeline = brick_eligible(op, nb, nb, alpha, beta);
      op->box_lo = op->box_n = nullptr;
    }
  if (!uses_cart(op) || (op->mesh->part.active && !part_pipeline) || mn[2] < 4 || no_pipeline)
    {
      for (int b = 0; b < nb; ++b)
        STFEM_CUDA_CHECK(cudaMemcpyAsync(op->d_scratch[nb + b], src_host[b], bytes, cudaMemcpyHostToDevice, ctx->stream));
      STFEM_FORWARD(stfem_op_vmult(op, d.data(), s.data(), transpose));
      for (int b = 0; b < nb; ++b)
        STFEM_CUDA_CHECK(cudaMemcpyAsync(dst_host[b], d[b], bytes, cudaMemcpyDeviceToHost, ctx->stream));
      return stream_sync_checked(ctx, "stfem_op_vmult_host");
    }
  // Pipelined over z slabs of cells: upload of slab s+1 (copy engine 1), cell kernel of slab s, download of the
  // node planes slab s has completed (copy engine 2) run concurrently; PCIe is used in both directions at once.
  STFEM_FORWARD(ctx_ensure_aux(ctx));
  const int    k       = op->degree;
  static const int slabs_env = std::getenv("STFEM_HOST_SLABS") ? std::atoi(std::getenv("STFEM_HOST_SLABS")) : 0; // tuning
  const int    slabs_max = slabs_env > 0 ? slabs_env : 16;
  const int    n_slabs = mn[2] < slabs_max ? mn[2] : slabs_max;
  const size_t plane   = (size_t)op->np[0] * op->np[1] * esz; // bytes of one z plane of nodes
  while (ctx->ev_pool.size() < (size_t)2 * n_slabs + 1)
    {
      cudaEvent_t e;
      STFEM_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ctx->ev_pool.push_back(e);
    }
  cudaStream_t up = ctx->aux[0], down = ctx->aux[1];
  // all three streams start after whatever is queued on the context stream
  STFEM_CUDA_CHECK(cudaEventRecord(ctx->ev_fork, ctx->stream));
  STFEM_CUDA_CHECK(cudaStreamWaitEvent(up, ctx->ev_fork, 0));
  STFEM_CUDA_CHECK(cudaStreamWaitEvent(down, ctx->ev_fork, 0));
  // brick kernel: every slab writes its node planes once (the first plane of a slab adds to the partial sum the slab below
  // left there), so dst needs no zero fill
  int  lo0[3] = {0, 0, 0}, nn0[3] = {mn[0], mn[1], 1};
  op->box_lo = lo0;
  op->box_n  = nn0;
  const bool brick = brick_eligible(op, nb, nb, alpha, beta);
  op->box_lo = op->box_n = nullptr;
  if (!brick)
    for (int b = 0; b < nb; ++b) STFEM_CUDA_CHECK(cudaMemsetAsync(d[b], 0, bytes, ctx->stream));
  for (int sl = 0; sl < n_slabs; ++sl)
    {
      const int z0 = (int)((long long)mn[2] * sl / n_slabs), z1 = (int)((long long)mn[2] * (sl + 1) / n_slabs);
      // upload node planes (k z0, k z1] (+ plane 0 for the first slab)
      const size_t p0 = sl == 0 ? 0 : (size_t)k * z0 + 1, p1 = (size_t)k * z1 + 1;
      for (int b = 0; b < nb; ++b)
        STFEM_CUDA_CHECK(cudaMemcpyAsync((char *)op->d_scratch[nb + b] + p0 * plane, (const char *)src_host[b] + p0 * plane, (p1 - p0) * plane,
                                         cudaMemcpyHostToDevice, up));
      STFEM_CUDA_CHECK(cudaEventRecord(ctx->ev_pool[2 * sl], up));
      STFEM_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_pool[2 * sl], 0));
      int lo[3] = {0, 0, z0}, nn[3] = {mn[0], mn[1], z1 - z0};
      op->box_lo = lo;
      op->box_n  = nn;
      int rc = brick ? brick_launch(op, d.data(), s.data(), nb, alpha, beta, 0, nullptr, sl > 0) :
               op->number_type == STFEM_F64 ? dispatch_degree<3, double>(op, d.data(), s.data(), nb, nb, alpha, beta) :
                                              dispatch_degree<3, float>(op, d.data(), s.data(), nb, nb, alpha, beta);
      op->box_lo = op->box_n = nullptr;
      if (rc != STFEM_OK) return rc;
      STFEM_CUDA_CHECK(cudaEventRecord(ctx->ev_pool[2 * sl + 1], ctx->stream));
      STFEM_CUDA_CHECK(cudaStreamWaitEvent(down, ctx->ev_pool[2 * sl + 1], 0));
      // node planes [k z0, k z1) are complete now (the last slab also completes the top plane)
      const size_t q0 = (size_t)k * z0, q1 = sl == n_slabs - 1 ? (size_t)k * z1 + 1 : (size_t)k * z1;
      for (int b = 0; b < nb; ++b)
        STFEM_CUDA_CHECK(cudaMemcpyAsync((char *)dst_host[b] + q0 * plane, (const char *)d[b] + q0 * plane, (q1 - q0) * plane, cudaMemcpyDeviceToHost, down));
    }
  if (part_pipeline)
    {
      // Partitioned mesh: the planes downloaded above hold this rank's partial sums at the interface nodes.  Sum them over
      // the ranks (one exchange, after the last slab), then download the interface nodes once more - behind the bulk copies
      // on the same stream, so the final values land last: z faces as whole planes, y faces as strided rows, x faces
      // packed by a kernel into a pinned staging buffer and scattered by the host after the synchronisation.
      const stfem::PartitionInfo &prt = op->mesh->part;
      if (op->number_type == STFEM_F64)
        STFEM_FORWARD(halo_compress_add<double>(ctx, prt, op->halo, d.data(), nb, op->np, 3, ctx->stream));
      else
        STFEM_FORWARD(halo_compress_add<float>(ctx, prt, op->halo, d.data(), nb, op->np, 3, ctx->stream));
      const long long n_rows = (long long)op->np[1] * op->np[2];
      const int       n_xf   = (prt.neighbor[0][0] >= 0 ? 1 : 0) + (prt.neighbor[0][1] >= 0 ? 1 : 0);
      const size_t    xbytes = (size_t)n_rows * nb * esz;
      if (n_xf > 0)
        {
          if (op->xface_bytes < 2 * xbytes)
            {
              if (op->d_xface) cudaFree(op->d_xface);
              if (op->h_xface) cudaFreeHost(op->h_xface);
              STFEM_CUDA_CHECK(cudaMalloc(&op->d_xface, 2 * xbytes));
              STFEM_CUDA_CHECK(cudaHostAlloc(&op->h_xface, 2 * xbytes, cudaHostAllocDefault));
              op->xface_bytes = 2 * xbytes;
            }
          BlockPtrs bp;
          for (int b = 0; b < STFEM_MAX_BLOCKS; ++b) bp.p[b] = b < nb ? d[b] : nullptr;
          int slot = 0;
          for (int sd = 0; sd < 2; ++sd)
            if (prt.neighbor[0][sd] >= 0)
              {
                const int ix = sd == 0 ? 0 : op->np[0] - 1;
                const int grid = (int)std::min<long long>((n_rows * nb + 255) / 256, (long long)ctx->sm_count * 4);
                if (esz == 8)
                  k_pack_xface<double><<<grid, 256, 0, ctx->stream>>>(bp, nb, n_rows, op->np[0], ix, (double *)((char *)op->d_xface + slot * xbytes));
                else
                  k_pack_xface<float><<<grid, 256, 0, ctx->stream>>>(bp, nb, n_rows, op->np[0], ix, (float *)((char *)op->d_xface + slot * xbytes));
                ctx->launches++;
                ++slot;
              }
          STFEM_CUDA_CHECK(cudaGetLastError());
        }
      STFEM_CUDA_CHECK(cudaEventRecord(ctx->ev_pool[2 * n_slabs], ctx->stream));
      STFEM_CUDA_CHECK(cudaStreamWaitEvent(down, ctx->ev_pool[2 * n_slabs], 0));
      const size_t row = (size_t)op->np[0] * esz;
      for (int b = 0; b < nb; ++b)
        {
          for (int sd = 0; sd < 2; ++sd)
            {
              if (prt.neighbor[2][sd] >= 0)
                {
                  const size_t o = (sd == 0 ? 0 : (size_t)op->np[2] - 1) * plane;
                  STFEM_CUDA_CHECK(cudaMemcpyAsync((char *)dst_host[b] + o, (const char *)d[b] + o, plane, cudaMemcpyDeviceToHost, down));
                }
              if (prt.neighbor[1][sd] >= 0)
                {
                  const size_t o = (sd == 0 ? 0 : (size_t)op->np[1] - 1) * row;
                  STFEM_CUDA_CHECK(cudaMemcpy2DAsync((char *)dst_host[b] + o, plane, (const char *)d[b] + o, plane, row, (size_t)op->np[2], cudaMemcpyDeviceToHost, down));
                }
            }
        }
      if (n_xf > 0) STFEM_CUDA_CHECK(cudaMemcpyAsync(op->h_xface, op->d_xface, n_xf * xbytes, cudaMemcpyDeviceToHost, down));
      STFEM_CUDA_CHECK(cudaEventRecord(ctx->ev_join, down));
      STFEM_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
      STFEM_FORWARD(stream_sync_checked(ctx, "stfem_op_vmult_host"));
      int slot = 0;
      for (int sd = 0; sd < 2; ++sd)
        if (prt.neighbor[0][sd] >= 0)
          {
            const size_t ix = sd == 0 ? 0 : (size_t)op->np[0] - 1;
            for (int b = 0; b < nb; ++b)
              {
                const char *src = (const char *)op->h_xface + slot * xbytes + (size_t)b * n_rows * esz;
                char       *dh  = (char *)dst_host[b] + ix * esz;
                if (esz == 8)
                  for (long long r = 0; r < n_rows; ++r) *(double *)(dh + (size_t)r * row) = ((const double *)src)[r];
                else
                  for (long long r = 0; r < n_rows; ++r) *(float *)(dh + (size_t)r * row) = ((const float *)src)[r];
              }
            ++slot;
          }
      return STFEM_OK;
    }
  STFEM_CUDA_CHECK(cudaEventRecord(ctx->ev_join, down));
  STFEM_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  return stream_sync_checked(ctx, "stfem_op_vmult_host");
}

/* Measurement helper: the copy floor of stfem_op_vmult_host - the operator's nb blocks uploaded and downloaded concurrently
 * (two copy streams, no kernel), host wall time of `reps` rounds in milliseconds per round. */
int stfem_op_host_copy_floor(stfem_op_t op, void *const *dst_host, const void *const *src_host, int reps, double *ms_per_round)
{
  STFEM_REQUIRE(op && dst_host && src_host && ms_per_round && reps >= 1, "stfem_op_host_copy_floor: bad arguments");
  stfem_ctx   *ctx   = op->mesh->ctx;
  const int    nb    = op->nb_rows;
  const size_t bytes = (size_t)op->N * (op->number_type == STFEM_F64 ? 8 : 4);
  STFEM_CUDA_CHECK(cudaSetDevice(ctx->device));
  STFEM_FORWARD(ctx_ensure_aux(ctx));
  if (op->d_scratch.size() < (size_t)2 * nb)
    {
      for (void *p : op->d_scratch) cudaFree(p);
      op->d_scratch.assign(2 * nb, nullptr);
      for (auto &p : op->d_scratch) STFEM_CUDA_CHECK(cudaMalloc(&p, bytes + 16));
    }
  STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  const auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < reps; ++r)
    {
      for (int b = 0; b < nb; ++b)
        {
          STFEM_CUDA_CHECK(cudaMemcpyAsync(op->d_scratch[nb + b], src_host[b], bytes, cudaMemcpyHostToDevice, ctx->aux[0]));
          STFEM_CUDA_CHECK(cudaMemcpyAsync(dst_host[b], op->d_scratch[b], bytes, cudaMemcpyDeviceToHost, ctx->aux[1]));
        }
      STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->aux[0]));
      STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->aux[1]));
    }
  *ms_per_round = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
  return STFEM_OK;
}

int stfem_op_set_timing(stfem_op_t op, int enable)
{
  STFEM_REQUIRE(op, "null op");
  op->timing = enable != 0;
  return STFEM_OK;
}

float stfem_op_last_kernel_ms(stfem_op_t op) { return op ? op->last_ms : -1.f; }

/* kernel_variant 60 (experimental fast-diagonalisation form of the Cartesian operator): the modes of the reference-cell
 * pencil (Kh, Mh) of FE_Q(degree) with QGauss(degree+1).  Host only.  V: (degree+1)^2 row-major, lam: degree+1. */
int stfem_cart_fd_modes(int degree, double *V, double *lam)
{
  STFEM_REQUIRE(degree >= 1 && degree <= 6 && V && lam, "stfem_cart_fd_modes: bad arguments");
  const int           n1 = degree + 1;
  ShapeHost           sh(degree);
  std::vector<double> v, l;
  cart_fd_reference_matrices(sh, n1, v, l);
  std::copy(v.begin(), v.end(), V);
  std::copy(l.begin(), l.end(), lam);
  return STFEM_OK;
}

} // extern "C"
