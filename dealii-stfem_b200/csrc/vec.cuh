// Device block vectors and the BLAS-1 style kernels of the Krylov / smoother updates.
// Replaces deal.II LinearAlgebra::distributed::BlockVector operations used by the reference's hot path
// (add/sadd/equ, dot, l2_norm; tensorproduct_add of include/operators.h:211-283; the double<->float
// copies of GMG::vmult, include/stmg.h:1340-1342).  A block vector is ONE contiguous device array of
// nb * N numbers (block b at offset b*N), so vector updates are single launches over nb*N entries and
// the per-block pointers the operator ABI expects are just offsets.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <vector>

#include "common.hpp"
#include "dist.cuh"

namespace stfem
{
  template <typename T>
  struct BlockVec
  {
    stfem_ctx *ctx = nullptr;
    int        nb  = 0;
    long long  n   = 0; // entries per block
    T         *d   = nullptr;
    bool       owner = false;
    std::vector<void *> ptrs; // per-block pointers (host array of device pointers)

    BlockVec() = default;
    BlockVec(const BlockVec &) = delete;
    BlockVec &operator=(const BlockVec &) = delete;
    BlockVec(BlockVec &&o) noexcept { *this = std::move(o); }
    BlockVec &operator=(BlockVec &&o) noexcept
    {
      release();
      ctx = o.ctx; nb = o.nb; n = o.n; d = o.d; owner = o.owner; ptrs = std::move(o.ptrs);
      o.d = nullptr; o.owner = false;
      return *this;
    }
    ~BlockVec() { release(); }
    void release()
    {
      if (owner && d) cudaFree(d);
      d = nullptr; owner = false;
    }
    int alloc(stfem_ctx *c, int nb_, long long n_)
    {
      release();
      ctx = c; nb = nb_; n = n_;
      STFEM_CUDA_CHECK(cudaMalloc(&d, sizeof(T) * (size_t)nb * n + 16));
      owner = true;
      ptrs.resize(nb);
      for (int b = 0; b < nb; ++b) ptrs[b] = d + (size_t)b * n;
      return zero();
    }
    int zero() { STFEM_CUDA_CHECK(cudaMemsetAsync(d, 0, sizeof(T) * (size_t)nb * n, ctx->stream)); return STFEM_OK; }
    long long size() const { return (long long)nb * n; }
    void *const *block_ptrs() { return ptrs.data(); }
    const void *const *cblock_ptrs() const { return (const void *const *)ptrs.data(); }
  };

  // ------------------------------------------------------------------ elementwise kernels
  template <typename T>
  __global__ void k_axpy(long long n, T a, const T *__restrict__ x, T *__restrict__ y)
  {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      y[i] += a * x[i];
  }
  // y = s*y + a*x
  template <typename T>
  __global__ void k_sadd(long long n, T s, T a, const T *__restrict__ x, T *__restrict__ y)
  {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      y[i] = s * y[i] + a * x[i];
  }
  // z = a*x + b*y
  template <typename T>
  __global__ void k_lincomb(long long n, T a, const T *__restrict__ x, T b, const T *__restrict__ y, T *__restrict__ z)
  {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      z[i] = a * x[i] + b * y[i];
  }
  template <typename T>
  __global__ void k_scale(long long n, T a, T *__restrict__ y)
  {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      y[i] *= a;
  }
  template <typename TO, typename TI>
  __global__ void k_convert(long long n, const TI *__restrict__ x, TO *__restrict__ y)
  {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      y[i] = (TO)x[i];
  }
  // w += sum_k c[k] * V_k   (V_k = base + k*stride), coefficients by value
  constexpr int MAXK = 16;
  template <typename T>
  struct MultiCoef { T c[MAXK]; const T *v[MAXK]; int m; };
  template <typename T>
  __global__ void k_multi_axpy(long long n, MultiCoef<T> mc, T *__restrict__ w)
  {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      {
        T s = w[i];
#pragma unroll 4
        for (int k = 0; k < mc.m; ++k) s += mc.c[k] * mc.v[k][i];
        w[i] = s;
      }
  }
  // w += sum_k c[k] V_k  and  *nrm2 += ||w_new||^2 in the same pass (Gram-Schmidt: the norm of the orthogonalised
  // vector comes for free); mask: see DotMask
  template <typename T>
  __global__ void k_multi_axpy_norm(long long n, MultiCoef<T> mc, T *__restrict__ w, double *__restrict__ nrm2, int mask_active, int np0, int np1,
                                    int np2, unsigned skip_high, long long n_block)
  {
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      {
        T s = w[i];
#pragma unroll 4
        for (int k = 0; k < mc.m; ++k) s += mc.c[k] * mc.v[k][i];
        w[i] = s;
        bool skip = false;
        if (mask_active)
          {
            long long r  = i % n_block;
            const int ix = (int)(r % np0);
            r /= np0;
            const int iy = (int)(r % np1), iz = (int)(r / np1);
            skip = ((skip_high & 1u) && ix == np0 - 1) || ((skip_high & 2u) && iy == np1 - 1) || ((skip_high & 4u) && iz == np2 - 1);
          }
        if (!skip) acc += (double)s * (double)s;
      }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double sh[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0)
      {
        double v = 0;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v += sh[wv];
        atomicAdd(nrm2, v);
      }
  }
  // out = scale * (w + sum_k c[k] V_k): the Gram-Schmidt update and the normalisation of the new basis vector in one pass
  // (w is read, not written: the next Arnoldi step overwrites it anyway)
  template <typename T>
  __global__ void k_multi_axpy_scale_out(long long n, MultiCoef<T> mc, T scale, const T *__restrict__ w, T *__restrict__ out)
  {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      {
        T s = w[i];
#pragma unroll 4
        for (int k = 0; k < mc.m; ++k) s += mc.c[k] * mc.v[k][i];
        out[i] = scale * s;
      }
  }
  // y = a * x
  template <typename T>
  __global__ void k_scale_copy(long long n, T a, const T *__restrict__ x, T *__restrict__ y)
  {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      y[i] = a * x[i];
  }
  // multi-GPU: interface DoFs are duplicated on both ranks; a dot product counts them on the lower rank only
  struct DotMask
  {
    int       active = 0;
    int       np[3] = {1, 1, 1};
    unsigned  skip_high = 0; // bit d: the plane index np[d]-1 belongs to the neighbour rank
    long long n_block = 0;
  };
  __device__ __forceinline__ bool dot_masked(const DotMask &m, long long i)
  {
    long long r  = i % m.n_block;
    const int ix = (int)(r % m.np[0]);
    r /= m.np[0];
    const int iy = (int)(r % m.np[1]);
    const int iz = (int)(r / m.np[1]);
    return ((m.skip_high & 1u) && ix == m.np[0] - 1) || ((m.skip_high & 2u) && iy == m.np[1] - 1) || ((m.skip_high & 4u) && iz == m.np[2] - 1);
  }

  // w += sum_k c[k] V_k  and, with the UPDATED w,  out[k] += <w, V_k> (k < m),  out[m] += ||w||^2  in the same pass: the
  // first Gram-Schmidt update fused with the inner products of the re-orthogonalisation pass
  template <typename T>
  __global__ void k_multi_axpy_dot(long long n, MultiCoef<T> mc, T *__restrict__ w, double *__restrict__ out, DotMask mask)
  {
    double acc[MAXK + 1];
#pragma unroll
    for (int k = 0; k <= MAXK; ++k) acc[k] = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      {
        T vk[MAXK];
        T s = w[i];
#pragma unroll
        for (int k = 0; k < MAXK; ++k)
          if (k < mc.m)
            {
              vk[k] = mc.v[k][i];
              s += mc.c[k] * vk[k];
            }
        w[i] = s;
        if (mask.active && dot_masked(mask, i)) continue;
        const double sd = (double)s;
#pragma unroll
        for (int k = 0; k < MAXK; ++k)
          if (k < mc.m) acc[k] += sd * (double)vk[k];
        acc[MAXK] += sd * sd;
      }
    __shared__ double sh[MAXK + 1][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k <= MAXK; ++k)
      if (k < mc.m || k == MAXK)
        {
          double v = acc[k];
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0) sh[k][warp] = v;
        }
    __syncthreads();
    if (threadIdx.x <= mc.m)
      {
        const int k = threadIdx.x < mc.m ? threadIdx.x : MAXK;
        double    v = 0;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v += sh[k][wv];
        atomicAdd(out + threadIdx.x, v);
      }
  }
  // out[k] += <w, V_k>  accumulated in double; warp shuffle + one atomic per warp
  template <typename T>
  __global__ void k_multi_dot(long long n, MultiCoef<T> mc, const T *__restrict__ w, double *__restrict__ out, DotMask mask)
  {
    double acc[MAXK];
#pragma unroll
    for (int k = 0; k < MAXK; ++k) acc[k] = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      {
        if (mask.active && dot_masked(mask, i)) continue;
        const double wi = (double)w[i];
#pragma unroll
        for (int k = 0; k < MAXK; ++k)
          if (k < mc.m) acc[k] += wi * (double)mc.v[k][i];
      }
    __shared__ double sh[MAXK][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
      if (k < mc.m)
        {
          double v = acc[k];
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0) sh[k][warp] = v;
        }
    __syncthreads();
    if (threadIdx.x < mc.m)
      {
        double v = 0;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v += sh[threadIdx.x][wv];
        atomicAdd(out + threadIdx.x, v);
      }
  }
  // dst_i (+)= sum_j P(i,j) src_j over blocks (tensorproduct / tensorproduct_add, operators.h:252-283)
  template <typename T>
  struct SmallMat { T a[STFEM_MAX_BLOCKS * STFEM_MAX_BLOCKS]; int m, n; };
  template <typename T>
  __global__ void k_block_matmul(long long n, SmallMat<T> P, const T *__restrict__ src, T *__restrict__ dst, int add)
  {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      {
        T s[STFEM_MAX_BLOCKS];
        for (int j = 0; j < P.n; ++j) s[j] = src[(size_t)j * n + i];
        for (int r = 0; r < P.m; ++r)
          {
            T acc = add ? dst[(size_t)r * n + i] : T(0);
            for (int j = 0; j < P.n; ++j) acc += P.a[r * P.n + j] * s[j];
            dst[(size_t)r * n + i] = acc;
          }
      }
  }

  // ---- Gram-Schmidt kernels with the number of basis vectors M as a compile-time constant and TWO entries per thread and
  // sweep (16-byte loads): the generic kernels above keep (M + 1) 8-byte loads per thread in flight, i.e. about 16 KB per SM
  // for M = 3 - less than half of what the HBM latency-bandwidth product asks for (measured: k_multi_axpy_dot 2.4 TB/s,
  // k_multi_dot 4.3 TB/s).  Same arithmetic, same summation structure (per-thread partial sums in double, warp shuffle,
  // one atomic per CTA and vector).
  template <typename T> struct Pair;
  template <> struct Pair<double> { using type = double2; };
  template <> struct Pair<float> { using type = float2; };

  template <int NACC>
  __device__ __forceinline__ void block_reduce_atomic(double (&acc)[NACC], double *__restrict__ out)
  {
    __shared__ double sh[NACC][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NACC; ++k)
      {
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh[k][warp] = v;
      }
    __syncthreads();
    if (threadIdx.x < NACC)
      {
        double v = 0;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v += sh[threadIdx.x][wv];
        atomicAdd(out + threadIdx.x, v);
      }
  }

  // out[k] += <w, V_k>, k < M
  template <typename T, int M>
  __global__ void __launch_bounds__(256) k_multi_dot_m(long long n, MultiCoef<T> mc, const T *__restrict__ w, double *__restrict__ out, DotMask mask)
  {
    using T2 = typename Pair<T>::type;
    double acc[M];
#pragma unroll
    for (int k = 0; k < M; ++k) acc[k] = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x * 2;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += stride)
      {
        if (i + 1 < n)
          {
            const T2 w2 = *reinterpret_cast<const T2 *>(w + i);
            T2       v2[M];
#pragma unroll
            for (int k = 0; k < M; ++k) v2[k] = *reinterpret_cast<const T2 *>(mc.v[k] + i);
            const bool   m0 = mask.active && dot_masked(mask, i), m1 = mask.active && dot_masked(mask, i + 1);
            const double a0 = m0 ? 0.0 : (double)w2.x, a1 = m1 ? 0.0 : (double)w2.y;
#pragma unroll
            for (int k = 0; k < M; ++k) acc[k] += a0 * (double)v2[k].x + a1 * (double)v2[k].y;
          }
        else if (!(mask.active && dot_masked(mask, i)))
          {
            const double a0 = (double)w[i];
#pragma unroll
            for (int k = 0; k < M; ++k) acc[k] += a0 * (double)mc.v[k][i];
          }
      }
    block_reduce_atomic<M>(acc, out);
  }

  // w += sum_k c[k] V_k, then with the updated w: out[k] += <w, V_k> (k < M), out[M] += ||w||^2
  template <typename T, int M>
  __global__ void __launch_bounds__(256) k_multi_axpy_dot_m(long long n, MultiCoef<T> mc, T *__restrict__ w, double *__restrict__ out, DotMask mask)
  {
    using T2 = typename Pair<T>::type;
    double acc[M + 1];
#pragma unroll
    for (int k = 0; k <= M; ++k) acc[k] = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x * 2;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += stride)
      {
        if (i + 1 < n)
          {
            T2 w2 = *reinterpret_cast<const T2 *>(w + i);
            T2 v2[M];
#pragma unroll
            for (int k = 0; k < M; ++k) v2[k] = *reinterpret_cast<const T2 *>(mc.v[k] + i);
#pragma unroll
            for (int k = 0; k < M; ++k)
              {
                w2.x += mc.c[k] * v2[k].x;
                w2.y += mc.c[k] * v2[k].y;
              }
            *reinterpret_cast<T2 *>(w + i) = w2;
            const bool   m0 = mask.active && dot_masked(mask, i), m1 = mask.active && dot_masked(mask, i + 1);
            const double a0 = m0 ? 0.0 : (double)w2.x, a1 = m1 ? 0.0 : (double)w2.y;
#pragma unroll
            for (int k = 0; k < M; ++k) acc[k] += a0 * (double)v2[k].x + a1 * (double)v2[k].y;
            acc[M] += a0 * a0 + a1 * a1;
          }
        else
          {
            T s = w[i];
            T vk[M];
#pragma unroll
            for (int k = 0; k < M; ++k)
              {
                vk[k] = mc.v[k][i];
                s += mc.c[k] * vk[k];
              }
            w[i] = s;
            if (!(mask.active && dot_masked(mask, i)))
              {
                const double sd = (double)s;
#pragma unroll
                for (int k = 0; k < M; ++k) acc[k] += sd * (double)vk[k];
                acc[M] += sd * sd;
              }
          }
      }
    block_reduce_atomic<M + 1>(acc, out);
  }

  // ------------------------------------------------------------------ host wrappers
  inline int grid_for(stfem_ctx *ctx, long long n, int threads)
  {
    long long g = (n + threads - 1) / threads;
    const long long cap = (long long)ctx->sm_count * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
  }

  template <typename T>
  inline void v_axpy(BlockVec<T> &y, T a, const BlockVec<T> &x)
  {
    k_axpy<T><<<grid_for(y.ctx, y.size(), 256), 256, 0, y.ctx->stream>>>(y.size(), a, x.d, y.d);
    y.ctx->launches++;
  }
  template <typename T>
  inline void v_sadd(BlockVec<T> &y, T s, T a, const BlockVec<T> &x)
  {
    k_sadd<T><<<grid_for(y.ctx, y.size(), 256), 256, 0, y.ctx->stream>>>(y.size(), s, a, x.d, y.d);
    y.ctx->launches++;
  }
  template <typename T>
  inline void v_lincomb(BlockVec<T> &z, T a, const BlockVec<T> &x, T b, const BlockVec<T> &y)
  {
    k_lincomb<T><<<grid_for(z.ctx, z.size(), 256), 256, 0, z.ctx->stream>>>(z.size(), a, x.d, b, y.d, z.d);
    z.ctx->launches++;
  }
  template <typename T>
  inline void v_scale(BlockVec<T> &y, T a)
  {
    k_scale<T><<<grid_for(y.ctx, y.size(), 256), 256, 0, y.ctx->stream>>>(y.size(), a, y.d);
    y.ctx->launches++;
  }
  template <typename T>
  inline int v_copy(BlockVec<T> &y, const BlockVec<T> &x)
  {
    STFEM_CUDA_CHECK(cudaMemcpyAsync(y.d, x.d, sizeof(T) * (size_t)x.size(), cudaMemcpyDeviceToDevice, y.ctx->stream));
    return STFEM_OK;
  }
  template <typename TO, typename TI>
  inline void v_convert(BlockVec<TO> &y, const BlockVec<TI> &x)
  {
    k_convert<TO, TI><<<grid_for(y.ctx, y.size(), 256), 256, 0, y.ctx->stream>>>(y.size(), x.d, y.d);
    y.ctx->launches++;
  }

  // host-visible scratch for reductions (pinned), one per context user
  struct DotScratch
  {
    double *d = nullptr, *h = nullptr;
    DotMask mask; // set by the owner for partitioned meshes
    void    set_partition(const PartitionInfo &part, const int np[3], int dim)
    {
      mask         = DotMask();
      mask.active  = part.active ? 1 : 0;
      mask.n_block = 1;
      for (int d = 0; d < 3; ++d)
        {
          mask.np[d] = d < dim ? np[d] : 1;
          mask.n_block *= mask.np[d];
          if (d < dim && part.neighbor[d][1] >= 0) mask.skip_high |= 1u << d;
        }
    }
    int init()
    {
      if (d) return STFEM_OK;
      STFEM_CUDA_CHECK(cudaMalloc(&d, sizeof(double) * 256));
      STFEM_CUDA_CHECK(cudaMallocHost(&h, sizeof(double) * 256));
      return STFEM_OK;
    }
    ~DotScratch()
    {
      if (d) cudaFree(d);
      if (h) cudaFreeHost(h);
    }
  };

  // the specialised Gram-Schmidt kernels read pairs of entries: every vector must start on a 16-byte boundary
  // (STFEM_GS_GENERIC=1 keeps the generic kernels, for A/B timing)
  template <typename T>
  inline bool gs_specialised(const T *w, const MultiCoef<T> &mc)
  {
    static const bool generic = std::getenv("STFEM_GS_GENERIC") != nullptr;
    if (generic) return false;
    bool ok = ((unsigned long long)w & 15ull) == 0;
    for (int k = 0; k < mc.m; ++k) ok = ok && ((unsigned long long)mc.v[k] & 15ull) == 0;
    return ok;
  }
  // out[k] = <w, V[k]> for k < m (any m: chunks of MAXK).  Synchronises the stream.
  template <typename T>
  inline int v_multi_dot(DotScratch &sc, const BlockVec<T> &w, const std::vector<const BlockVec<T> *> &V, double *out)
  {
    STFEM_FORWARD(sc.init());
    stfem_ctx *ctx = w.ctx;
    const int  m   = (int)V.size();
    if (m > 256) { set_error("multi_dot: too many vectors"); return STFEM_ERR_INVALID; }
    STFEM_CUDA_CHECK(cudaMemsetAsync(sc.d, 0, sizeof(double) * m, ctx->stream));
    for (int k0 = 0; k0 < m; k0 += MAXK)
      {
        MultiCoef<T> mc;
        mc.m = std::min(MAXK, m - k0);
        for (int k = 0; k < mc.m; ++k) mc.v[k] = V[k0 + k]->d;
        const int grid = grid_for(ctx, (w.size() + 1) / 2, 256);
        switch (gs_specialised(w.d, mc) ? mc.m : 0)
          {
#define STFEM_MD_CASE(M_) case M_: k_multi_dot_m<T, M_><<<grid, 256, 0, ctx->stream>>>(w.size(), mc, w.d, sc.d + k0, sc.mask); break;
            STFEM_MD_CASE(1) STFEM_MD_CASE(2) STFEM_MD_CASE(3) STFEM_MD_CASE(4) STFEM_MD_CASE(5) STFEM_MD_CASE(6) STFEM_MD_CASE(7) STFEM_MD_CASE(8)
            STFEM_MD_CASE(9) STFEM_MD_CASE(10)
#undef STFEM_MD_CASE
            default: k_multi_dot<T><<<grid_for(ctx, w.size(), 256), 256, 0, ctx->stream>>>(w.size(), mc, w.d, sc.d + k0, sc.mask);
          }
        ctx->launches++;
      }
    if (sc.mask.active && ctx->n_ranks > 1 && ctx->nccl_comm) // vectors on agglomerated (global) levels are complete on every rank
      {
        NcclApi *api = nccl_api();
        STFEM_REQUIRE(api, "multi_dot: NCCL unavailable");
        STFEM_NCCL_CHECK(api->AllReduce(sc.d, sc.d, (size_t)m, NcclApi::kDouble, NcclApi::kSum, (nccl_comm_t)ctx->nccl_comm, ctx->stream));
      }
    STFEM_CUDA_CHECK(cudaMemcpyAsync(sc.h, sc.d, sizeof(double) * m, cudaMemcpyDeviceToHost, ctx->stream));
    STFEM_FORWARD(stream_sync_checked(ctx, "multi_dot"));
    for (int k = 0; k < m; ++k) out[k] = sc.h[k];
    return STFEM_OK;
  }
  template <typename T>
  inline int v_dot(DotScratch &sc, const BlockVec<T> &x, const BlockVec<T> &y, double *out)
  {
    std::vector<const BlockVec<T> *> V{&y};
    return v_multi_dot(sc, x, V, out);
  }
  // w += sum_k c[k] V[k]
  template <typename T>
  inline void v_multi_axpy(BlockVec<T> &w, const std::vector<const BlockVec<T> *> &V, const double *c)
  {
    const int m = (int)V.size();
    for (int k0 = 0; k0 < m; k0 += MAXK)
      {
        MultiCoef<T> mc;
        mc.m = std::min(MAXK, m - k0);
        for (int k = 0; k < mc.m; ++k)
          {
            mc.v[k] = V[k0 + k]->d;
            mc.c[k] = (T)c[k0 + k];
          }
        k_multi_axpy<T><<<grid_for(w.ctx, w.size(), 256), 256, 0, w.ctx->stream>>>(w.size(), mc, w.d);
        w.ctx->launches++;
      }
  }
  // w += sum_k c[k] V[k] (<= MAXK vectors per launch; the last launch also accumulates ||w||^2), returns the norm^2
  template <typename T>
  inline int v_multi_axpy_norm(DotScratch &sc, BlockVec<T> &w, const std::vector<const BlockVec<T> *> &V, const double *c, double *nrm2)
  {
    STFEM_FORWARD(sc.init());
    stfem_ctx *ctx = w.ctx;
    const int  m   = (int)V.size();
    STFEM_CUDA_CHECK(cudaMemsetAsync(sc.d, 0, sizeof(double), ctx->stream));
    for (int k0 = 0; k0 < m || k0 == 0; k0 += MAXK)
      {
        MultiCoef<T> mc;
        mc.m = std::max(0, std::min(MAXK, m - k0));
        for (int k = 0; k < mc.m; ++k)
          {
            mc.v[k] = V[k0 + k]->d;
            mc.c[k] = (T)c[k0 + k];
          }
        const bool last = k0 + MAXK >= m;
        if (last)
          k_multi_axpy_norm<T><<<grid_for(ctx, w.size(), 256), 256, 0, ctx->stream>>>(w.size(), mc, w.d, sc.d, sc.mask.active, sc.mask.np[0], sc.mask.np[1],
                                                                                       sc.mask.np[2], sc.mask.skip_high, sc.mask.n_block);
        else
          k_multi_axpy<T><<<grid_for(ctx, w.size(), 256), 256, 0, ctx->stream>>>(w.size(), mc, w.d);
        ctx->launches++;
        if (last) break;
      }
    if (sc.mask.active && ctx->n_ranks > 1 && ctx->nccl_comm)
      {
        NcclApi *api = nccl_api();
        STFEM_REQUIRE(api, "multi_axpy_norm: NCCL unavailable");
        STFEM_NCCL_CHECK(api->AllReduce(sc.d, sc.d, 1, NcclApi::kDouble, NcclApi::kSum, (nccl_comm_t)ctx->nccl_comm, ctx->stream));
      }
    STFEM_CUDA_CHECK(cudaMemcpyAsync(sc.h, sc.d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    STFEM_FORWARD(stream_sync_checked(ctx, "multi_axpy_norm"));
    *nrm2 = sc.h[0];
    return STFEM_OK;
  }
  // w += sum_k c[k] V[k] (at most MAXK vectors), then out[k] = <w, V[k]> and out[m] = ||w||^2 of the updated w.  Synchronises.
  template <typename T>
  inline int v_multi_axpy_dot(DotScratch &sc, BlockVec<T> &w, const std::vector<const BlockVec<T> *> &V, const double *c, double *out)
  {
    STFEM_FORWARD(sc.init());
    stfem_ctx *ctx = w.ctx;
    const int  m   = (int)V.size();
    STFEM_REQUIRE(m <= MAXK, "multi_axpy_dot: too many vectors");
    STFEM_CUDA_CHECK(cudaMemsetAsync(sc.d, 0, sizeof(double) * (m + 1), ctx->stream));
    MultiCoef<T> mc;
    mc.m = m;
    for (int k = 0; k < m; ++k)
      {
        mc.v[k] = V[k]->d;
        mc.c[k] = (T)c[k];
      }
    const int grid = grid_for(ctx, (w.size() + 1) / 2, 256);
    switch (gs_specialised(w.d, mc) ? m : 0)
      {
#define STFEM_MAD_CASE(M_) case M_: k_multi_axpy_dot_m<T, M_><<<grid, 256, 0, ctx->stream>>>(w.size(), mc, w.d, sc.d, sc.mask); break;
        STFEM_MAD_CASE(1) STFEM_MAD_CASE(2) STFEM_MAD_CASE(3) STFEM_MAD_CASE(4) STFEM_MAD_CASE(5) STFEM_MAD_CASE(6) STFEM_MAD_CASE(7) STFEM_MAD_CASE(8)
        STFEM_MAD_CASE(9)
#undef STFEM_MAD_CASE
        default: k_multi_axpy_dot<T><<<grid_for(ctx, w.size(), 256), 256, 0, ctx->stream>>>(w.size(), mc, w.d, sc.d, sc.mask);
      }
    ctx->launches++;
    if (sc.mask.active && ctx->n_ranks > 1 && ctx->nccl_comm)
      {
        NcclApi *api = nccl_api();
        STFEM_REQUIRE(api, "multi_axpy_dot: NCCL unavailable");
        STFEM_NCCL_CHECK(api->AllReduce(sc.d, sc.d, (size_t)(m + 1), NcclApi::kDouble, NcclApi::kSum, (nccl_comm_t)ctx->nccl_comm, ctx->stream));
      }
    STFEM_CUDA_CHECK(cudaMemcpyAsync(sc.h, sc.d, sizeof(double) * (m + 1), cudaMemcpyDeviceToHost, ctx->stream));
    STFEM_FORWARD(stream_sync_checked(ctx, "multi_axpy_dot"));
    for (int k = 0; k <= m; ++k) out[k] = sc.h[k];
    return STFEM_OK;
  }
  // out = scale * (w + sum_k c[k] V[k]) for at most MAXK vectors (the caller falls back to separate passes beyond)
  template <typename T>
  inline void v_multi_axpy_scale_out(BlockVec<T> &out, const BlockVec<T> &w, const std::vector<const BlockVec<T> *> &V, const double *c, double scale)
  {
    MultiCoef<T> mc;
    mc.m = (int)V.size();
    for (int k = 0; k < mc.m; ++k)
      {
        mc.v[k] = V[k]->d;
        mc.c[k] = (T)c[k];
      }
    k_multi_axpy_scale_out<T><<<grid_for(w.ctx, w.size(), 256), 256, 0, w.ctx->stream>>>(w.size(), mc, (T)scale, w.d, out.d);
    w.ctx->launches++;
  }
  template <typename T>
  inline void v_scale_copy(BlockVec<T> &y, T a, const BlockVec<T> &x)
  {
    k_scale_copy<T><<<grid_for(y.ctx, y.size(), 256), 256, 0, y.ctx->stream>>>(y.size(), a, x.d, y.d);
    y.ctx->launches++;
  }
  // dst (+)= P src across blocks; dst has P.m blocks, src P.n blocks, same n
  template <typename T>
  inline void v_block_matmul(BlockVec<T> &dst, const std::vector<double> &P, int m, int n, const BlockVec<T> &src, bool add)
  {
    SmallMat<T> S;
    S.m = m; S.n = n;
    for (int i = 0; i < m * n; ++i) S.a[i] = (T)P[i];
    k_block_matmul<T><<<grid_for(dst.ctx, dst.n, 256), 256, 0, dst.ctx->stream>>>(dst.n, S, src.d, dst.d, add ? 1 : 0);
    dst.ctx->launches++;
  }
} // namespace stfem
