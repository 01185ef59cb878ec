// Host side of the brick kernel (st_vmult_brick.cuh): tensor maps of the source blocks, launch.
#pragma once
#include <cstdlib>

#include "op.hpp"
#include "st_vmult_brick.cuh"

namespace stfem
{
  // cuTensorMapEncodeTiled through the runtime's driver entry point: no link dependency on libcuda
  typedef CUresult (*brick_encode_fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  inline brick_encode_fn_t brick_encode_fn()
  {
    static brick_encode_fn_t fn = nullptr;
    static bool              tried = false;
    if (!tried)
      {
        tried = true;
        void                           *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
          fn = (brick_encode_fn_t)p;
        else
          cudaGetLastError();
      }
    return fn;
  }

  template <typename T>
  inline bool brick_encode(const BrickMapDesc &d, CUtensorMap *map)
  {
    brick_encode_fn_t fn = brick_encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2]    = {d.dim0, d.dim1};
    const cuuint64_t strides[1] = {d.stride1};
    const cuuint32_t box[2]     = {(cuuint32_t)d.box0, (cuuint32_t)d.box1};
    const cuuint32_t estr[2]    = {1, 1};
    const CUresult   r = fn(map, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(d.base), dims,
                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
  }

  // Alpha / Beta by value: the host copy behind one of the operator's device matrices (nullptr: unknown pointer)
  inline const std::vector<double> *brick_host_matrix(const stfem_op *op, const void *dev)
  {
    if (dev == op->d_alpha) return &op->Alpha;
    if (dev == op->d_beta) return &op->Beta;
    if (dev == op->d_alphaT) return &op->AlphaT;
    if (dev == op->d_betaT) return &op->BetaT;
    if (dev == op->d_alpha_neg) return &op->AlphaNeg;
    if (dev == op->d_beta_neg) return &op->BetaNeg;
    return nullptr;
  }

  template <int N1, int NB, typename T, int CX, int CY, int MINB, bool SPLIT = false>
  static int launch_brick(stfem_op *op, void *const *dst, const void *const *src, const std::vector<double> &Alpha, const std::vector<double> &Beta,
                          int mode, const void *const *rhs, int zlo, int zhi, bool first_plane_acc, int use_tma, int n_chunks)
  {
    using C = BrickCfg<T, N1, NB, CX, CY, SPLIT>;
    stfem_mesh *m = op->mesh;
    if (zhi <= zlo) return STFEM_OK;
    BrickArgs<T, N1, NB> a;
    std::memset(&a, 0, sizeof(a));
    const ShapeHost &sh = *op->shape;
    double           h[3];
    for (int d = 0; d < 3; ++d) h[d] = (m->upper[d] - m->lower[d]) / m->n[d];
    unsigned iface = 0; // faces shared with other ranks
    if (m->part.active)
      for (int d = 0; d < 3; ++d)
        for (int sd = 0; sd < 2; ++sd)
          if (m->part.neighbor[d][sd] >= 0) iface |= 1u << (2 * d + sd);
    brick_fill_args<T, N1, NB, CX, CY>(a, sh.S.data(), sh.D.data(), sh.wq.data(), h, m->n, m->dirichlet, Alpha.data(), Beta.data(), src, dst, zlo, zhi,
                                       mode, rhs, first_plane_acc, n_chunks, (long long)m->ctx->sm_count * MINB, iface);
    STFEM_REQUIRE(a.n_cls <= C::MAXCLS, "st_vmult (brick): %d row classes", a.n_cls);
    a.use_tma = use_tma;
    // per-SM CTA counter behind the alternating warp roles: opt-in (kernel_variant 90); measured 2 % SLOWER than without
    // (0.920 against 0.903 ms, Q4 x 2 blocks FP64, 96^3 cells), so the schedulers are not what the X phase waits for
    a.sm_counter = nullptr;
    if (!SPLIT && op->variant == 90 && C::NWARPS > C::XW && C::NWARPS % 4 == 0)
      {
        stfem_ctx *ctx = m->ctx;
        if (!ctx->d_sm_counter)
          {
            STFEM_CUDA_CHECK(cudaMalloc(&ctx->d_sm_counter, 1024 * sizeof(unsigned)));
            STFEM_CUDA_CHECK(cudaMemsetAsync(ctx->d_sm_counter, 0, 1024 * sizeof(unsigned), ctx->stream));
          }
        a.sm_counter = ctx->d_sm_counter;
      }
    for (int b = 0; b < NB && a.use_tma; ++b)
      for (int c = 0; c < a.n_cls; ++c)
        if (!brick_encode<T>(a.desc[b][c], &a.maps[b][c]))
          {
            a.use_tma = 0; // the driver refused the descriptor: plain loads
            break;
          }
    cudaStream_t stream = op->launch_stream ? op->launch_stream : m->ctx->stream;
    if (a.n_chunks > 1 && mode != 1)
      {
        BrickZeroArgs<T, NB> z;
        for (int b = 0; b < NB; ++b) z.dst[b] = (T *)dst[b];
        z.plane  = (long long)a.np[0] * a.np[1];
        z.first  = (N1 - 1) * (zlo + a.layers_per_chunk);
        z.stride = (N1 - 1) * a.layers_per_chunk;
        z.count  = a.n_chunks - 1;
        const long long total = z.plane * z.count * NB;
        brick_zero_planes_kernel<T, NB><<<(unsigned)std::min<long long>((total + 255) / 256, (long long)m->ctx->sm_count * 8), 256, 0, stream>>>(z);
        m->ctx->launches++;
      }
    const size_t smem = (size_t)C::smem_bytes(a.n_cls);
    const bool   gen  = mode != 0 || first_plane_acc; // plain product: the instance without the read / mode logic in the store phase
    auto         kern = gen ? st_vmult_brick_kernel<T, N1, NB, CX, CY, MINB, SPLIT, true> : st_vmult_brick_kernel<T, N1, NB, CX, CY, MINB, SPLIT, false>;
    STFEM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid   = (long long)a.tiles_x * a.tiles_y * a.n_chunks;
    STFEM_REQUIRE(grid < (1ll << 31), "st_vmult (brick): grid too large");
    kern<<<(unsigned)grid, C::NTHREADS, smem, stream>>>(a);
    m->ctx->launches++;
    STFEM_CUDA_CHECK(cudaGetLastError());
    return STFEM_OK;
  }
} // namespace stfem
