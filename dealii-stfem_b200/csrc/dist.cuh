// Multi-GPU layer: one process per GPU, box partition of the structured hex mesh, shared interface DoFs
// duplicated on both sides.  Replaces what deal.II does for the reference inside MatrixFree::cell_loop
// (ghost update + compress(add), called at include/operators.h:1016), in PreconditionVanka::vmult
// (include/stmg.h:843, 870-871) and in the MPI_Allreduce of every dot product.
//
//   compress_add : after a local cell loop every interface DoF holds a partial sum; the partial sums of the
//                  ranks sharing it are added so that all copies hold the full value.  Done direction by
//                  direction (x, then y, then z faces, each including the already summed edges), i.e. 6
//                  messages per rank instead of 26; ncclSend/ncclRecv grouped per direction over NVLink.
//   dot products : interface DoFs are counted on the lower rank only, then ncclAllReduce(sum, double).
// NCCL is loaded with dlopen("libnccl.so.2") so the single-GPU library has no link dependency on it.
#pragma once
#include <dlfcn.h>

#include "common.hpp"

namespace stfem
{
  // minimal NCCL surface (ABI-stable C functions)
  typedef struct ncclComm *nccl_comm_t;
  struct NcclUniqueId { char internal[128]; };
  struct NcclApi
  {
    void *handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    static constexpr int kDouble = 8, kFloat = 7, kChar = 0, kSum = 0; // ncclDataType_t / ncclRedOp_t values
  };
  NcclApi *nccl_api(); // loads on first use; nullptr + error message if libnccl is unavailable

#define STFEM_NCCL_CHECK(expr)                                                                     \
  do                                                                                               \
    {                                                                                              \
      int r__ = (expr);                                                                            \
      if (r__ != 0)                                                                                \
        {                                                                                          \
          stfem::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                      \
                           stfem::nccl_api()->GetErrorString(r__));                                \
          return STFEM_ERR_CUDA;                                                                   \
        }                                                                                          \
    }                                                                                              \
  while (0)

  using Partition = PartitionInfo;

  // block pointers by value (kernel parameter): no host->device pointer upload, so the exchange can be captured
  // in a CUDA graph
  struct BlockPtrs { void *p[STFEM_MAX_BLOCKS]; };

  // pack / unpack-add one index plane of a [np2][np1][np0] array (all blocks), plane normal to `axis`
  template <typename T>
  __global__ void k_pack_plane(BlockPtrs blocks, int nb, int np0, int np1, int np2, int axis, int index, T *__restrict__ out)
  {
    const int       a = axis == 0 ? np1 : np0, b = axis == 2 ? np1 : np2; // plane extents (fast, slow)
    const long long per = (long long)a * b, total = per * nb;
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x)
      {
        const int blk = (int)(gid / per);
        const long long r = gid % per;
        const int u = (int)(r % a), v = (int)(r / a);
        long long off;
        if (axis == 0) off = (long long)index + (long long)np0 * (u + (long long)np1 * v);
        else if (axis == 1) off = (long long)u + (long long)np0 * (index + (long long)np1 * v);
        else off = (long long)u + (long long)np0 * (v + (long long)np1 * index);
        out[gid] = ((const T *)blocks.p[blk])[off];
      }
  }
  template <typename T>
  __global__ void k_unpack_add_plane(BlockPtrs blocks, int nb, int np0, int np1, int np2, int axis, int index, const T *__restrict__ in)
  {
    const int       a = axis == 0 ? np1 : np0, b = axis == 2 ? np1 : np2;
    const long long per = (long long)a * b, total = per * nb;
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x)
      {
        const int blk = (int)(gid / per);
        const long long r = gid % per;
        const int u = (int)(r % a), v = (int)(r / a);
        long long off;
        if (axis == 0) off = (long long)index + (long long)np0 * (u + (long long)np1 * v);
        else if (axis == 1) off = (long long)u + (long long)np0 * (index + (long long)np1 * v);
        else off = (long long)u + (long long)np0 * (v + (long long)np1 * index);
        ((T *)blocks.p[blk])[off] += in[gid];
      }
  }

  // multiply the entries on rank-interface planes by 1/2 per shared direction (so that a sum over ranks of a
  // quantity that is complete on every rank counts it once): used before the restriction of a residual
  template <typename T>
  __global__ void k_scale_interfaces(BlockPtrs blocks, int nb, int np0, int np1, int np2, unsigned shared_faces)
  {
    const long long per = (long long)np0 * np1 * np2, total = per * nb;
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x)
      {
        const int       blk = (int)(gid / per);
        long long       r   = gid % per;
        const int       ix = (int)(r % np0);
        r /= np0;
        const int iy = (int)(r % np1), iz = (int)(r / np1);
        T         w  = T(1);
        if (((shared_faces & 1u) && ix == 0) || ((shared_faces & 2u) && ix == np0 - 1)) w *= T(0.5);
        if (((shared_faces & 4u) && iy == 0) || ((shared_faces & 8u) && iy == np1 - 1)) w *= T(0.5);
        if (((shared_faces & 16u) && iz == 0) || ((shared_faces & 32u) && iz == np2 - 1)) w *= T(0.5);
        if (w != T(1)) ((T *)blocks.p[blk])[gid % per] *= w;
      }
  }

  // brick <-> global copies of the coarse-level agglomeration (csrc/mg.cuh): the local brick [b][bz][by][bx] sits at
  // node offset (o0, o1, o2) of the global array [b][gz][gy][gx]
  template <typename T, bool TO_GLOBAL>
  __global__ void k_brick_global(T *__restrict__ brick, T *__restrict__ global, int nb, int b0, int b1, int b2, int g0, int g1, int g2, int o0,
                                 int o1, int o2)
  {
    const long long per = (long long)b0 * b1 * b2, total = per * nb;
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x)
      {
        const int blk = (int)(gid / per);
        long long r   = gid % per;
        const int ix  = (int)(r % b0);
        r /= b0;
        const int       iy = (int)(r % b1), iz = (int)(r / b1);
        const long long g  = (long long)blk * g0 * g1 * g2 + (long long)(ix + o0) + (long long)g0 * ((iy + o1) + (long long)g1 * (iz + o2));
        if (TO_GLOBAL)
          global[g] += brick[gid];
        else
          brick[gid] = global[g];
      }
  }

  struct HaloBuffers
  {
    void  *send[2] = {nullptr, nullptr}, *recv[2] = {nullptr, nullptr};
    size_t bytes = 0;
    ~HaloBuffers()
    {
      for (int s = 0; s < 2; ++s)
        {
          if (send[s]) cudaFree(send[s]);
          if (recv[s]) cudaFree(recv[s]);
        }
    }
  };

  // sum the partial values of the interface DoFs over the ranks sharing them (compress(add) + ghost update)
  template <typename T>
  int halo_compress_add(stfem_ctx *ctx, const Partition &part, HaloBuffers &hb, void *const *blocks, int nb, const int np[3], int dim,
                        cudaStream_t stream = nullptr);
  template <typename T>
  int halo_scale_interfaces(stfem_ctx *ctx, const Partition &part, HaloBuffers &hb, void *const *blocks, int nb, const int np[3], int dim);
} // namespace stfem
