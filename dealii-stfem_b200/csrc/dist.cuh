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

  // the same on the two faces normal to `axis` only (one launch per direction with interfaces: a node on an edge or corner is
  // halved once per direction it is shared in); work proportional to the faces, not to the vector
  template <typename T>
  __global__ void k_scale_interface_faces(BlockPtrs blocks, int nb, int np0, int np1, int np2, int axis, unsigned sides)
  {
    const int       a = axis == 0 ? np1 : np0, b = axis == 2 ? np1 : np2; // plane extents (fast, slow)
    const int       nd = axis == 0 ? np0 : (axis == 1 ? np1 : np2);
    const long long per = (long long)a * b, total = per * nb * 2;
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x)
      {
        const int side = (int)(gid / (per * nb));
        if (!((sides >> side) & 1u) || (side == 1 && nd == 1)) continue;
        const long long g   = gid % (per * nb);
        const int       blk = (int)(g / per);
        const long long r   = g % per;
        const int       u = (int)(r % a), v = (int)(r / a), index = side == 0 ? 0 : nd - 1;
        long long       off;
        if (axis == 0) off = (long long)index + (long long)np0 * (u + (long long)np1 * v);
        else if (axis == 1) off = (long long)u + (long long)np0 * (index + (long long)np1 * v);
        else off = (long long)u + (long long)np0 * (v + (long long)np1 * index);
        ((T *)blocks.p[blk])[off] *= T(0.5);
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

  // ---- single-round exchange: every rank of the 3x3x3 neighbourhood that shares nodes with this brick (faces, edges,
  // corners) gets this rank's partial sums of exactly the shared nodes in ONE grouped send/receive; the receiver then adds
  // the up to 8 partial sums of a node IN THE ORDER OF THE GLOBAL RANKS, so that every copy of an interface DoF ends up
  // with bit-identical values (a sum in arrival order would differ in the last bit between the ranks).
  struct HaloPlan
  {
    int       n_seg = 0;            // neighbours sharing nodes with this brick
    int       rank[26];             // their ranks
    int       off[26][3];           // their offsets in the process grid, each in {-1, 0, 1}
    int       lo[26][3], ext[26][3]; // node box shared with them
    long long start[27];            // element offsets of the segments in the send / receive buffer (all blocks)
    int       seg_of[27];           // (ox+1) + 3 (oy+1) + 9 (oz+1) -> segment index or -1
    int       np[3] = {0, 0, 0}, nb = 0, my_rank = 0;
    unsigned  has = 0;              // bit 2d+s: a neighbour exists on side s of direction d
  };

  template <typename T>
  __host__ __device__ inline void halo_pack_element(const BlockPtrs &blocks, const HaloPlan &pl, long long gid, T *out)
  {
    int s = 0;
    while (s + 1 < pl.n_seg && gid >= pl.start[s + 1]) ++s;
    const long long e   = gid - pl.start[s];
    const long long box = (long long)pl.ext[s][0] * pl.ext[s][1] * pl.ext[s][2];
    const int       blk = (int)(e / box);
    long long       r   = e % box;
    const int       u = (int)(r % pl.ext[s][0]);
    r /= pl.ext[s][0];
    const int       v = (int)(r % pl.ext[s][1]), w = (int)(r / pl.ext[s][1]);
    const long long node = (long long)(pl.lo[s][0] + u) + (long long)pl.np[0] * ((pl.lo[s][1] + v) + (long long)pl.np[1] * (pl.lo[s][2] + w));
    out[gid] = ((const T *)blocks.p[blk])[node];
  }

  // the element of the segment whose offset equals the node's full set of interface flags sums all partial sums of that
  // node (own + every neighbour sharing it) in rank order and writes the result; all other elements do nothing
  template <typename T>
  __host__ __device__ inline void halo_unpack_element(const BlockPtrs &blocks, const HaloPlan &pl, long long gid, const T *in)
  {
    int s = 0;
    while (s + 1 < pl.n_seg && gid >= pl.start[s + 1]) ++s;
    const long long e   = gid - pl.start[s];
    const long long box = (long long)pl.ext[s][0] * pl.ext[s][1] * pl.ext[s][2];
    const int       blk = (int)(e / box);
    long long       r   = e % box;
    int             c[3];
    c[0] = pl.lo[s][0] + (int)(r % pl.ext[s][0]);
    r /= pl.ext[s][0];
    c[1] = pl.lo[s][1] + (int)(r % pl.ext[s][1]);
    c[2] = pl.lo[s][2] + (int)(r / pl.ext[s][1]);
    int f[3];
    for (int d = 0; d < 3; ++d)
      {
        f[d] = (c[d] == 0 && ((pl.has >> (2 * d)) & 1u)) ? -1 : ((c[d] == pl.np[d] - 1 && ((pl.has >> (2 * d + 1)) & 1u)) ? 1 : 0);
        if (f[d] != pl.off[s][d]) return;
      }
    T  *dst = (T *)blocks.p[blk] + ((long long)c[0] + (long long)pl.np[0] * (c[1] + (long long)pl.np[1] * c[2]));
    int ranks[8];
    T   vals[8];
    int n    = 1;
    ranks[0] = pl.my_rank;
    vals[0]  = *dst;
    for (int m = 1; m < 8; ++m) // non-empty subsets of the node's interface directions
      {
        int  o[3];
        bool ok = true;
        for (int d = 0; d < 3; ++d)
          {
            o[d] = ((m >> d) & 1) ? f[d] : 0;
            if (((m >> d) & 1) && f[d] == 0) ok = false;
          }
        if (!ok) continue;
        const int t = pl.seg_of[(o[0] + 1) + 3 * (o[1] + 1) + 9 * (o[2] + 1)];
        if (t < 0) continue;
        const long long bt = (long long)pl.ext[t][0] * pl.ext[t][1] * pl.ext[t][2];
        const long long et = (long long)(c[0] - pl.lo[t][0]) + (long long)pl.ext[t][0] * ((c[1] - pl.lo[t][1]) + (long long)pl.ext[t][1] * (c[2] - pl.lo[t][2]));
        ranks[n] = pl.rank[t];
        vals[n]  = in[pl.start[t] + (long long)blk * bt + et];
        ++n;
      }
    for (int a = 1; a < n; ++a) // insertion sort by rank
      {
        const int ra = ranks[a];
        const T   va = vals[a];
        int       b  = a - 1;
        while (b >= 0 && ranks[b] > ra)
          {
            ranks[b + 1] = ranks[b];
            vals[b + 1]  = vals[b];
            --b;
          }
        ranks[b + 1] = ra;
        vals[b + 1]  = va;
      }
    T sum = vals[0];
    for (int a = 1; a < n; ++a) sum += vals[a];
    *dst = sum;
  }

  template <typename T>
  __global__ void k_halo_pack(BlockPtrs blocks, HaloPlan pl, T *__restrict__ out)
  {
    const long long total = pl.start[pl.n_seg];
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x)
      halo_pack_element<T>(blocks, pl, gid, out);
  }
  template <typename T>
  __global__ void k_halo_unpack_sum(BlockPtrs blocks, HaloPlan pl, const T *__restrict__ in)
  {
    const long long total = pl.start[pl.n_seg];
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x)
      halo_unpack_element<T>(blocks, pl, gid, in);
  }

  // ---- the same exchange without NCCL: every rank maps its neighbours' receive buffers (CUDA IPC, NVLink peer access) and
  // the pack kernel STORES the partial sums straight into them; a one-warp kernel then raises a sequence flag in every
  // neighbour's memory, and the unpack kernel waits for the flags of all neighbours before it sums.  Two receive buffers
  // alternate (a neighbour can be at most one exchange ahead, see capi_dist.cu), the sequence number lives in device memory
  // so that the three kernels replay inside CUDA graphs.  One exchange = 3 small kernels and ~2 NVLink latencies instead of
  // the launch + proxy latency of a grouped ncclSend / ncclRecv.
  struct HaloP2PArgs
  {
    void               *peer_data[26];  // neighbours' receive areas (both parities), mapped into this process
    unsigned           *peer_flag[26];  // the flag in the neighbour's memory this rank raises
    long long           peer_start[26]; // where this rank's segment starts in the neighbour's receive area (elements)
    long long           peer_total[26]; // elements per parity in the neighbour's receive area
    void               *my_data;        // this rank's receive area: [2][total]
    volatile unsigned  *my_flags;       // [26] raised by the neighbours
    unsigned           *seq;            // number of exchanges completed so far
  };

  template <typename T>
  __global__ void k_halo_pack_p2p(BlockPtrs blocks, HaloPlan pl, HaloP2PArgs pa)
  {
    const long long total  = pl.start[pl.n_seg];
    const unsigned  parity = (*pa.seq + 1u) & 1u;
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x)
      {
        int s = 0;
        while (s + 1 < pl.n_seg && gid >= pl.start[s + 1]) ++s;
        const long long e   = gid - pl.start[s];
        const long long box = (long long)pl.ext[s][0] * pl.ext[s][1] * pl.ext[s][2];
        const int       blk = (int)(e / box);
        long long       r   = e % box;
        const int       u = (int)(r % pl.ext[s][0]);
        r /= pl.ext[s][0];
        const int       v = (int)(r % pl.ext[s][1]), w = (int)(r / pl.ext[s][1]);
        const long long node = (long long)(pl.lo[s][0] + u) + (long long)pl.np[0] * ((pl.lo[s][1] + v) + (long long)pl.np[1] * (pl.lo[s][2] + w));
        ((T *)pa.peer_data[s])[(long long)parity * pa.peer_total[s] + pa.peer_start[s] + e] = ((const T *)blocks.p[blk])[node];
      }
  }
  // after the pack kernel (stream order): make its stores visible system-wide, raise the flags, count the exchange
  static __global__ void k_halo_signal_p2p(HaloPlan pl, HaloP2PArgs pa)
  {
    __threadfence_system();
    const unsigned next = *pa.seq + 1u;
    __syncwarp();
    if ((int)threadIdx.x < pl.n_seg)
      {
        *(volatile unsigned *)pa.peer_flag[threadIdx.x] = next;
        __threadfence_system();
      }
    __syncwarp();
    if (threadIdx.x == 0) *pa.seq = next;
  }
  template <typename T>
  __global__ void k_halo_unpack_sum_p2p(BlockPtrs blocks, HaloPlan pl, HaloP2PArgs pa)
  {
    const unsigned expected = *pa.seq;
    if ((int)threadIdx.x < pl.n_seg)
      {
        const long long t0 = clock64();
        while ((int)(pa.my_flags[threadIdx.x] - expected) < 0)
          if (clock64() - t0 > 20000000000ll) __trap(); // ~10 s: a neighbour never sent - fail instead of hanging the GPU
      }
    __syncthreads();
    __threadfence_system();
    const long long total = pl.start[pl.n_seg];
    const T        *in    = (const T *)pa.my_data + (long long)(expected & 1u) * total;
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x)
      halo_unpack_element<T>(blocks, pl, gid, in);
  }

  struct HaloP2P
  {
    bool        tried = false, ok = false;
    char       *local = nullptr; // [flags 256 B][seq 256 B][receive area: 2 x total elements]
    size_t      elem = 0;
    void       *peer_base[26];
    int         n_peer = 0;
    HaloP2PArgs args;
    ~HaloP2P()
    {
      for (int i = 0; i < n_peer; ++i)
        if (peer_base[i]) cudaIpcCloseMemHandle(peer_base[i]);
      if (local) cudaFree(local);
    }
  };

  struct HaloBuffers
  {
    void  *send[2] = {nullptr, nullptr}, *recv[2] = {nullptr, nullptr};
    size_t bytes = 0;
    void  *send_all = nullptr, *recv_all = nullptr; // single-round exchange
    size_t bytes_all = 0;
    HaloPlan plan;                                  // cached for (np, nb)
    HaloP2P  p2p;                                   // peer-memory exchange of that plan
    bool     no_p2p = false;                        // one-off exchanges (set-up): NCCL send / receive, no peer-memory set-up
    ~HaloBuffers()
    {
      for (int s = 0; s < 2; ++s)
        {
          if (send[s]) cudaFree(send[s]);
          if (recv[s]) cudaFree(recv[s]);
        }
      if (send_all) cudaFree(send_all);
      if (recv_all) cudaFree(recv_all);
    }
  };

  // sum the partial values of the interface DoFs over the ranks sharing them (compress(add) + ghost update)
  template <typename T>
  int halo_compress_add(stfem_ctx *ctx, const Partition &part, HaloBuffers &hb, void *const *blocks, int nb, const int np[3], int dim,
                        cudaStream_t stream = nullptr);
  template <typename T>
  int halo_scale_interfaces(stfem_ctx *ctx, const Partition &part, HaloBuffers &hb, void *const *blocks, int nb, const int np[3], int dim);
} // namespace stfem
