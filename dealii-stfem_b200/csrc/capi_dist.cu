// C ABI + implementation of the multi-GPU layer (see dist.cuh).
#include <algorithm>
#include <cstdlib>
#include <new>

#include "dist.cuh"
#include "op.hpp"

namespace stfem
{
  static NcclApi g_nccl;
  static bool    g_nccl_tried = false;

  NcclApi *nccl_api()
  {
    if (g_nccl.handle) return &g_nccl;
    if (g_nccl_tried) return nullptr;
    g_nccl_tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h)
      {
        set_error("NCCL not available: %s", dlerror());
        return nullptr;
      }
    bool ok = true;
    auto sym = [&](const char *name) {
      void *p = dlsym(h, name);
      if (!p) ok = false;
      return p;
    };
    g_nccl.GetUniqueId    = (int (*)(NcclUniqueId *))sym("ncclGetUniqueId");
    g_nccl.CommInitRank   = (int (*)(nccl_comm_t *, int, NcclUniqueId, int))sym("ncclCommInitRank");
    g_nccl.CommDestroy    = (int (*)(nccl_comm_t))sym("ncclCommDestroy");
    g_nccl.Send           = (int (*)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t))sym("ncclSend");
    g_nccl.Recv           = (int (*)(void *, size_t, int, int, nccl_comm_t, cudaStream_t))sym("ncclRecv");
    g_nccl.AllReduce      = (int (*)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t))sym("ncclAllReduce");
    g_nccl.Broadcast      = (int (*)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t))sym("ncclBroadcast");
    g_nccl.GroupStart     = (int (*)())sym("ncclGroupStart");
    g_nccl.GroupEnd       = (int (*)())sym("ncclGroupEnd");
    g_nccl.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
    if (!ok)
      {
        set_error("NCCL library lacks required symbols");
        return nullptr;
      }
    g_nccl.handle = h;
    return &g_nccl;
  }

  // the neighbours of this brick in the process grid and the node boxes shared with them
  static void halo_build_plan(HaloPlan &pl, const Partition &part, int nb, const int np[3], int dim, int my_rank)
  {
    pl        = HaloPlan();
    pl.nb     = nb;
    pl.my_rank = my_rank;
    for (int d = 0; d < 3; ++d) pl.np[d] = d < dim ? np[d] : 1;
    for (int i = 0; i < 27; ++i) pl.seg_of[i] = -1;
    for (int d = 0; d < dim; ++d)
      for (int sd = 0; sd < 2; ++sd)
        if (part.neighbor[d][sd] >= 0) pl.has |= 1u << (2 * d + sd);
    pl.start[0] = 0;
    for (int oz = -1; oz <= 1; ++oz)
      for (int oy = -1; oy <= 1; ++oy)
        for (int ox = -1; ox <= 1; ++ox)
          {
            const int o[3] = {ox, oy, oz};
            if (ox == 0 && oy == 0 && oz == 0) continue;
            int  c[3];
            bool ok = true;
            for (int d = 0; d < 3; ++d)
              {
                c[d] = part.coords[d] + o[d];
                if ((d >= dim && o[d] != 0) || c[d] < 0 || c[d] >= part.grid[d]) ok = false;
              }
            if (!ok) continue;
            const int sgm = pl.n_seg++;
            pl.rank[sgm]  = part.rank_of(c);
            long long box = 1;
            for (int d = 0; d < 3; ++d)
              {
                pl.off[sgm][d] = o[d];
                pl.lo[sgm][d]  = o[d] > 0 ? pl.np[d] - 1 : 0;
                pl.ext[sgm][d] = o[d] == 0 ? pl.np[d] : 1;
                box *= pl.ext[sgm][d];
              }
            pl.seg_of[(ox + 1) + 3 * (oy + 1) + 9 * (oz + 1)] = sgm;
            pl.start[sgm + 1] = pl.start[sgm] + box * nb;
          }
  }

  // Peer-memory set-up of a plan (collective over the ranks of the communicator): allocate the receive area, hand its IPC
  // handle + this rank's segment offsets to every neighbour (ncclSend / ncclRecv, once), map theirs.  Falls back to NCCL for
  // ALL ranks if any rank fails (STFEM_HALO_NCCL=1 forces that).
  struct HaloHello
  {
    cudaIpcMemHandle_t handle;
    long long          start, total; // this rank's receive area: where the addressee's segment starts, elements per parity (-1: no peer access here)
    long long          slot;         // the flag of this rank the addressee has to raise
    char               pad[128 - sizeof(cudaIpcMemHandle_t) - 3 * sizeof(long long)];
  };
  static int halo_p2p_setup(stfem_ctx *ctx, HaloBuffers &hb, size_t elem)
  {
    HaloP2P  &pp = hb.p2p;
    HaloPlan &pl = hb.plan;
    NcclApi  *api = nccl_api();
    pp.tried = true;
    pp.ok    = false;
    pp.elem  = elem;
    static const bool force_nccl = std::getenv("STFEM_HALO_NCCL") != nullptr;
    const long long   total = pl.start[pl.n_seg];
    int               good = force_nccl ? 0 : 1;
    cudaStreamSynchronize(ctx->stream);
    if (good)
      {
        const size_t bytes = 512 + (size_t)2 * total * elem;
        if (cudaMalloc(&pp.local, bytes < (2u << 20) ? (2u << 20) : bytes) != cudaSuccess || cudaMemset(pp.local, 0, 512) != cudaSuccess) good = 0;
      }
    std::vector<HaloHello> mine(pl.n_seg), theirs(pl.n_seg);
    HaloHello             *d_buf = nullptr;
    if (cudaMalloc(&d_buf, sizeof(HaloHello) * 2 * (pl.n_seg > 0 ? pl.n_seg : 1)) != cudaSuccess) good = 0;
    for (int sgm = 0; sgm < pl.n_seg; ++sgm)
      {
        std::memset(&mine[sgm], 0, sizeof(HaloHello));
        if (good && cudaIpcGetMemHandle(&mine[sgm].handle, pp.local) != cudaSuccess) good = 0;
        mine[sgm].start = pl.start[sgm];
        mine[sgm].slot  = sgm;
        mine[sgm].total = good ? total : -1; // -1: this rank cannot do it
      }
    cudaGetLastError();
    if (d_buf)
      {
        cudaMemcpy(d_buf, mine.data(), sizeof(HaloHello) * pl.n_seg, cudaMemcpyHostToDevice);
        STFEM_NCCL_CHECK(api->GroupStart());
        for (int sgm = 0; sgm < pl.n_seg; ++sgm)
          {
            STFEM_NCCL_CHECK(api->Send(d_buf + sgm, sizeof(HaloHello), NcclApi::kChar, pl.rank[sgm], (nccl_comm_t)ctx->nccl_comm, ctx->stream));
            STFEM_NCCL_CHECK(api->Recv(d_buf + pl.n_seg + sgm, sizeof(HaloHello), NcclApi::kChar, pl.rank[sgm], (nccl_comm_t)ctx->nccl_comm, ctx->stream));
          }
        STFEM_NCCL_CHECK(api->GroupEnd());
        STFEM_FORWARD(stream_sync_checked(ctx, "halo p2p set-up"));
        cudaMemcpy(theirs.data(), d_buf + pl.n_seg, sizeof(HaloHello) * pl.n_seg, cudaMemcpyDeviceToHost);
      }
    pp.n_peer = 0;
    for (int sgm = 0; sgm < pl.n_seg && good; ++sgm)
      {
        pp.peer_base[sgm] = nullptr;
        if (theirs[sgm].total < 0 || cudaIpcOpenMemHandle(&pp.peer_base[sgm], theirs[sgm].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
          {
            good = 0;
            cudaGetLastError();
            break;
          }
        pp.n_peer = sgm + 1;
        pp.args.peer_data[sgm]  = (char *)pp.peer_base[sgm] + 512;
        pp.args.peer_flag[sgm]  = (unsigned *)pp.peer_base[sgm] + theirs[sgm].slot;
        pp.args.peer_start[sgm] = theirs[sgm].start;
        pp.args.peer_total[sgm] = theirs[sgm].total;
      }
    if (d_buf) cudaFree(d_buf);
    // everybody or nobody
    double flag = good ? 1.0 : 0.0, *d_flag = nullptr;
    STFEM_CUDA_CHECK(cudaMalloc(&d_flag, sizeof(double)));
    STFEM_CUDA_CHECK(cudaMemcpy(d_flag, &flag, sizeof(double), cudaMemcpyHostToDevice));
    STFEM_NCCL_CHECK(api->AllReduce(d_flag, d_flag, 1, NcclApi::kDouble, 3 /* ncclMin */, (nccl_comm_t)ctx->nccl_comm, ctx->stream));
    STFEM_FORWARD(stream_sync_checked(ctx, "halo p2p set-up"));
    STFEM_CUDA_CHECK(cudaMemcpy(&flag, d_flag, sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d_flag);
    pp.ok = flag > 0.5;
    static bool told = false;
    if (!told && std::getenv("STFEM_HALO_VERBOSE"))
      {
        told = true;
        std::fprintf(stderr, "stfem rank %d: interface exchange over %s (%d neighbours, %lld numbers)\n", ctx->rank,
                     pp.ok ? "peer memory (CUDA IPC over NVLink)" : "NCCL send/recv", pl.n_seg, total);
      }
    if (pp.ok)
      {
        pp.args.my_flags = (volatile unsigned *)pp.local;
        pp.args.seq      = (unsigned *)(pp.local + 256);
        pp.args.my_data  = pp.local + 512;
      }
    return STFEM_OK;
  }

  // all interface partial sums in one grouped send / receive (see HaloPlan)
  template <typename T>
  static int halo_exchange_single_round(stfem_ctx *ctx, const Partition &part, HaloBuffers &hb, void *const *blocks, int nb, const int np[3], int dim,
                                        cudaStream_t stream)
  {
    NcclApi *api = nccl_api();
    HaloPlan &pl = hb.plan;
    if (pl.nb != nb || pl.np[0] != np[0] || pl.np[1] != (dim > 1 ? np[1] : 1) || pl.np[2] != (dim > 2 ? np[2] : 1) || pl.n_seg == 0)
      {
        halo_build_plan(pl, part, nb, np, dim, ctx->rank);
        hb.p2p.~HaloP2P();
        new (&hb.p2p) HaloP2P();
      }
    if (pl.n_seg == 0) return STFEM_OK;
    const long long total = pl.start[pl.n_seg];
    if (!hb.p2p.tried || hb.p2p.elem != sizeof(T))
      {
        if (hb.p2p.tried)
          {
            hb.p2p.~HaloP2P();
            new (&hb.p2p) HaloP2P();
          }
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(stream, &cap);
        STFEM_REQUIRE(cap == cudaStreamCaptureStatusNone, "halo exchange: the first exchange of a vector shape must not happen inside a graph capture");
        if (hb.no_p2p)
          {
            hb.p2p.tried = true;
            hb.p2p.ok    = false;
            hb.p2p.elem  = sizeof(T);
          }
        else
          STFEM_FORWARD(halo_p2p_setup(ctx, hb, sizeof(T)));
      }
    if (hb.p2p.ok)
      {
        BlockPtrs bp;
        for (int b = 0; b < STFEM_MAX_BLOCKS; ++b) bp.p[b] = b < nb ? blocks[b] : nullptr;
        const int threads = 256;
        const int grid    = (int)std::min<long long>((total + threads - 1) / threads, (long long)ctx->sm_count * 4);
        k_halo_pack_p2p<T><<<grid, threads, 0, stream>>>(bp, pl, hb.p2p.args);
        k_halo_signal_p2p<<<1, 32, 0, stream>>>(pl, hb.p2p.args);
        k_halo_unpack_sum_p2p<T><<<grid, threads, 0, stream>>>(bp, pl, hb.p2p.args);
        ctx->launches += 3;
        STFEM_CUDA_CHECK(cudaGetLastError());
        return STFEM_OK;
      }
    const size_t    need  = (size_t)total * sizeof(T);
    if (hb.bytes_all < need)
      {
        if (hb.send_all) cudaFree(hb.send_all);
        if (hb.recv_all) cudaFree(hb.recv_all);
        STFEM_CUDA_CHECK(cudaMalloc(&hb.send_all, need));
        STFEM_CUDA_CHECK(cudaMalloc(&hb.recv_all, need));
        hb.bytes_all = need;
      }
    BlockPtrs bp;
    for (int b = 0; b < STFEM_MAX_BLOCKS; ++b) bp.p[b] = b < nb ? blocks[b] : nullptr;
    const int threads = 256;
    const int grid    = (int)std::min<long long>((total + threads - 1) / threads, (long long)ctx->sm_count * 8);
    k_halo_pack<T><<<grid, threads, 0, stream>>>(bp, pl, (T *)hb.send_all);
    ctx->launches++;
    STFEM_NCCL_CHECK(api->GroupStart());
    for (int sgm = 0; sgm < pl.n_seg; ++sgm)
      {
        const size_t cnt = (size_t)(pl.start[sgm + 1] - pl.start[sgm]) * sizeof(T);
        STFEM_NCCL_CHECK(api->Send((const char *)hb.send_all + (size_t)pl.start[sgm] * sizeof(T), cnt, NcclApi::kChar, pl.rank[sgm], (nccl_comm_t)ctx->nccl_comm, stream));
        STFEM_NCCL_CHECK(api->Recv((char *)hb.recv_all + (size_t)pl.start[sgm] * sizeof(T), cnt, NcclApi::kChar, pl.rank[sgm], (nccl_comm_t)ctx->nccl_comm, stream));
      }
    STFEM_NCCL_CHECK(api->GroupEnd());
    k_halo_unpack_sum<T><<<grid, threads, 0, stream>>>(bp, pl, (const T *)hb.recv_all);
    ctx->launches++;
    STFEM_CUDA_CHECK(cudaGetLastError());
    return STFEM_OK;
  }

  template <typename T>
  int halo_compress_add(stfem_ctx *ctx, const Partition &part, HaloBuffers &hb, void *const *blocks, int nb, const int np[3], int dim,
                        cudaStream_t stream)
  {
    if (!part.active) return STFEM_OK;
    if (!stream) stream = ctx->stream;
    NcclApi *api = nccl_api();
    STFEM_REQUIRE(api && ctx->nccl_comm, "halo exchange: context has no NCCL communicator (stfem_ctx_comm_init)");
    STFEM_REQUIRE(nb <= STFEM_MAX_BLOCKS, "halo: too many blocks");
    static const bool rounds = std::getenv("STFEM_HALO_ROUNDS") != nullptr; // the direction-by-direction exchange of round 1
    if (!rounds) return halo_exchange_single_round<T>(ctx, part, hb, blocks, nb, np, dim, stream);
    // buffers sized for the largest face
    size_t maxface = 0;
    for (int d = 0; d < dim; ++d)
      {
        size_t f = 1;
        for (int e = 0; e < dim; ++e)
          if (e != d) f *= (size_t)np[e];
        maxface = std::max(maxface, f);
      }
    const size_t need = maxface * nb * sizeof(T);
    if (hb.bytes < need)
      {
        for (int s = 0; s < 2; ++s)
          {
            if (hb.send[s]) cudaFree(hb.send[s]);
            if (hb.recv[s]) cudaFree(hb.recv[s]);
            STFEM_CUDA_CHECK(cudaMalloc(&hb.send[s], need));
            STFEM_CUDA_CHECK(cudaMalloc(&hb.recv[s], need));
          }
        hb.bytes = need;
      }
    STFEM_REQUIRE(nb <= STFEM_MAX_BLOCKS, "halo: too many blocks");
    BlockPtrs bp;
    for (int b = 0; b < STFEM_MAX_BLOCKS; ++b) bp.p[b] = b < nb ? blocks[b] : nullptr;
    const int np2 = dim == 3 ? np[2] : 1;
    for (int d = 0; d < dim; ++d)
      {
        size_t face = 1;
        for (int e = 0; e < dim; ++e)
          if (e != d) face *= (size_t)np[e];
        const size_t count = face * nb;
        const int    threads = 256;
        const int    grid = (int)std::min<size_t>((count + threads - 1) / threads, (size_t)ctx->sm_count * 8);
        bool any = false;
        for (int s = 0; s < 2; ++s)
          if (part.neighbor[d][s] >= 0)
            {
              k_pack_plane<T><<<grid, threads, 0, stream>>>(bp, nb, np[0], np[1], np2, d, s == 0 ? 0 : np[d] - 1, (T *)hb.send[s]);
              ctx->launches++;
              any = true;
            }
        if (!any) continue;
        STFEM_NCCL_CHECK(api->GroupStart());
        for (int s = 0; s < 2; ++s)
          if (part.neighbor[d][s] >= 0)
            {
              STFEM_NCCL_CHECK(api->Send(hb.send[s], count * sizeof(T), NcclApi::kChar, part.neighbor[d][s], (nccl_comm_t)ctx->nccl_comm, stream));
              STFEM_NCCL_CHECK(api->Recv(hb.recv[s], count * sizeof(T), NcclApi::kChar, part.neighbor[d][s], (nccl_comm_t)ctx->nccl_comm, stream));
            }
        STFEM_NCCL_CHECK(api->GroupEnd());
        for (int s = 0; s < 2; ++s)
          if (part.neighbor[d][s] >= 0)
            {
              k_unpack_add_plane<T><<<grid, threads, 0, stream>>>(bp, nb, np[0], np[1], np2, d, s == 0 ? 0 : np[d] - 1, (const T *)hb.recv[s]);
              ctx->launches++;
            }
      }
    STFEM_CUDA_CHECK(cudaGetLastError());
    return STFEM_OK;
  }

  template <typename T>
  int halo_scale_interfaces(stfem_ctx *ctx, const Partition &part, HaloBuffers &hb, void *const *blocks, int nb, const int np[3], int dim)
  {
    if (!part.active) return STFEM_OK;
    STFEM_REQUIRE(nb <= STFEM_MAX_BLOCKS, "halo: too many blocks");
    BlockPtrs bp;
    for (int b = 0; b < STFEM_MAX_BLOCKS; ++b) bp.p[b] = b < nb ? blocks[b] : nullptr;
    unsigned faces = 0;
    for (int d = 0; d < dim; ++d)
      for (int s = 0; s < 2; ++s)
        if (part.neighbor[d][s] >= 0) faces |= 1u << (2 * d + s);
    const int np2 = dim == 3 ? np[2] : 1;
    for (int d = 0; d < dim; ++d)
      {
        const unsigned sides = (faces >> (2 * d)) & 3u;
        if (!sides) continue;
        long long face = 1;
        for (int e = 0; e < dim; ++e)
          if (e != d) face *= np[e];
        const long long total = face * nb * 2;
        const int       grid  = (int)std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 8);
        k_scale_interface_faces<T><<<grid, 256, 0, ctx->stream>>>(bp, nb, np[0], np[1], np2, d, sides);
        ctx->launches++;
      }
    STFEM_CUDA_CHECK(cudaGetLastError());
    return STFEM_OK;
  }
  template int halo_scale_interfaces<double>(stfem_ctx *, const Partition &, HaloBuffers &, void *const *, int, const int[3], int);
  template int halo_scale_interfaces<float>(stfem_ctx *, const Partition &, HaloBuffers &, void *const *, int, const int[3], int);

  template int halo_compress_add<double>(stfem_ctx *, const Partition &, HaloBuffers &, void *const *, int, const int[3], int, cudaStream_t);
  template int halo_compress_add<float>(stfem_ctx *, const Partition &, HaloBuffers &, void *const *, int, const int[3], int, cudaStream_t);
} // namespace stfem

using namespace stfem;

extern "C" {

int stfem_comm_unique_id(char *id128)
{
  STFEM_REQUIRE(id128, "null id buffer");
  NcclApi *api = nccl_api();
  if (!api) return STFEM_ERR_UNSUPPORTED;
  NcclUniqueId id;
  STFEM_NCCL_CHECK(api->GetUniqueId(&id));
  std::memcpy(id128, id.internal, 128);
  return STFEM_OK;
}

int stfem_ctx_comm_init(stfem_ctx_t ctx, int rank, int n_ranks, const char *id128)
{
  STFEM_REQUIRE(ctx && id128, "stfem_ctx_comm_init: null argument");
  STFEM_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "stfem_ctx_comm_init: bad rank %d of %d", rank, n_ranks);
  NcclApi *api = nccl_api();
  if (!api) return STFEM_ERR_UNSUPPORTED;
  STFEM_CUDA_CHECK(cudaSetDevice(ctx->device));
  NcclUniqueId id;
  std::memcpy(id.internal, id128, 128);
  nccl_comm_t comm = nullptr;
  STFEM_NCCL_CHECK(api->CommInitRank(&comm, n_ranks, id, rank));
  ctx->nccl_comm = comm;
  ctx->rank      = rank;
  ctx->n_ranks   = n_ranks;
  return STFEM_OK;
}

int stfem_ctx_comm_destroy(stfem_ctx_t ctx)
{
  if (ctx && ctx->nccl_comm)
    {
      NcclApi *api = nccl_api();
      if (api) api->CommDestroy((nccl_comm_t)ctx->nccl_comm);
      ctx->nccl_comm = nullptr;
    }
  return STFEM_OK;
}

// host values reduced over the ranks of the context's communicator (in place); op 0 = sum, 1 = max, 2 = min.
// A context without a communicator (single GPU) returns the values unchanged.
int stfem_ctx_allreduce(stfem_ctx_t ctx, double *values, int n, int op)
{
  STFEM_REQUIRE(ctx && values && n >= 0 && n <= 256 && op >= 0 && op <= 2, "stfem_ctx_allreduce: bad arguments");
  if (!ctx->nccl_comm || ctx->n_ranks <= 1 || n == 0) return STFEM_OK;
  NcclApi *api = nccl_api();
  STFEM_REQUIRE(api, "stfem_ctx_allreduce: NCCL unavailable");
  STFEM_CUDA_CHECK(cudaSetDevice(ctx->device));
  double *d = nullptr;
  STFEM_CUDA_CHECK(cudaMalloc(&d, sizeof(double) * n));
  STFEM_CUDA_CHECK(cudaMemcpyAsync(d, values, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  static const int red[3] = {0, 2, 3}; // ncclSum, ncclMax, ncclMin
  int rc = api->AllReduce(d, d, (size_t)n, NcclApi::kDouble, red[op], (nccl_comm_t)ctx->nccl_comm, ctx->stream);
  if (rc == 0)
    {
      STFEM_CUDA_CHECK(cudaMemcpyAsync(values, d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
      STFEM_FORWARD(stream_sync_checked(ctx, "stfem_ctx_allreduce"));
    }
  cudaFree(d);
  if (rc != 0)
    {
      set_error("stfem_ctx_allreduce: ncclAllReduce failed: %s", api->GetErrorString(rc));
      return STFEM_ERR_CUDA;
    }
  return STFEM_OK;
}

int stfem_ctx_rank(stfem_ctx_t ctx) { return ctx ? ctx->rank : 0; }
int stfem_ctx_n_ranks(stfem_ctx_t ctx) { return ctx ? ctx->n_ranks : 1; }

// Box partition helper (host logic, also used by the CPU tests): brick of process `coords` in a
// proc_grid of a global mesh of n_global cells; writes local cell counts, the global cell offset, the local
// bounding box and the Dirichlet mask (physical boundary faces only).
int stfem_partition_brick(int dim, const int *n_global, const double *lower, const double *upper, const int *proc_grid,
                          const int *coords, int *n_local, int *cell_offset, double *local_lower, double *local_upper,
                          unsigned *dirichlet_faces)
{
  STFEM_REQUIRE(n_global && proc_grid && coords && n_local && cell_offset, "stfem_partition_brick: null argument");
  unsigned mask = 0;
  for (int d = 0; d < dim; ++d)
    {
      STFEM_REQUIRE(proc_grid[d] >= 1 && coords[d] >= 0 && coords[d] < proc_grid[d], "stfem_partition_brick: bad process grid");
      STFEM_REQUIRE(n_global[d] % proc_grid[d] == 0, "stfem_partition_brick: %d cells not divisible by %d processes in direction %d",
                    n_global[d], proc_grid[d], d);
      n_local[d]     = n_global[d] / proc_grid[d];
      cell_offset[d] = coords[d] * n_local[d];
      const double lo = lower ? lower[d] : 0.0, up = upper ? upper[d] : 1.0, h = (up - lo) / n_global[d];
      if (local_lower) local_lower[d] = lo + h * cell_offset[d];
      if (local_upper) local_upper[d] = lo + h * (cell_offset[d] + n_local[d]);
      if (coords[d] == 0) mask |= 1u << (2 * d);
      if (coords[d] == proc_grid[d] - 1) mask |= 1u << (2 * d + 1);
    }
  if (dirichlet_faces) *dirichlet_faces = mask;
  return STFEM_OK;
}

int stfem_mesh_set_partition(stfem_mesh_t mesh, const int *proc_grid, const int *coords)
{
  STFEM_REQUIRE(mesh && proc_grid && coords, "stfem_mesh_set_partition: null argument");
  Partition &p = mesh->part;
  int        total = 1;
  for (int d = 0; d < 3; ++d)
    {
      p.grid[d]   = d < mesh->dim ? proc_grid[d] : 1;
      p.coords[d] = d < mesh->dim ? coords[d] : 0;
      total *= p.grid[d];
    }
  STFEM_REQUIRE(total == mesh->ctx->n_ranks, "stfem_mesh_set_partition: process grid has %d entries, communicator %d ranks", total,
                mesh->ctx->n_ranks);
  STFEM_REQUIRE(p.rank_of(p.coords) == mesh->ctx->rank, "stfem_mesh_set_partition: coords do not match the rank (x fastest)");
  for (int d = 0; d < 3; ++d)
    for (int s = 0; s < 2; ++s)
      {
        int c[3] = {p.coords[0], p.coords[1], p.coords[2]};
        c[d] += s == 0 ? -1 : 1;
        p.neighbor[d][s] = (c[d] < 0 || c[d] >= p.grid[d]) ? -1 : p.rank_of(c);
      }
  p.active = total > 1;
  return STFEM_OK;
}

// Ghost layer of a partitioned mesh: one layer of cells across every face shared with another rank (deal.II's ghost cells).
// Only the dense cell-patch smoother needs it: its patch matrices take the contributions of the neighbouring cells on the
// shared DoFs (compute_block_matrix.h:50-139), and the neighbours of a cell at a rank interface live on the other rank.
namespace
{
  void ghost_extent(const stfem_mesh *m, int *ne, long long *n_cells, long long *n_vertices)
  {
    *n_cells = *n_vertices = 1;
    for (int d = 0; d < 3; ++d)
      {
        ne[d] = m->n[d];
        if (d < m->dim)
          {
            ne[d] += (m->part.neighbor[d][0] >= 0 ? 1 : 0) + (m->part.neighbor[d][1] >= 0 ? 1 : 0);
            *n_cells *= ne[d];
            *n_vertices *= ne[d] + 1;
          }
      }
  }
} // namespace

int stfem_mesh_set_ghost_vertices(stfem_mesh_t mesh, const double *vertices_ext)
{
  STFEM_REQUIRE(mesh && vertices_ext, "stfem_mesh_set_ghost_vertices: null argument");
  STFEM_REQUIRE(mesh->part.active, "stfem_mesh_set_ghost_vertices: the mesh has no partition (stfem_mesh_set_partition first)");
  int       ne[3];
  long long nc, nv;
  ghost_extent(mesh, ne, &nc, &nv);
  mesh->h_vertices_ghost.assign(vertices_ext, vertices_ext + nv * mesh->dim);
  return STFEM_OK;
}

int stfem_op_set_ghost_coefficients(stfem_op_t op, const double *coeff_cell_ext, const double *coeff_q_ext)
{
  STFEM_REQUIRE(op, "stfem_op_set_ghost_coefficients: null operator");
  stfem_mesh *mesh = op->mesh;
  STFEM_REQUIRE(mesh->part.active, "stfem_op_set_ghost_coefficients: the mesh has no partition");
  int       ne[3];
  long long nc, nv;
  ghost_extent(mesh, ne, &nc, &nv);
  const int n1 = op->degree + 1, nq = mesh->dim == 3 ? n1 * n1 * n1 : n1 * n1;
  if (coeff_cell_ext) op->h_coeff_cell_ghost.assign(coeff_cell_ext, coeff_cell_ext + nc);
  if (coeff_q_ext) op->h_coeff_q_ghost.assign(coeff_q_ext, coeff_q_ext + nc * nq);
  return STFEM_OK;
}

// TEST HOOK (host only, no GPU, no NCCL): the single-round interface exchange among the n_ranks = prod(proc_grid) bricks of a
// box partition, run in this process with the very element functions the pack / unpack kernels execute
// (halo_pack_element / halo_unpack_element, csrc/dist.cuh) and memcpy in place of ncclSend / ncclRecv.
// data: [n_ranks][nb][np0*np1*np2] doubles, the partial sums of every rank's brick; summed in place.
int stfem_halo_emulate_host(int dim, const int *proc_grid, const int *np, int nb, double *data)
{
  STFEM_REQUIRE(dim >= 1 && dim <= 3 && proc_grid && np && data && nb >= 1 && nb <= STFEM_MAX_BLOCKS, "stfem_halo_emulate_host: bad arguments");
  int       n_ranks = 1;
  long long N = 1;
  for (int d = 0; d < dim; ++d)
    {
      n_ranks *= proc_grid[d];
      N *= np[d];
    }
  std::vector<HaloPlan>            plans(n_ranks);
  std::vector<std::vector<double>> send(n_ranks), recv(n_ranks);
  std::vector<BlockPtrs>           bps(n_ranks);
  for (int r = 0; r < n_ranks; ++r)
    {
      Partition part;
      int       rr = r;
      for (int d = 0; d < 3; ++d)
        {
          part.grid[d]   = d < dim ? proc_grid[d] : 1;
          part.coords[d] = rr % part.grid[d];
          rr /= part.grid[d];
        }
      for (int d = 0; d < 3; ++d)
        for (int sd = 0; sd < 2; ++sd)
          {
            int c[3] = {part.coords[0], part.coords[1], part.coords[2]};
            c[d] += sd == 0 ? -1 : 1;
            part.neighbor[d][sd] = (c[d] < 0 || c[d] >= part.grid[d]) ? -1 : part.rank_of(c);
          }
      part.active = n_ranks > 1;
      halo_build_plan(plans[r], part, nb, np, dim, r);
      for (int b = 0; b < STFEM_MAX_BLOCKS; ++b) bps[r].p[b] = b < nb ? data + ((size_t)r * nb + b) * N : nullptr;
      const long long total = plans[r].start[plans[r].n_seg];
      send[r].assign(total, 0.0);
      recv[r].assign(total, 0.0);
      for (long long g = 0; g < total; ++g) halo_pack_element<double>(bps[r], plans[r], g, send[r].data());
    }
  // "send / receive": segment s of rank r goes to the segment of the opposite offset on rank plans[r].rank[s]
  for (int r = 0; r < n_ranks; ++r)
    for (int sgm = 0; sgm < plans[r].n_seg; ++sgm)
      {
        const int      peer = plans[r].rank[sgm];
        const int     *o    = plans[r].off[sgm];
        const int      t    = plans[peer].seg_of[(-o[0] + 1) + 3 * (-o[1] + 1) + 9 * (-o[2] + 1)];
        STFEM_REQUIRE(t >= 0 && plans[peer].rank[t] == r, "stfem_halo_emulate_host: neighbour relation is not symmetric");
        const long long cnt = plans[r].start[sgm + 1] - plans[r].start[sgm];
        STFEM_REQUIRE(cnt == plans[peer].start[t + 1] - plans[peer].start[t], "stfem_halo_emulate_host: segment sizes differ");
        std::memcpy(recv[peer].data() + plans[peer].start[t], send[r].data() + plans[r].start[sgm], sizeof(double) * cnt);
      }
  for (int r = 0; r < n_ranks; ++r)
    {
      const long long total = plans[r].start[plans[r].n_seg];
      for (long long g = 0; g < total; ++g) halo_unpack_element<double>(bps[r], plans[r], g, recv[r].data());
    }
  return STFEM_OK;
}

// sum of interface partial values (for tests and callers that fill vectors cell-wise themselves)
int stfem_op_halo_add(stfem_op_t op, void *const *blocks, int nb)
{
  STFEM_REQUIRE(op && blocks, "stfem_op_halo_add: null argument");
  if (op->number_type == STFEM_F64)
    return halo_compress_add<double>(op->mesh->ctx, op->mesh->part, op->halo, blocks, nb, op->np, op->mesh->dim);
  return halo_compress_add<float>(op->mesh->ctx, op->mesh->part, op->halo, blocks, nb, op->np, op->mesh->dim);
}

} // extern "C"
