// C ABI + implementation of the multi-GPU layer (see dist.cuh).
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

  template <typename T>
  int halo_compress_add(stfem_ctx *ctx, const Partition &part, HaloBuffers &hb, void *const *blocks, int nb, const int np[3], int dim,
                        cudaStream_t stream)
  {
    if (!part.active) return STFEM_OK;
    if (!stream) stream = ctx->stream;
    NcclApi *api = nccl_api();
    STFEM_REQUIRE(api && ctx->nccl_comm, "halo exchange: context has no NCCL communicator (stfem_ctx_comm_init)");
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
    const long long total = (long long)np[0] * np[1] * (dim == 3 ? np[2] : 1) * nb;
    const int       grid  = (int)std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 8);
    k_scale_interfaces<T><<<grid, 256, 0, ctx->stream>>>(bp, nb, np[0], np[1], dim == 3 ? np[2] : 1, faces);
    ctx->launches++;
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

// sum of interface partial values (for tests and callers that fill vectors cell-wise themselves)
int stfem_op_halo_add(stfem_op_t op, void *const *blocks, int nb)
{
  STFEM_REQUIRE(op && blocks, "stfem_op_halo_add: null argument");
  if (op->number_type == STFEM_F64)
    return halo_compress_add<double>(op->mesh->ctx, op->mesh->part, op->halo, blocks, nb, op->np, op->mesh->dim);
  return halo_compress_add<float>(op->mesh->ctx, op->mesh->part, op->halo, blocks, nb, op->np, op->mesh->dim);
}

} // extern "C"
