// C ABI: context, device memory helpers, mesh.
#include <chrono>
#include <cstdlib>
#include <thread>

#include "common.hpp"

namespace stfem
{
  static thread_local char g_err[1024] = "";
  void set_error(const char *fmt, ...)
  {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
  }
  const char *get_error() { return g_err; }

  // Fail-fast synchronisation.  On a multi-GPU context a stream can wait for ever on a collective / send-receive pair the
  // peer ranks never enter (mismatched call sequences): instead of spinning until some watchdog kills the job minutes later,
  // the wait is bounded (STFEM_SYNC_TIMEOUT_S, default 60 s) and the call returns STFEM_ERR_CUDA with a message naming it.
  int stream_sync_checked(stfem_ctx *ctx, const char *what, cudaStream_t stream)
  {
    if (!stream) stream = ctx->stream;
    if (ctx->n_ranks <= 1 || !ctx->nccl_comm)
      {
        STFEM_CUDA_CHECK(cudaStreamSynchronize(stream));
        return STFEM_OK;
      }
    static const double limit = [] {
      const char *e = std::getenv("STFEM_SYNC_TIMEOUT_S");
      const double v = e ? std::atof(e) : 60.0;
      return v > 0 ? v : 60.0;
    }();
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spin = 0;; ++spin)
      {
        const cudaError_t q = cudaStreamQuery(stream);
        if (q == cudaSuccess) return STFEM_OK;
        if (q != cudaErrorNotReady)
          {
            set_error("%s: stream failed: %s", what, cudaGetErrorString(q));
            return STFEM_ERR_CUDA;
          }
        if ((spin & 1023u) == 1023u)
          {
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (dt > limit)
              {
                set_error("%s: rank %d of %d waited %.0f s for its stream - a collective or halo exchange the other ranks did not enter "
                          "(mismatched call sequence across ranks?)", what, ctx->rank, ctx->n_ranks, dt);
                return STFEM_ERR_CUDA;
              }
            if (dt > 0.002) std::this_thread::yield();
          }
      }
  }
} // namespace stfem

extern "C" {

const char *stfem_last_error(void) { return stfem::get_error(); }
const char *stfem_version(void) { return "stfem_b200 0.1 (sm_100a)"; }

int stfem_ctx_create(int device, stfem_ctx_t *out)
{
  STFEM_REQUIRE(out != nullptr, "stfem_ctx_create: out is null");
  int count = 0;
  STFEM_CUDA_CHECK(cudaGetDeviceCount(&count));
  STFEM_REQUIRE(device >= 0 && device < count, "stfem_ctx_create: device %d not available (%d devices)",
                device, count);
  STFEM_CUDA_CHECK(cudaSetDevice(device));
  auto *c   = new stfem_ctx();
  c->device = device;
  STFEM_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  STFEM_CUDA_CHECK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  STFEM_CUDA_CHECK(cudaEventCreate(&c->ev0));
  STFEM_CUDA_CHECK(cudaEventCreate(&c->ev1));
  STFEM_CUDA_CHECK(cudaEventCreate(&c->tm0));
  STFEM_CUDA_CHECK(cudaEventCreate(&c->tm1));
  *out = c;
  return STFEM_OK;
}

int stfem_ctx_destroy(stfem_ctx_t ctx)
{
  if (!ctx) return STFEM_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaEventDestroy(ctx->tm0);
  cudaEventDestroy(ctx->tm1);
  for (int i = 0; i < 2; ++i)
    if (ctx->aux[i]) cudaStreamDestroy(ctx->aux[i]);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  if (ctx->d_sm_counter) cudaFree(ctx->d_sm_counter);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return STFEM_OK;
}

int stfem_ctx_synchronize(stfem_ctx_t ctx)
{
  STFEM_REQUIRE(ctx, "null context");
  return stfem::stream_sync_checked(ctx, "stfem_ctx_synchronize", nullptr);
}

int stfem_ctx_timer_start(stfem_ctx_t ctx)
{
  STFEM_REQUIRE(ctx, "null context");
  STFEM_CUDA_CHECK(cudaEventRecord(ctx->tm0, ctx->stream));
  return STFEM_OK;
}

int stfem_ctx_timer_stop(stfem_ctx_t ctx, float *ms)
{
  STFEM_REQUIRE(ctx && ms, "null argument");
  STFEM_CUDA_CHECK(cudaEventRecord(ctx->tm1, ctx->stream));
  STFEM_FORWARD(stfem::stream_sync_checked(ctx, "stfem_ctx_timer_stop", nullptr));
  STFEM_CUDA_CHECK(cudaEventSynchronize(ctx->tm1));
  STFEM_CUDA_CHECK(cudaEventElapsedTime(ms, ctx->tm0, ctx->tm1));
  return STFEM_OK;
}

void *stfem_ctx_stream(stfem_ctx_t ctx) { return ctx ? (void *)ctx->stream : nullptr; }
long long stfem_ctx_launch_count(stfem_ctx_t ctx) { return ctx ? ctx->launches : 0; }

int stfem_dev_alloc(stfem_ctx_t ctx, size_t bytes, void **out)
{
  STFEM_REQUIRE(ctx && out, "stfem_dev_alloc: null argument");
  STFEM_CUDA_CHECK(cudaSetDevice(ctx->device));
  STFEM_CUDA_CHECK(cudaMalloc(out, bytes + 32)); // slack: kernels may over-read (never write) up to 16 bytes past the end
  return STFEM_OK;
}

int stfem_dev_free(stfem_ctx_t ctx, void *p)
{
  STFEM_REQUIRE(ctx, "null context");
  if (p)
    {
      STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      STFEM_CUDA_CHECK(cudaFree(p));
    }
  return STFEM_OK;
}

int stfem_dev_upload(stfem_ctx_t ctx, void *dst_dev, const void *src_host, size_t bytes)
{
  STFEM_REQUIRE(ctx, "null context");
  STFEM_CUDA_CHECK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return STFEM_OK;
}

int stfem_dev_download(stfem_ctx_t ctx, void *dst_host, const void *src_dev, size_t bytes)
{
  STFEM_REQUIRE(ctx, "null context");
  STFEM_CUDA_CHECK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return STFEM_OK;
}

int stfem_dev_memset(stfem_ctx_t ctx, void *dst_dev, int value, size_t bytes)
{
  STFEM_REQUIRE(ctx, "null context");
  STFEM_CUDA_CHECK(cudaMemsetAsync(dst_dev, value, bytes, ctx->stream));
  return STFEM_OK;
}

int stfem_dev_copy(stfem_ctx_t ctx, void *dst_dev, const void *src_dev, size_t bytes)
{
  STFEM_REQUIRE(ctx, "null context");
  STFEM_CUDA_CHECK(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return STFEM_OK;
}

int stfem_host_alloc_pinned(size_t bytes, void **out)
{
  STFEM_REQUIRE(out, "null out");
  STFEM_CUDA_CHECK(cudaMallocHost(out, bytes ? bytes : 1));
  return STFEM_OK;
}

int stfem_host_free_pinned(void *p)
{
  if (p) STFEM_CUDA_CHECK(cudaFreeHost(p));
  return STFEM_OK;
}

int stfem_mesh_create(stfem_ctx_t ctx, int dim, const int *n_cells, const double *lower,
                      const double *upper, const double *vertices, unsigned dirichlet_faces,
                      stfem_mesh_t *out)
{
  STFEM_REQUIRE(ctx && n_cells && out, "stfem_mesh_create: null argument");
  STFEM_REQUIRE(dim == 2 || dim == 3, "stfem_mesh_create: dim must be 2 or 3 (got %d)", dim);
  auto *m = new stfem_mesh();
  m->ctx  = ctx;
  m->dim  = dim;
  m->n_cells = 1;
  size_t nv  = 1;
  for (int d = 0; d < dim; ++d)
    {
      if (n_cells[d] < 1)
        {
          delete m;
          stfem::set_error("stfem_mesh_create: n_cells[%d] = %d", d, n_cells[d]);
          return STFEM_ERR_INVALID;
        }
      m->n[d]     = n_cells[d];
      m->lower[d] = lower ? lower[d] : 0.0;
      m->upper[d] = upper ? upper[d] : 1.0;
      m->n_cells *= n_cells[d];
      nv *= (size_t)n_cells[d] + 1;
    }
  m->dirichlet = dirichlet_faces;
  m->cartesian = vertices == nullptr;
  if (vertices)
    {
      m->h_vertices.assign(vertices, vertices + nv * dim);
      STFEM_CUDA_CHECK(cudaSetDevice(ctx->device));
      STFEM_CUDA_CHECK(cudaMalloc(&m->d_vertices, nv * dim * sizeof(double)));
      STFEM_CUDA_CHECK(cudaMemcpyAsync(m->d_vertices, vertices, nv * dim * sizeof(double),
                                       cudaMemcpyHostToDevice, ctx->stream));
      STFEM_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
  *out = m;
  return STFEM_OK;
}

int stfem_mesh_destroy(stfem_mesh_t mesh)
{
  if (!mesh) return STFEM_OK;
  if (mesh->d_vertices) cudaFree(mesh->d_vertices);
  delete mesh;
  return STFEM_OK;
}

} // extern "C"
