// Shared host-side plumbing of the C ABI: error reporting, context, mesh, operator objects.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/stfem_b200.h"

namespace stfem
{
  void        set_error(const char *fmt, ...);
  const char *get_error();

#define STFEM_CUDA_CHECK(expr)                                                                   \
  do                                                                                             \
    {                                                                                            \
      cudaError_t err__ = (expr);                                                                \
      if (err__ != cudaSuccess)                                                                  \
        {                                                                                        \
          stfem::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                    \
                           cudaGetErrorString(err__));                                           \
          return STFEM_ERR_CUDA;                                                                 \
        }                                                                                        \
    }                                                                                            \
  while (0)

#define STFEM_REQUIRE(cond, ...)                                                                 \
  do                                                                                             \
    {                                                                                            \
      if (!(cond))                                                                               \
        {                                                                                        \
          stfem::set_error(__VA_ARGS__);                                                         \
          return STFEM_ERR_INVALID;                                                              \
        }                                                                                        \
    }                                                                                            \
  while (0)

#define STFEM_FORWARD(expr)                                                                      \
  do                                                                                             \
    {                                                                                            \
      int rc__ = (expr);                                                                         \
      if (rc__ != STFEM_OK) return rc__;                                                         \
    }                                                                                            \
  while (0)
} // namespace stfem

struct stfem_ctx
{
  int          device  = 0;
  cudaStream_t stream  = nullptr;
  long long    launches = 0;
  int          sm_count = 0;
  cudaEvent_t  ev0 = nullptr, ev1 = nullptr; // per-operator kernel timing
  cudaEvent_t  tm0 = nullptr, tm1 = nullptr; // stfem_ctx_timer_*
  // auxiliary streams + events (created on first use): halo exchange overlapped with interior cells, and the
  // upload / compute / download pipeline of the host-buffer entry point
  cudaStream_t aux[2] = {nullptr, nullptr};
  cudaEvent_t  ev_fork = nullptr, ev_join = nullptr;
  std::vector<cudaEvent_t> ev_pool;
  void        *nccl_comm = nullptr;          // ncclComm_t when the context is part of a multi-GPU run
  int          rank = 0, n_ranks = 1;
  unsigned    *d_sm_counter = nullptr;       // one counter per SM: the brick kernel alternates its warp roles per SM with it
};

namespace stfem
{
  // cudaStreamSynchronize(stream ? stream : ctx->stream), bounded in time on multi-GPU contexts (capi_core.cu)
  int stream_sync_checked(stfem_ctx *ctx, const char *what, cudaStream_t stream = nullptr);
} // namespace stfem

namespace stfem
{
  struct PartitionInfo
  {
    bool active = false;
    int  grid[3] = {1, 1, 1}, coords[3] = {0, 0, 0};
    int  neighbor[3][2] = {{-1, -1}, {-1, -1}, {-1, -1}}; // rank of the low/high neighbour per direction, -1 = none
    int  rank_of(const int c[3]) const { return c[0] + grid[0] * (c[1] + grid[1] * c[2]); }
  };
} // namespace stfem

struct stfem_mesh
{
  stfem_ctx *ctx = nullptr;
  int        dim = 0;
  int        n[3] = {1, 1, 1};
  double     lower[3] = {0, 0, 0}, upper[3] = {1, 1, 1};
  long long  n_cells = 0;
  bool       cartesian = true;
  double    *d_vertices = nullptr; // device, (n+1)^dim * dim doubles, or null
  std::vector<double> h_vertices;
  unsigned   dirichlet = 0;
  stfem::PartitionInfo part; // box partition of a multi-GPU run (this mesh = the local brick)
  std::vector<double> h_vertices_ghost; // partitioned general meshes: vertices of the brick + one cell layer across shared faces
};
