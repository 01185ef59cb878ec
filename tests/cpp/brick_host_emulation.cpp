// Host emulation of st_vmult_brick_kernel (dealii-stfem_b200/csrc/st_vmult_brick.cuh): the kernel source is compiled as
// plain C++20 with shims for the CUDA execution model - one std::thread per CUDA thread of a CTA, std::barrier for
// __syncthreads(), a per-warp rendezvous for __shfl_sync, a global array for the dynamic shared memory, CTAs run one after
// another - so that the tile / chunk decomposition, the row-class addressing of the box loads (the plain-load path reads
// exactly the boxes the tensor maps describe), the lane mappings, the Dirichlet masks and the owner-writes stores can be
// checked against the oracle without a GPU (tests/test_brick_emulation.py).  Not a performance tool, not part of the product.
//
//   brick_host_emulation <degree> <nb> <nx> <ny> <nz> <hx> <hy> <hz> <dirichlet mask> <zlo> <zhi> <mode> <first_plane_acc>
//                        <n_chunks> <cx> <cy> <misalign> <in.bin> <out.bin> [f32]
// in.bin : Alpha[nb*nb], Beta[nb*nb], src[nb][N], dst0[nb][N]  (doubles; dst0 = initial content of dst);  out.bin: dst[nb][N]
// misalign: number of T elements the block vectors are shifted off a 16-byte boundary (tests the shifted row classes)
#include <algorithm>
#include <atomic>
#include <new>
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

struct Dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local Dim3 threadIdx, blockIdx;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __grid_constant__
#define __launch_bounds__(...)
#define __restrict__
#define __shared__
#define __align__(n) __attribute__((aligned(n)))
#define STFEM_HOST_EMULATION
#define STFEM_BRICK_STANDALONE

static std::barrier<> *g_barrier = nullptr;
inline void            __syncthreads() { g_barrier->arrive_and_wait(); }
struct WarpExchange
{
  std::barrier<> bar{32};
  double         buf[32];
};
static std::vector<std::unique_ptr<WarpExchange>> g_warps;
template <class T> inline T __shfl_sync(unsigned, T v, int src)
{
  WarpExchange &w    = *g_warps[threadIdx.x / 32];
  const int     lane = threadIdx.x % 32;
  w.buf[lane]        = (double)v;
  w.bar.arrive_and_wait();
  const T r = (T)w.buf[src & 31];
  w.bar.arrive_and_wait();
  return r;
}
static std::mutex g_atomic_mutex;
template <class T> inline void atomicAdd(T *p, T v)
{
  std::lock_guard<std::mutex> lock(g_atomic_mutex);
  *p += v;
}
using std::max;
using std::min;

namespace stfem
{
  __attribute__((aligned(128))) unsigned char brick_smem[1 << 18]; // the kernel's `extern __shared__` buffer
  // mbarrier emulation: the 8 bytes of a barrier hold {pending arrivals, phase}; `count` sits in a side table
  namespace brick_emu
  {
    struct Bar
    {
      std::atomic<int>      pending;
      std::atomic<unsigned> phase;
    };
    static_assert(sizeof(Bar) == 8, "emulated barrier must fit the 8 bytes of an mbarrier");
    static int g_count[64];
    inline int slot(const unsigned long long *bar) { return (int)(bar - reinterpret_cast<const unsigned long long *>(brick_smem)); }
    inline void mbar_init(unsigned long long *bar, unsigned count)
    {
      Bar *b = reinterpret_cast<Bar *>(bar);
      new (b) Bar();
      b->pending.store((int)count);
      b->phase.store(0u);
      g_count[slot(bar)] = (int)count;
    }
    inline void fence_init() {}
    inline void mbar_arrive(unsigned long long *bar)
    {
      Bar *b = reinterpret_cast<Bar *>(bar);
      if (b->pending.fetch_sub(1) == 1)
        {
          b->pending.store(g_count[slot(bar)]);
          b->phase.fetch_add(1u);
        }
    }
    inline void mbar_wait(unsigned long long *bar, unsigned parity)
    {
      Bar *b = reinterpret_cast<Bar *>(bar);
      while ((b->phase.load() & 1u) == parity) std::this_thread::yield();
    }
  } // namespace brick_emu
}

#include "../../dealii-stfem_b200/csrc/basis_host.hpp"
#include "../../dealii-stfem_b200/csrc/st_vmult_brick.cuh"

using namespace stfem;

struct Cmd
{
  int      n[3];
  double   h[3];
  unsigned mask;
  int      zlo, zhi, mode, first_plane_acc, n_chunks, misalign, flow_a;
};

template <int N1, int NB, typename T, int CX, int CY, bool SPLIT>
static int run(const Cmd &cmd, const std::vector<double> &in, std::vector<double> &out)
{
  using C = BrickCfg<T, N1, NB, CX, CY, SPLIT>;
  const int       degree = N1 - 1;
  const ShapeHost sh(degree);
  long long       N = 1;
  for (int d = 0; d < 3; ++d) N *= degree * cmd.n[d] + 1;
  const size_t need = 2 * NB * NB + 2 * (size_t)NB * N;
  if (in.size() != need)
    {
      std::fprintf(stderr, "input has %zu doubles, expected %zu\n", in.size(), need);
      return 2;
    }
  // block vectors like BlockVec: ONE array of NB*N numbers, here shifted off the 16-byte boundary by `misalign` elements
  std::vector<T> sbuf((size_t)NB * N + 64), dbuf((size_t)NB * N + 64);
  auto aligned = [&](std::vector<T> &v) {
    unsigned long long p = (unsigned long long)v.data();
    p                    = (p + 15ull) & ~15ull;
    return (T *)p + cmd.misalign;
  };
  T *src = aligned(sbuf), *dst = aligned(dbuf);
  for (size_t i = 0; i < (size_t)NB * N; ++i)
    {
      src[i] = (T)in[2 * NB * NB + i];
      dst[i] = (T)in[2 * NB * NB + (size_t)NB * N + i];
    }
  const void *sp[NB];
  void       *dp[NB];
  for (int b = 0; b < NB; ++b)
    {
      sp[b] = src + (size_t)b * N;
      dp[b] = dst + (size_t)b * N;
    }
  BrickArgs<T, N1, NB> a;
  std::memset(&a, 0, sizeof(a));
  // mode 2 (dst = rhs + A src): rhs = the initial content of dst, dst itself starts as garbage
  std::vector<T> rbuf;
  const void    *rp[NB];
  if (cmd.mode == 2)
    {
      rbuf.assign(dst, dst + (size_t)NB * N);
      for (size_t i = 0; i < (size_t)NB * N; ++i) dst[i] = (T)123.5;
    }
  for (int b = 0; b < NB; ++b) rp[b] = cmd.mode == 2 ? rbuf.data() + (size_t)b * N : nullptr;
  brick_fill_args<T, N1, NB, CX, CY>(a, sh.S.data(), sh.D.data(), sh.wq.data(), cmd.h, cmd.n, cmd.mask, in.data(), in.data() + NB * NB, sp, dp,
                                     cmd.zlo, cmd.zhi, cmd.mode, rp, cmd.first_plane_acc != 0, cmd.n_chunks, 296);
  a.use_tma = cmd.flow_a;
  if ((size_t)C::smem_bytes(a.n_cls) > sizeof(brick_smem)) return 3;
  const long long grid = (long long)a.tiles_x * a.tiles_y * a.n_chunks;
  std::printf("N1 %d NB %d tile %dx%d cells%s, %d threads, tiles %d x %d, %d chunks of %d layers, %d row classes (shifts", N1, NB, CX, CY, SPLIT ? " (split warps)" : "", C::NTHREADS,
              a.tiles_x, a.tiles_y, a.n_chunks, a.layers_per_chunk, a.n_cls);
  for (int b = 0; b < NB; ++b)
    for (int c = 0; c < a.n_cls; ++c) std::printf(" %d", a.shift[b][c]);
  std::printf("), smem %d bytes, grid %lld, flow %s\n", C::smem_bytes(a.n_cls), grid, a.use_tma == 2 ? "A2 (barrier pipeline)" : (a.use_tma ? "A1 (one CTA barrier per plane)" : "B (plain loads)"));
  // what launch_brick does before the kernel: the node planes shared by two z chunks start from zero in mode 0
  if (a.n_chunks > 1 && cmd.mode != 1)
    {
      const long long plane = (long long)a.np[0] * a.np[1];
      for (int b = 0; b < NB; ++b)
        for (int k = 1; k < a.n_chunks; ++k)
          for (long long e = 0; e < plane; ++e) ((T *)dp[b])[(long long)degree * (cmd.zlo + k * a.layers_per_chunk) * plane + e] = T(0);
    }
  g_warps.clear();
  for (int w = 0; w < C::NWARPS; ++w) g_warps.push_back(std::make_unique<WarpExchange>());
  for (long long blk = 0; blk < grid; ++blk)
    {
      std::memset(brick_smem, 0xff, sizeof(brick_smem)); // NaN pattern: reading something nobody wrote shows up in the result
      std::barrier<>           barrier(C::NTHREADS);
      std::vector<std::thread> pool;
      g_barrier = &barrier;
      for (int t = 0; t < C::NTHREADS; ++t)
        pool.emplace_back([&a, t, blk]() {
          threadIdx.x = (unsigned)t;
          blockIdx.x  = (unsigned)blk;
          st_vmult_brick_kernel<T, N1, NB, CX, CY, 1, SPLIT>(a);
        });
      for (auto &th : pool) th.join();
    }
  out.resize((size_t)NB * N);
  for (size_t i = 0; i < (size_t)NB * N; ++i) out[i] = (double)dst[i];
  return 0;
}

int main(int argc, char **argv)
{
  if (argc != 20 && argc != 21) return 1;
  const bool f32 = argc == 21 && std::string(argv[20]) == "f32";
  const int  degree = std::atoi(argv[1]), nb = std::atoi(argv[2]);
  Cmd        cmd;
  for (int d = 0; d < 3; ++d)
    {
      cmd.n[d] = std::atoi(argv[3 + d]);
      cmd.h[d] = std::atof(argv[6 + d]);
    }
  cmd.mask            = (unsigned)std::strtoul(argv[9], nullptr, 0);
  cmd.zlo             = std::atoi(argv[10]);
  cmd.zhi             = std::atoi(argv[11]);
  cmd.mode            = std::atoi(argv[12]);
  cmd.first_plane_acc = std::atoi(argv[13]);
  cmd.n_chunks        = std::atoi(argv[14]);
  const int cx = std::atoi(argv[15]), cy = std::atoi(argv[16]);
  cmd.misalign = std::atoi(argv[17]);
  // 1 (default): box loads by one thread + one CTA barrier per plane; 2: full / empty barrier pipeline; 0: plain loads by all threads
  cmd.flow_a   = std::getenv("BRICK_EMU_FLOW_B") ? 0 : (std::getenv("BRICK_EMU_PIPE") ? 2 : 1);
  std::vector<double> in, out;
  {
    FILE *f = std::fopen(argv[18], "rb");
    if (!f) return 1;
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    in.resize(sz / sizeof(double));
    if (std::fread(in.data(), sizeof(double), in.size(), f) != in.size()) return 1;
    std::fclose(f);
  }
  int        rc    = 4;
  const bool split = std::getenv("BRICK_EMU_SPLIT") != nullptr; // X and Y+Z phases on separate warps
  // the product's tiles (BrickTile<N1>) and small ones that give several tiles and ragged edges on tiny meshes
#define CASE(K_, NB_, CX_, CY_)                                                                              \
  if (degree == K_ && nb == NB_ && cx == CX_ && cy == CY_)                                                   \
    rc = split ? (f32 ? run<K_ + 1, NB_, float, CX_, CY_, true>(cmd, in, out) : run<K_ + 1, NB_, double, CX_, CY_, true>(cmd, in, out)) : \
                 (f32 ? run<K_ + 1, NB_, float, CX_, CY_, false>(cmd, in, out) : run<K_ + 1, NB_, double, CX_, CY_, false>(cmd, in, out));
  CASE(4, 2, 7, 4) CASE(4, 2, 3, 2) CASE(4, 1, 3, 2) CASE(4, 3, 3, 2) CASE(4, 1, 7, 4) CASE(4, 3, 7, 4)
  CASE(3, 2, 9, 5) CASE(3, 2, 4, 2) CASE(3, 3, 4, 2) CASE(3, 1, 4, 2)
  CASE(2, 2, 15, 6) CASE(2, 2, 3, 2) CASE(2, 3, 3, 2) CASE(2, 1, 3, 2)
#undef CASE
  if (rc != 0) return rc;
  FILE *f = std::fopen(argv[19], "wb");
  if (!f) return 1;
  std::fwrite(out.data(), sizeof(double), out.size(), f);
  std::fclose(f);
  return 0;
}
