// Host emulation of k_st_vmult_cart_fd (dealii-stfem_b200/csrc/st_vmult_cart_fd.cuh, kernel_variant 60): the kernel
// source is compiled as plain C++20 with shims for the CUDA execution model — one std::thread per CUDA thread of a CTA,
// std::barrier for __syncthreads(), a global buffer for the dynamic shared memory, CTAs run one after another — so that
// its thread mapping, shared-memory layout and Dirichlet masking can be checked against the oracle without a GPU
// (tests/test_cart_fd_modes.py).  Not a performance tool and not part of the product.
//
//   cart_fd_host_emulation <degree> <nb> <nx> <ny> <nz> <hx> <hy> <hz> <dirichlet mask> <in.bin> <out.bin> [f32]
// in.bin : Alpha[nb*nb], Beta[nb*nb], coeff[n_cells], src[nb][N]   (doubles);   out.bin: dst[nb][N]
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct Dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local Dim3 threadIdx, blockIdx;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __grid_constant__
#define __launch_bounds__(a, b)
#define __restrict__
#define __shared__
#define __align__(n) __attribute__((aligned(n)))
#define STFEM_MAX_BLOCKS 16
static std::barrier<> *g_barrier = nullptr;
static std::mutex      g_mutex;
inline void            __syncthreads() { g_barrier->arrive_and_wait(); }
template <class T> inline T    __ldg(const T *p) { return *p; }
template <class T> inline void atomicAdd(T *p, T v)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  *p += v;
}

namespace stfem
{
  __attribute__((aligned(16))) unsigned char smem_raw[1 << 18]; // the kernel's `extern __shared__` buffer
  // same constants as ExchLayout in st_vmult_cart.cuh (checked by the Python side against that file)
  template <int N1> struct ExchLayout;
  template <> struct ExchLayout<3> { static constexpr int LS = 3, CBS = 27; static constexpr bool blocked = true; };
  template <> struct ExchLayout<4> { static constexpr int LS = 5, CBS = 84; static constexpr bool blocked = false; };
  template <> struct ExchLayout<5> { static constexpr int LS = 5, CBS = 125; static constexpr bool blocked = true; };
} // namespace stfem

#define STFEM_CART_FD_STANDALONE
#include "../../dealii-stfem_b200/csrc/basis_host.hpp"
#include "../../dealii-stfem_b200/csrc/st_vmult_cart_fd.cuh"

using namespace stfem;

template <int N1, int NB, typename T>
static int run(int n[3], double h[3], unsigned mask, const std::vector<double> &in, std::vector<double> &out)
{
  const int       degree = N1 - 1;
  const ShapeHost sh(degree);
  std::vector<double> Mh(N1 * N1, 0.0), Kh(N1 * N1, 0.0), V(N1 * N1), lam(N1);
  for (int i = 0; i < N1; ++i)
    for (int j = 0; j < N1; ++j)
      for (int q = 0; q < N1; ++q)
        {
          Mh[i * N1 + j] += sh.wq[q] * sh.S[q * N1 + i] * sh.S[q * N1 + j];
          Kh[i * N1 + j] += sh.wq[q] * sh.D[q * N1 + i] * sh.D[q * N1 + j];
        }
  cartfd_host::pencil_modes(Mh.data(), Kh.data(), N1, V.data(), lam.data());
  // the argument block exactly as launch_cart_fd (csrc/capi_op.cu) fills it
  CartFdArgs<T, N1> a;
  const double           vol = h[0] * h[1] * h[2];
  long long              N = 1;
  for (int d = 0; d < 3; ++d)
    {
      a.n[d]  = n[d];
      a.np[d] = degree * n[d] + 1;
      N *= a.np[d];
    }
  for (int q = 0; q < N1; ++q)
    for (int i = 0; i < N1; ++i)
      {
        a.V[q * N1 + i]   = (T)V[q * N1 + i];
        a.Vt[i * N1 + q]  = (T)V[q * N1 + i];
        a.Vtx[i * N1 + q] = (T)(V[q * N1 + i] * vol);
      }
  for (int d = 0; d < 3; ++d)
    for (int q = 0; q < N1; ++q) a.lam[d][q] = (T)(lam[q] / (h[d] * h[d]));
  a.n_cells   = (long long)n[0] * n[1] * n[2];
  a.dirichlet = mask;
  const size_t need = 2 * NB * NB + a.n_cells + (size_t)NB * N;
  if (in.size() != need)
    {
      std::fprintf(stderr, "input has %zu doubles, expected %zu\n", in.size(), need);
      return 2;
    }
  // the operator's number type: everything the kernel touches is converted like the library does at op_create / upload
  std::vector<T> tin(in.begin(), in.end()), tout((size_t)NB * N, T(0));
  const T       *alpha = tin.data(), *beta = alpha + NB * NB, *coeff = beta + NB * NB, *src = coeff + a.n_cells;
  for (int b = 0; b < STFEM_MAX_BLOCKS; ++b)
    {
      a.src[b] = b < NB ? src + (size_t)b * N : nullptr;
      a.dst[b] = b < NB ? tout.data() + (size_t)b * N : nullptr;
    }
  a.alpha      = alpha;
  a.beta       = beta;
  a.coeff_cell = coeff;
  const int tpc = NB * N1;
  int       best = 1;
  double    best_score = -1;
  for (int c = 1; c * tpc <= 128; ++c)
    {
      const int    thr = c * tpc;
      const double eff = (double)thr / (((thr + 31) / 32) * 32);
      const double score = eff >= 0.9 ? 2.0 - 1e-4 * thr : eff;
      if (score > best_score + 1e-9) { best_score = score; best = c; }
    }
  a.cells_per_cta     = best;
  const int       threads = best * tpc;
  const long long grid    = (a.n_cells + best - 1) / best;
  if ((size_t)best * NB * ExchLayout<N1>::CBS * sizeof(T) > sizeof(smem_raw)) return 3;
  std::printf("N1 %d NB %d cells %lld: %d cells per CTA, %d threads, grid %lld\n", N1, NB, a.n_cells, best, threads, grid);
  for (long long blk = 0; blk < grid; ++blk)
    {
      std::barrier<>           barrier(threads);
      std::vector<std::thread> pool;
      g_barrier = &barrier;
      for (int t = 0; t < threads; ++t)
        pool.emplace_back([&a, t, blk]() {
          threadIdx.x = (unsigned)t;
          blockIdx.x  = (unsigned)blk;
          k_st_vmult_cart_fd<N1, NB, T>(a);
        });
      for (auto &th : pool) th.join();
    }
  out.assign(tout.begin(), tout.end());
  return 0;
}

int main(int argc, char **argv)
{
  if (argc != 12 && argc != 13) return 1;
  const bool f32 = argc == 13 && std::string(argv[12]) == "f32";
  const int degree = std::atoi(argv[1]), nb = std::atoi(argv[2]);
  int       n[3] = {std::atoi(argv[3]), std::atoi(argv[4]), std::atoi(argv[5])};
  double    h[3] = {std::atof(argv[6]), std::atof(argv[7]), std::atof(argv[8])};
  const unsigned mask = (unsigned)std::strtoul(argv[9], nullptr, 0);
  std::vector<double> in, out;
  {
    FILE *f = std::fopen(argv[10], "rb");
    if (!f) return 1;
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    in.resize(sz / sizeof(double));
    if (std::fread(in.data(), sizeof(double), in.size(), f) != in.size()) return 1;
    std::fclose(f);
  }
  int rc = 4;
#define CASE(K_, NB_) \
  if (degree == K_ && nb == NB_) rc = f32 ? run<K_ + 1, NB_, float>(n, h, mask, in, out) : run<K_ + 1, NB_, double>(n, h, mask, in, out);
  CASE(2, 2) CASE(2, 3) CASE(3, 2) CASE(3, 3) CASE(4, 2) CASE(4, 3)
#undef CASE
  if (rc != 0) return rc;
  FILE *f = std::fopen(argv[11], "wb");
  if (!f) return 1;
  std::fwrite(out.data(), sizeof(double), out.size(), f);
  std::fclose(f);
  return 0;
}
