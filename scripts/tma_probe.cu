// Stand-alone probe of the tensor-map box loads the brick kernel relies on (csrc/st_vmult_brick.cuh): encodes the row-class
// descriptors with the library's own host code, loads one box per class with cp.async.bulk.tensor into shared memory and
// compares with the plain-load semantics (brick_box_load_plain).  Reports instead of trapping when a load never completes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I dealii-stfem_b200/csrc -I include -o tma_probe scripts/tma_probe.cu
#include <cstdio>
#include <cstring>
#include <vector>

#include "common.hpp"
#include "st_vmult_brick.cuh"

namespace stfem
{
  void        set_error(const char *, ...) {}
  const char *get_error() { return ""; }
} // namespace stfem
struct stfem_op;
#define STFEM_PROBE
typedef CUresult (*encode_fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

using namespace stfem;

struct ProbeArgs
{
  alignas(64) CUtensorMap map;
  BrickMapDesc desc;
  int          c0, c1, box_elems, esz;
  void        *out;
  int         *status;
};

template <typename T>
__global__ void probe_kernel(const __grid_constant__ ProbeArgs a)
{
  extern __shared__ __align__(128) unsigned char sm[];
  unsigned long long *bar  = reinterpret_cast<unsigned long long *>(sm);
  T                  *tile = reinterpret_cast<T *>(sm + 128);
  if (threadIdx.x == 0)
    {
      brick_hw::mbar_init(bar, 1);
      brick_hw::fence_init();
    }
  __syncthreads();
  if (threadIdx.x == 0)
    {
      brick_hw::mbar_expect_tx(bar, (unsigned)(a.box_elems * sizeof(T)));
      brick_hw::tma_load_2d(tile, &a.map, a.c0, a.c1, bar);
    }
  unsigned ok = 0;
  for (unsigned spin = 0; !ok && spin < (1u << 20); ++spin)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(brick_hw::smem_addr(bar)), "r"(0u)
                 : "memory");
  if (threadIdx.x == 0) *a.status = ok ? 1 : -1;
  if (ok)
    for (int e = threadIdx.x; e < a.box_elems; e += blockDim.x) reinterpret_cast<T *>(a.out)[e] = tile[e];
}

template <typename T>
int probe(encode_fn_t fn, int np0, int np1, int np2, int misalign, int box0, int box1, int x_first, long long row_first)
{
  const long long N = (long long)np0 * np1 * np2;
  std::vector<T>  h(N + 64);
  for (long long i = 0; i < N; ++i) h[i] = (T)(1 + (i % 100003) * 0.5);
  T *d_all = nullptr;
  cudaMalloc(&d_all, (N + 64) * sizeof(T));
  T *d = d_all + misalign;
  cudaMemcpy(d, h.data(), N * sizeof(T), cudaMemcpyHostToDevice);
  const long long pitch = (long long)np0 * sizeof(T);
  const int       n_cls = 16 / brick_gcd(16, pitch % 16 == 0 ? 16 : pitch % 16);
  BrickMapDesc    desc[4];
  int             shift[4];
  brick_describe_block<T>(d, np0, (long long)np1 * np2, n_cls, box0, box1, desc, shift);
  int   bad = 0;
  T    *d_out = nullptr;
  int  *d_status = nullptr;
  cudaMalloc(&d_out, (size_t)box0 * box1 * sizeof(T));
  cudaMalloc(&d_status, sizeof(int));
  for (int c = 0; c < n_cls; ++c)
    {
      ProbeArgs a;
      std::memset(&a, 0, sizeof(a));
      const cuuint64_t dims[2]    = {desc[c].dim0, desc[c].dim1};
      const cuuint64_t strides[1] = {desc[c].stride1};
      const cuuint32_t box[2]     = {(cuuint32_t)box0, (cuuint32_t)box1};
      const cuuint32_t estr[2]    = {1, 1};
      const CUresult   r = fn(&a.map, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(desc[c].base),
                              dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      std::printf("  class %d of %d: base %p shift %d dims %llu x %llu stride %llu box %d x %d -> encode rc %d\n", c, n_cls, desc[c].base, shift[c],
                  desc[c].dim0, desc[c].dim1, desc[c].stride1, box0, box1, (int)r);
      if (r != CUDA_SUCCESS)
        {
          ++bad;
          continue;
        }
      // first row >= row_first of this class
      const long long rc = row_first + (((long long)c - row_first) & (n_cls - 1));
      a.desc      = desc[c];
      constexpr int EPV = 16 / (int)sizeof(T);
      a.c0              = (x_first + shift[c]) & ~(EPV - 1); // boxes must start on a 16-byte boundary of global memory
      const int lead    = (x_first + shift[c]) - a.c0;
      a.c1        = (int)((rc - c) / n_cls);
      a.box_elems = box0 * box1;
      a.out       = d_out;
      a.status    = d_status;
      cudaMemset(d_status, 0, sizeof(int));
      cudaMemset(d_out, 0xff, (size_t)box0 * box1 * sizeof(T));
      const size_t smem = 128 + (size_t)box0 * box1 * sizeof(T) + 128;
      cudaFuncSetAttribute(probe_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      probe_kernel<T><<<1, 128, smem>>>(a);
      const cudaError_t e = cudaDeviceSynchronize();
      int status = 0;
      cudaMemcpy(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost);
      std::vector<T> got((size_t)box0 * box1);
      cudaMemcpy(got.data(), d_out, got.size() * sizeof(T), cudaMemcpyDeviceToHost);
      long long wrong = 0;
      for (int i1 = 0; i1 < box1; ++i1)
        for (int i0 = 0; i0 + lead < box0; ++i0)
          {
            const long long x = x_first + i0, row = rc + (long long)i1 * n_cls;
            // what the kernel expects: element x of tensor row `row`, 0 outside [0,np0) x [0, rows) -- except that a shifted
            // class exposes `shift` elements of the previous row at x < 0 (masked by the kernel, skipped here)
            if (x < 0) continue;
            T want = T(0);
            if (x < np0 && row >= 0 && row < (long long)np1 * np2) want = h[row * np0 + x];
            if (got[(size_t)i1 * box0 + lead + i0] != want) ++wrong;
          }
      std::printf("    load at (%d, %d): cuda '%s', barrier %s, %lld wrong elements\n", a.c0, a.c1, cudaGetErrorString(e),
                  status == 1 ? "completed" : (status == -1 ? "NEVER COMPLETED" : "not reached"), wrong);
      if (e != cudaSuccess || status != 1 || wrong) ++bad;
      if (e != cudaSuccess) return bad;
    }
  cudaFree(d_all);
  cudaFree(d_out);
  cudaFree(d_status);
  return bad;
}

int main()
{
  void                           *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  std::printf("cudaGetDriverEntryPoint: %s, query result %d, fn %p\n", cudaGetErrorString(e), (int)q, p);
  if (!p) return 1;
  encode_fn_t fn  = (encode_fn_t)p;
  int         bad = 0;
  std::printf("FP64, 33 x 21 x 13 nodes, box 34 x 11, interior\n");
  bad += probe<double>(fn, 33, 21, 13, 0, 34, 11, 0, 30);
  std::printf("FP64, misaligned base, box starting at x = -4, rows from -4\n");
  bad += probe<double>(fn, 33, 21, 13, 1, 34, 11, -4, -4);
  std::printf("FP64, 385^2 x 9 nodes, box past the end of the rows\n");
  bad += probe<double>(fn, 385, 385, 9, 0, 34, 11, 360, 385ll * 385 * 9 - 8);
  std::printf("FP32, 33 x 21 x 13 nodes (4 classes), box 36 x 6\n");
  bad += probe<float>(fn, 33, 21, 13, 0, 36, 6, 0, 30);
  bad += probe<float>(fn, 33, 21, 13, 3, 36, 6, -4, -4);
  std::printf("FP64, even pitch (1 class)\n");
  bad += probe<double>(fn, 34, 10, 4, 0, 34, 21, -4, 3);
  std::printf("%s\n", bad ? "PROBE FAILED" : "PROBE OK");
  std::fflush(stdout);
  return bad ? 2 : 0;
}
