// st_vmult, BRICK variant (3D, Cartesian mesh, constant coefficient): the fused space-time operator
//     dst_j (+)= sum_s  Alpha(j,s) K src_s + Beta(j,s) M src_s
// (reference: SystemMatrix::vmult over MatrixFreeOperator cell loops, include/operators.h:536-559, 1112-1173)
// organised so that every DoF is READ ONCE FROM HBM AND WRITTEN ONCE, without atomics and without zeroing dst:
//
//  * On a Cartesian mesh with constant coefficient the assembled operator factorises over the directions,
//        M = Mz (x) My (x) Mx,    K = Mz (x) My (x) Kx + Mz (x) Ky (x) Mx + Kz (x) My (x) Mx,
//    with the ASSEMBLED 1D matrices (block-banded: one (k+1)^2 block per cell, overlapping in the shared vertex).
//    Homogeneous Dirichlet constraints are separable too (read constrained nodes as 0, write 0 to constrained rows).
//  * One CTA owns a tile of CX x CY cells in x-y and marches through z plane by plane ("2.5D blocking"):
//      load   TMA box load (cp.async.bulk.tensor, mbarrier completion) of the (K(CX+1)+1) x (K(CY+1)+1) node tile of every
//             source block of plane z into a ring of shared-memory stages, two to three planes ahead;
//      X      per tile row and cell (one lane each): a_s = Mx u_s, b_s = Kx u_s by the even-odd decomposition of the
//             (k+1)^2 cell matrices, temporal contraction  P_j = sum_s Beta(j,s) a_s + Alpha(j,s) b_s,
//             Q_j = sum_s Alpha(j,s) a_s;  the partial sum of the vertex shared with the left cell comes by warp shuffle;
//      Y+Z    one lane per x node, one warp per (cell row, dst block):  c = My P + Ky Q,  d = My Q  on its K (+1) nodes,
//             then  out(z') += Mz(z',z) c + Kz(z',z) d  into register accumulators of the K+1 planes of the current cell
//             layer; when a layer is complete its K planes are written with plain coalesced stores.
//    The only redundancy is the halo cell on the low side of the tile in x and y (re-read through L2, its X-phase rows
//    recomputed).  The z range of a launch is cut into chunks (one CTA each) to fill the GPU evenly; the node plane between
//    two chunks receives one partial sum from either side by atomic adds (two addends: order-independent result).
//  * Row pitches of the block vectors ((k n + 1) numbers) are not multiples of 16 bytes; the tensor maps therefore view a
//    block as 16/gcd(16, pitch) interleaved row classes (2 for FP64, 4 for FP32 when the pitch is odd), each with a 16-byte
//    aligned base and a row stride that is a multiple of 16 bytes.  A box must also START on a 16-byte boundary of global
//    memory (measured: cp.async.bulk.tensor raises an illegal-instruction error otherwise), so every class' box starts at the
//    boundary at or below the first node of the tile row and the X phase skips the 0..3 lead-in elements.  Out-of-range box
//    elements are zero-filled by TMA.
#pragma once
#ifndef STFEM_HOST_EMULATION
#include <cuda.h>
#include <cuda_runtime.h>
#endif

#include <type_traits>

#include "../../include/stfem_b200.h"

namespace stfem
{
  // what the host needs to encode one tensor map (and what the host emulation of the kernel reads instead of it)
  struct BrickMapDesc
  {
    const void        *base;    // 16-byte aligned
    unsigned long long dim0, dim1;    // elements, rows
    unsigned long long stride1; // bytes
    int                box0, box1;
  };

  template <int N1> struct BrickTile;
  template <> struct BrickTile<3> { static constexpr int CX = 15, CY = 6; };
  template <> struct BrickTile<4> { static constexpr int CX = 9, CY = 5; };
  template <> struct BrickTile<5> { static constexpr int CX = 7, CY = 4; };

  // SPLIT: the X phase and the Y+Z phase run on separate warps (XW + YW warps, one CTA per SM) instead of sharing them
  template <typename T, int N1, int NB, int CX_, int CY_, bool SPLIT = false>
  struct BrickCfg
  {
    static constexpr int K = N1 - 1, CX = CX_, CY = CY_;
    static constexpr int TX = K * CX, TY = K * CY;           // owned nodes of a tile (+ the last plane of the mesh)
    static constexpr int WX = TX + K + 1, WY = TY + K + 1;   // input nodes: one halo cell on the low side
    static constexpr int EPV = 16 / (int)sizeof(T);
    // box width: a multiple of 16 bytes that also covers the lead-in elements of a box whose start has to be moved down to
    // a 16-byte boundary of global memory (TMA rejects - illegal instruction - boxes that start anywhere else)
    static constexpr int WXP = ((WX + EPV - 1 + EPV - 1) / EPV) * EPV;
    static constexpr int MAXCLS = EPV;                       // row classes: 1, 2 (,4 for FP32)
    static constexpr int CXL = CX + 1;                       // cells per tile row incl. the halo cell
    static constexpr int RPW = 32 / CXL;                     // tile rows per warp in the X phase
    static constexpr int XW = (WY + RPW - 1) / RPW;          // warps with X work
    static constexpr int YW = CY * NB;                       // warps of the Y+Z phase: one per (cell row, dst block)
    static constexpr int NWARPS = SPLIT ? XW + YW : (XW > YW ? XW : YW);
    static constexpr int NTHREADS = 32 * NWARPS;
    static constexpr int STAGES = 3;  // planes of source tiles in flight
    static constexpr int NPQ = 3;     // P/Q buffers: X runs one plane ahead of Y+Z with a plane of slack on either side
    // pitch of a P/Q row: an odd number of 16-byte words, so that the 16-byte stores of the X phase (one run of K values per
    // lane) spread over all banks; measured faster than scalar stores on an odd pitch in FP64 (fewer issue slots)
    static constexpr int PXP = ((TX + 1 + EPV - 1) / EPV) * EPV + ((((TX + 1 + EPV - 1) / EPV) & 1) ? 0 : EPV);
    static constexpr int PQ_FIELD = WY * PXP;                // one field of one dst block
    static constexpr int PQ_BUF = 2 * NB * PQ_FIELD;
    static_assert(CXL <= 32 && TX + 1 <= 32, "tile does not fit the lane mappings");
    static_assert(RPW >= 1, "tile row does not fit a warp");
    // bytes of one (source block, row class) sub-tile for n_cls classes, padded to 128 bytes
    __host__ __device__ static constexpr int sub_bytes(int n_cls) { return ((((WY + n_cls - 1) / n_cls) * WXP * (int)sizeof(T)) + 127) / 128 * 128; }
    __host__ __device__ static constexpr int stage_bytes(int n_cls) { return NB * n_cls * sub_bytes(n_cls); }
    __host__ __device__ static constexpr int smem_bytes(int n_cls) { return 256 + STAGES * stage_bytes(n_cls) + NPQ * PQ_BUF * (int)sizeof(T); }
  };

  template <typename T, int N1, int NB>
  struct BrickArgs
  {
    static constexpr int NE = N1 - N1 / 2, NO = N1 / 2;
#ifndef STFEM_HOST_EMULATION
    alignas(64) CUtensorMap maps[NB][4];
#endif
    BrickMapDesc desc[NB][4]; // the same in plain form (host emulation, non-TMA load path)
    int          shift[NB][4]; // elements between the aligned base of a class and its first row
    int          n_cls, cls_shift; // row classes, log2
    unsigned     shift_pack;       // shift[s][c] as 2 bits at position 2 * (4 s + c)
    int          zero_top_row;     // the top node row in y is a Dirichlet row no chunk owns: the last chunk writes its zeros
    // x: even/odd parts of vol*Mh and vol*Kh/hx^2; y, z: full matrices Mh, Kh/hy^2, Kh/hz^2 (row = output node)
    T Mxe[NE * NE], Mxo[NO * NO + 1], Kxe[NE * NE], Kxo[NO * NO + 1];
    T M[N1 * N1], Ky[N1 * N1], Kz[N1 * N1];
    T alpha[NB * NB], beta[NB * NB]; // row-major [dst j][src s]
    int n[3], np[3];
    int zlo, zhi;        // cell layers [zlo, zhi) of this launch
    int layers_per_chunk, n_chunks, tiles_x, tiles_y;
    int mode;            // 0: dst = A src, 1: dst += A src, 2: dst = rhs + A src
    int first_plane_acc; // plane K*zlo is accumulated even in mode 0 (z-slab pipeline: it holds the partial sum of the slab below)
    int use_tma;
    unsigned *sm_counter; // one counter per SM (see xrot in the kernel); null: no rotation of the X warps
    unsigned dirichlet;
    unsigned iface;      // bit 2d+s: face s of direction d is shared with another rank (mode 2: rhs enters with weight 1/multiplicity)
    const T *src[NB];
    T       *dst[NB];
    const T *rhs[NB]; // mode 2
  };


  // ------------------------------------------------------------------------------------------------ host helpers (plain C++)
  inline int brick_gcd(long long a, long long b) { return b == 0 ? (int)a : brick_gcd(b, a % b); }

  // the row classes of one block vector: class c = rows r = c (mod n_cls), each a 2D tensor with an aligned base
  template <typename T>
  inline void brick_describe_block(const void *base, int np0, long long n_rows, int n_cls, int box0, int box1, BrickMapDesc *desc, int *shift)
  {
    for (int c = 0; c < n_cls; ++c)
      {
        const unsigned long long first = (unsigned long long)base + (unsigned long long)c * np0 * sizeof(T);
        desc[c].base    = (const void *)(first & ~15ull);
        shift[c]        = (int)((first & 15ull) / sizeof(T));
        desc[c].dim0    = (unsigned long long)np0 + shift[c];
        desc[c].dim1    = (unsigned long long)((n_rows - c + n_cls - 1) / n_cls);
        desc[c].stride1 = (unsigned long long)n_cls * np0 * sizeof(T);
        desc[c].box0    = box0;
        desc[c].box1    = box1;
      }
  }

  // even / odd parts of a centrosymmetric n1 x n1 matrix (row-major, row = output): see brick_eo_apply2
  template <typename T>
  inline void brick_even_odd(const long double *A, int n1, T *Ae, T *Ao)
  {
    const int K = n1 - 1, NO = n1 / 2, NE = n1 - NO;
    for (int i = 0; i < NE; ++i)
      for (int k = 0; k < NE; ++k) Ae[i * NE + k] = (T)((k < NO) ? 0.5L * (A[i * n1 + k] + A[i * n1 + K - k]) : A[i * n1 + k]);
    for (int i = 0; i < NO; ++i)
      for (int k = 0; k < NO; ++k) Ao[i * NO + k] = (T)(0.5L * (A[i * n1 + k] - A[i * n1 + K - k]));
  }

  // number of z chunks: minimise  waves x (layers per chunk + start-up)  for `slots` resident CTAs
  inline int brick_choose_chunks(long long tiles, int layers, long long slots)
  {
    int    best = 1;
    double best_cost = 1e300;
    for (int nc = 1; nc <= layers && nc <= 64; ++nc)
      {
        const int       lpc   = (layers + nc - 1) / nc;
        const int       real  = (layers + lpc - 1) / lpc;
        const long long ctas  = tiles * real;
        const long long waves = (ctas + slots - 1) / slots;
        const double    cost  = (double)waves * (lpc + 0.75); // 0.75: pipeline fill / drain of a CTA
        if (cost < best_cost - 1e-9)
          {
            best_cost = cost;
            best      = real;
          }
      }
    return best;
  }

  // everything of the argument block except the tensor maps.  S, D, wq: shape values / derivatives [q][i] and weights of
  // QGauss(k+1) for the GLL-nodal FE_Q(k) basis on [0,1]; h: cell sizes; n_chunks <= 0: choose for `slots` resident CTAs
  template <typename T, int N1, int NB, int CX, int CY>
  inline void brick_fill_args(BrickArgs<T, N1, NB> &a, const double *S, const double *D, const double *wq, const double h[3], const int n[3],
                              unsigned dirichlet, const double *Alpha, const double *Beta, const void *const *src, void *const *dst, int zlo,
                              int zhi, int mode, const void *const *rhs, bool first_plane_acc, int n_chunks, long long slots, unsigned iface = 0)
  {
    using C = BrickCfg<T, N1, NB, CX, CY>;
    constexpr int K = N1 - 1;
    const double  vol = h[0] * h[1] * h[2];
    for (int d = 0; d < 3; ++d)
      {
        a.n[d]  = n[d];
        a.np[d] = K * n[d] + 1;
      }
    long double Mh[N1 * N1], Kh[N1 * N1], Mx[N1 * N1], Kx[N1 * N1];
    for (int i = 0; i < N1; ++i)
      for (int j = 0; j < N1; ++j)
        {
          long double mm = 0, kk = 0;
          for (int q = 0; q < N1; ++q)
            {
              mm += (long double)wq[q] * S[q * N1 + i] * S[q * N1 + j];
              kk += (long double)wq[q] * D[q * N1 + i] * D[q * N1 + j];
            }
          Mh[i * N1 + j] = mm;
          Kh[i * N1 + j] = kk;
        }
    // enforce the exact symmetries the even-odd form relies on (they hold up to rounding of the quadrature sums)
    for (int i = 0; i < N1; ++i)
      for (int j = 0; j < N1; ++j)
        {
          const long double ms = 0.25L * (Mh[i * N1 + j] + Mh[j * N1 + i] + Mh[(K - i) * N1 + K - j] + Mh[(K - j) * N1 + K - i]);
          const long double ks = 0.25L * (Kh[i * N1 + j] + Kh[j * N1 + i] + Kh[(K - i) * N1 + K - j] + Kh[(K - j) * N1 + K - i]);
          Mx[i * N1 + j]   = ms * vol;
          Kx[i * N1 + j]   = ks * vol / (h[0] * h[0]);
          a.M[i * N1 + j]  = (T)ms;
          a.Ky[i * N1 + j] = (T)(ks / (h[1] * h[1]));
          a.Kz[i * N1 + j] = (T)(ks / (h[2] * h[2]));
        }
    brick_even_odd<T>(Mx, N1, a.Mxe, a.Mxo);
    brick_even_odd<T>(Kx, N1, a.Kxe, a.Kxo);
    for (int i = 0; i < NB * NB; ++i)
      {
        a.alpha[i] = (T)Alpha[i];
        a.beta[i]  = (T)Beta[i];
      }
    a.zlo             = zlo;
    a.zhi             = zhi;
    a.mode            = mode;
    a.first_plane_acc = first_plane_acc ? 1 : 0;
    a.dirichlet       = dirichlet;
    a.iface           = mode == 2 ? iface : 0u;
    a.tiles_x         = (n[0] + CX - 1) / CX;
    // the top node row (y = K n1) is node 0 of the chunk of the non-existing cell n1: one more chunk, i.e. one more tile row
    // when n1 is a multiple of CY - unless that row is a Dirichlet row, whose zeros the last chunk writes
    a.zero_top_row    = (n[1] % CY == 0 && (dirichlet & 8u)) ? 1 : 0;
    a.tiles_y         = a.zero_top_row ? n[1] / CY : (n[1] + 1 + CY - 1) / CY;
    const int layers  = zhi - zlo;
    if (n_chunks <= 0) n_chunks = brick_choose_chunks((long long)a.tiles_x * a.tiles_y, layers, slots);
    if (n_chunks > layers) n_chunks = layers;
    a.layers_per_chunk = (layers + n_chunks - 1) / n_chunks;
    a.n_chunks         = (layers + a.layers_per_chunk - 1) / a.layers_per_chunk;
    const long long pitch = (long long)a.np[0] * sizeof(T);
    a.n_cls               = 16 / brick_gcd(16, pitch % 16 == 0 ? 16 : pitch % 16);
    a.cls_shift            = a.n_cls == 1 ? 0 : (a.n_cls == 2 ? 1 : 2);
    const int       HB     = (C::WY + a.n_cls - 1) / a.n_cls;
    const long long n_rows = (long long)a.np[1] * a.np[2];
    a.use_tma              = 0;
    a.shift_pack           = 0;
    for (int b = 0; b < NB; ++b)
      {
        a.src[b] = (const T *)src[b];
        a.dst[b] = (T *)dst[b];
        a.rhs[b] = mode == 2 ? (const T *)rhs[b] : nullptr;
        brick_describe_block<T>(src[b], a.np[0], n_rows, a.n_cls, C::WXP, HB, a.desc[b], a.shift[b]);
        for (int c = 0; c < a.n_cls; ++c) a.shift_pack |= (unsigned)a.shift[b][c] << (2 * (4 * b + c));
      }
  }

  // ------------------------------------------------------------------------------------------------ async-copy primitives
#ifndef STFEM_HOST_EMULATION
  namespace brick_hw
  {
    __device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
    __device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
    {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
    }
    __device__ __forceinline__ void fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
    {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    }
    // bounded wait: a lost transaction must end in a trap, not in a hung GPU.  Plain polling: measured faster than a
    // suspend-time hint (slow wake-up) and than __nanosleep back-off
    __device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
    {
      unsigned ok = 0;
      for (unsigned spin = 0; !ok; ++spin)
        {
          asm volatile("{\n"
                       ".reg .pred p;\n"
                       "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                       "selp.u32 %0, 1, 0, p;\n"
                       "}"
                       : "=r"(ok)
                       : "r"(smem_addr(bar)), "r"(parity)
                       : "memory");
          if (!ok && spin > (1u << 22)) __trap();
        }
    }
    __device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, unsigned long long *bar)
    {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_addr(dst)),
                   "l"((unsigned long long)map), "r"(c0), "r"(c1), "r"(smem_addr(bar))
                   : "memory");
    }
  } // namespace brick_hw
#endif

  // mbarrier wrappers of the kernel: hardware (brick_hw) or the host emulation's (brick_emu, tests/cpp)
#ifndef STFEM_HOST_EMULATION
  namespace brick_hw
  {
    __device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
    {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
    }
  } // namespace brick_hw
  namespace brick_sync = brick_hw;
#else
  namespace brick_sync = brick_emu;
#endif

  // plain (synchronous) version of one box load: element (i0, i1) of the box = tensor element (c0 + i0, c1 + i1), 0 outside
  template <typename T>
  __device__ __forceinline__ void brick_box_load_plain(T *dst, const BrickMapDesc &d, int c0, int c1, int tid, int nthreads)
  {
    const int total = d.box0 * d.box1;
    for (int e = tid; e < total; e += nthreads)
      {
        const int       i0 = e % d.box0, i1 = e / d.box0;
        const long long g0 = (long long)c0 + i0, g1 = (long long)c1 + i1;
        T               v  = T(0);
        if (g0 >= 0 && g0 < (long long)d.dim0 && g1 >= 0 && g1 < (long long)d.dim1)
          v = *reinterpret_cast<const T *>(reinterpret_cast<const char *>(d.base) + (size_t)g1 * d.stride1 + (size_t)g0 * sizeof(T));
        dst[e] = v;
      }
  }

  // y = A u for a centrosymmetric (k+1)^2 matrix pair given by their even / odd parts, both applied to the same u
  template <typename T, int N1>
  __device__ __forceinline__ void brick_eo_apply2(const T (&u)[N1], const T *__restrict__ Ae, const T *__restrict__ Ao, const T *__restrict__ Be,
                                                  const T *__restrict__ Bo, T (&ya)[N1], T (&yb)[N1])
  {
    constexpr int K = N1 - 1, NO = N1 / 2, NE = N1 - NO;
    T             e[NE], o[NO > 0 ? NO : 1];
#pragma unroll
    for (int k = 0; k < NO; ++k)
      {
        e[k] = u[k] + u[K - k];
        o[k] = u[k] - u[K - k];
      }
    if (N1 & 1) e[NE - 1] = u[NO];
#pragma unroll
    for (int i = 0; i < NE; ++i)
      {
        T ea = Ae[i * NE] * e[0], eb = Be[i * NE] * e[0];
#pragma unroll
        for (int k = 1; k < NE; ++k)
          {
            ea += Ae[i * NE + k] * e[k];
            eb += Be[i * NE + k] * e[k];
          }
        if (i < NO)
          {
            T oa = Ao[i * NO] * o[0], ob = Bo[i * NO] * o[0];
#pragma unroll
            for (int k = 1; k < NO; ++k)
              {
                oa += Ao[i * NO + k] * o[k];
                ob += Bo[i * NO + k] * o[k];
              }
            ya[i]     = ea + oa;
            ya[K - i] = ea - oa;
            yb[i]     = eb + ob;
            yb[K - i] = eb - ob;
          }
        else
          {
            ya[i] = ea;
            yb[i] = eb;
          }
      }
  }

  // lane <-> (tile row within the warp, cell) of the X phase.  For 8 cells per row the two low bits of the cell index are
  // swapped so that the 8 lanes of a quarter warp read (16-byte accesses) / write conflict-free.
  template <int CXL, int RPW>
  __device__ __forceinline__ int brick_cell_of(int cb)
  {
    if (CXL == 8) return (cb & 4) | ((cb & 1) << 1) | ((cb >> 1) & 1);
    return cb;
  }

  // K consecutive values to shared memory, as 16-byte stores where the type and K allow (p is 16-byte aligned then)
  template <typename T, int K>
  __device__ __forceinline__ void brick_store_run(T *p, const T *v)
  {
#ifndef STFEM_HOST_EMULATION
    if constexpr (sizeof(T) == 8 && K % 2 == 0)
      {
#pragma unroll
        for (int i = 0; i < K; i += 2) *reinterpret_cast<double2 *>(p + i) = make_double2((double)v[i], (double)v[i + 1]);
        return;
      }
    else if constexpr (sizeof(T) == 4 && K % 4 == 0)
      {
#pragma unroll
        for (int i = 0; i < K; i += 4) *reinterpret_cast<float4 *>(p + i) = make_float4((float)v[i], (float)v[i + 1], (float)v[i + 2], (float)v[i + 3]);
        return;
      }
#endif
#pragma unroll
    for (int i = 0; i < K; ++i) p[i] = v[i];
  }

  // mode 0 with several z chunks: the node planes shared by two chunks start from zero (both chunks add their partial sums)
  template <typename T, int NB>
  struct BrickZeroArgs
  {
    T        *dst[NB];
    long long plane;  // numbers per node plane
    int       first, stride, count; // planes first + k * stride, k < count
  };
#ifndef STFEM_HOST_EMULATION
  template <typename T, int NB>
  __global__ void brick_zero_planes_kernel(const BrickZeroArgs<T, NB> a)
  {
    const long long total = a.plane * a.count * NB;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
      {
        const long long in_plane = e % a.plane;
        const long long r        = e / a.plane;
        const int       k = (int)(r % a.count), b = (int)(r / a.count);
        a.dst[b][(long long)(a.first + k * a.stride) * a.plane + in_plane] = T(0);
      }
  }
#endif

  // GEN = false: the plain product only (mode 0, no accumulated first plane, no interface weights) - the store phase then
  // has nothing to read and no mode to test (the headline instance)
  template <typename T, int N1, int NB, int CX, int CY, int MINB, bool SPLIT = false, bool GEN = true>
  __global__ void __launch_bounds__((BrickCfg<T, N1, NB, CX, CY, SPLIT>::NTHREADS), MINB)
    st_vmult_brick_kernel(const __grid_constant__ BrickArgs<T, N1, NB> a)
  {
    using C = BrickCfg<T, N1, NB, CX, CY, SPLIT>;
    constexpr int K = C::K, TX = C::TX, WY = C::WY, WXP = C::WXP, PXP = C::PXP, S = C::STAGES, EPV = C::EPV;
    constexpr int CXL = C::CXL, RPW = C::RPW;
    extern __shared__ __align__(128) unsigned char brick_smem[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(brick_smem);
    const int           cshift = a.cls_shift, cmask = (1 << cshift) - 1; // row classes: 1 << cshift
    const int           HB = (WY + cmask) >> cshift;                    // box rows per class
    const int           sub_elems = C::sub_bytes(1 << cshift) / (int)sizeof(T);
    const int           stage_elems = (NB << cshift) * sub_elems;
    T                  *tiles = reinterpret_cast<T *>(brick_smem + 256);
    T                  *zero_row = reinterpret_cast<T *>(brick_smem + 128); // N1 <= 8 zeros: what an invalid lane of the X phase reads
    T                  *pq    = tiles + S * stage_elems;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // The X phase occupies XW of the NWARPS warps, i.e. (warps go round-robin to the four schedulers of an SM) it loads the
    // schedulers unevenly.  Two CTAs share an SM: every other CTA an SM receives (counted per SM) starts its X warps two
    // schedulers further on, so that the pair loads all four evenly.
    int xrot = 0;
#ifndef STFEM_HOST_EMULATION
    if (!SPLIT && a.sm_counter)
      {
        __shared__ int s_xrot;
        if (tid == 0)
          {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            s_xrot = (atomicAdd(a.sm_counter + smid, 1u) & 1u) ? 2 % C::NWARPS : 0;
          }
        __syncthreads();
        xrot = s_xrot;
      }
#endif
    const int xwarp = SPLIT ? warp : (warp - xrot + C::NWARPS) % C::NWARPS; // index of this warp among the X warps

    // ---- this CTA's tile and z chunk
    int       bid   = blockIdx.x;
    const int tix   = bid % a.tiles_x;
    bid /= a.tiles_x;
    const int tiy   = bid % a.tiles_y;
    const int chunk = bid / a.tiles_y;
    const int cx0 = tix * CX, cy0 = tiy * CY; // first owned cell
    const int x0 = cx0 * K, y0 = cy0 * K;     // first owned node
    const int cz0 = a.zlo + chunk * a.layers_per_chunk;
    const int cz1 = min(a.zhi, cz0 + a.layers_per_chunk);
    if (cz0 >= cz1) return;
    // The node plane between two z chunks of a launch gets a partial sum from either chunk: both add theirs atomically
    // (two addends: the result does not depend on their order); in mode 0 the launcher zeroes those planes first.
    const int start = cz0;
    const int nq    = K * (cz1 - start) + 1; // node planes this CTA walks through
    const int np0 = a.np[0], np1 = a.np[1], np2 = a.np[2];
    const unsigned dm = a.dirichlet;
    const int      xb = x0 - K;                    // first node of the tile rows (may be negative)

    // ---- barriers: [0,S) tile full (TMA transaction), [S,2S) tile empty (all X warps read it),
    //      [2S, 2S+NPQ) P/Q full (all X warps wrote the plane), [2S+NPQ, 2S+2NPQ) P/Q empty (all Y+Z warps read it)
    constexpr int NPQ = C::NPQ;
    unsigned long long *bar_tfull = bars, *bar_tempty = bars + S, *bar_pfull = bars + 2 * S, *bar_pempty = bars + 2 * S + NPQ;
    static_assert(2 * S + 2 * NPQ <= 16, "barriers do not fit the 128 bytes reserved for them");

    // ---- plane loads: plane index qq of this CTA into stage qq % S.  Tensor row of tile row 0: r0_first + qq * np1
    const int r0_first = (K * start) * np1 + (y0 - K);
    // flow A (TMA): one thread issues the box loads, completion on bar_tfull
    auto issue_tma = [&](int qq) {
#ifndef STFEM_HOST_EMULATION
      T        *st = tiles + (qq % S) * stage_elems;
      const int r0 = r0_first + qq * np1;
      brick_hw::mbar_expect_tx(&bar_tfull[qq % S], (unsigned)((NB << cshift) * HB * WXP * (int)sizeof(T)));
      for (int s = 0; s < NB; ++s)
        for (int c = 0; c <= cmask; ++c)
          {
            const int rc = r0 + ((c - r0) & cmask); // first row >= r0 of class c
            const int sh = (a.shift_pack >> (2 * (4 * s + c))) & 3;
            brick_hw::tma_load_2d(st + ((s << cshift) + c) * sub_elems, &a.maps[s][c], (xb + sh) & ~(EPV - 1), (rc - c) >> cshift, &bar_tfull[qq % S]);
          }
#else
      // host emulation of flow A: the issuing thread copies the boxes itself, then completes the barrier
      T        *st = tiles + (qq % S) * stage_elems;
      const int r0 = r0_first + qq * np1;
      for (int s = 0; s < NB; ++s)
        for (int c = 0; c <= cmask; ++c)
          {
            const int rc = r0 + ((c - r0) & cmask);
            const int sh = (a.shift_pack >> (2 * (4 * s + c))) & 3;
            brick_box_load_plain<T>(st + ((s << cshift) + c) * sub_elems, a.desc[s][c], (xb + sh) & ~(EPV - 1), (rc - c) >> cshift, 0, 1);
          }
      brick_sync::mbar_arrive(&bar_tfull[qq % S]);
#endif
    };
    // flow B (plain loads, no tensor maps): all threads copy the boxes, CTA-wide barriers around the phases
    auto load_plain = [&](int qq) {
      T        *st = tiles + (qq % S) * stage_elems;
      const int r0 = r0_first + qq * np1;
      for (int s = 0; s < NB; ++s)
        for (int c = 0; c <= cmask; ++c)
          {
            const int rc = r0 + ((c - r0) & cmask);
            const int sh = (a.shift_pack >> (2 * (4 * s + c))) & 3;
            brick_box_load_plain<T>(st + ((s << cshift) + c) * sub_elems, a.desc[s][c], (xb + sh) & ~(EPV - 1), (rc - c) >> cshift, tid, C::NTHREADS);
          }
    };

    const bool flow_a = a.use_tma != 0;      // tensor-map loads (1: one CTA barrier per plane, 2: full / empty barrier pipeline)
    const bool pipe   = a.use_tma == 2;
    if (tid < 8) zero_row[tid] = T(0);
    if (!flow_a) __syncthreads();
    if (flow_a)
      {
        if (tid == 0)
          {
            for (int s = 0; s < S; ++s)
              {
                brick_sync::mbar_init(&bar_tfull[s], 1);
                brick_sync::mbar_init(&bar_tempty[s], 32 * C::XW);
              }
            for (int b = 0; b < NPQ; ++b)
              {
                brick_sync::mbar_init(&bar_pfull[b], 32 * C::XW);
                brick_sync::mbar_init(&bar_pempty[b], 32 * C::YW);
              }
            brick_sync::fence_init();
          }
        __syncthreads();
        if (tid == C::NTHREADS - 32)
          for (int qq = 0; qq < S && qq < nq; ++qq) issue_tma(qq);
      }

    // ---- X phase identity of this thread: lane = (tile row within the warp, cell), all plane-independent pieces hoisted
    const bool x_lane_ok = lane < RPW * CXL;
    const int  xr  = x_lane_ok ? lane % RPW : 0;
    const int  xci = brick_cell_of<CXL, RPW>(x_lane_ok ? lane / RPW : 0); // 0 = halo cell on the low side
    const int  xyl = xwarp * RPW + xr;                                    // tile row
    const bool x_row_in = xwarp < C::XW && x_lane_ok && xyl < WY;
    bool       x_valid0;
    bool       x_zero_u0, x_zero_uK;
    {
      const int yg = y0 - K + xyl, cxg = cx0 - 1 + xci;
      x_valid0  = x_row_in && yg >= 0 && yg < np1 && !((dm & 4u) && yg == 0) && !((dm & 8u) && yg == np1 - 1) && cxg >= 0 && cxg < a.n[0];
      x_zero_u0 = (dm & 1u) && cxg == 0;
      x_zero_uK = (dm & 2u) && cxg == a.n[0] - 1;
    }
    int x_src_lane = lane; // lane holding the left neighbour cell of the same tile row
    if (x_lane_ok && xci >= 1)
      {
        int cbl = 0;
#pragma unroll
        for (int t = 0; t < CXL; ++t)
          if (brick_cell_of<CXL, RPW>(t) == xci - 1) cbl = t;
        x_src_lane = xr + RPW * cbl;
      }
    const bool x_store  = x_row_in && xci >= 1;
    unsigned   x_lead_pack = 0; // lead-in elements of the box of (source block, row class), 2 bits each
    for (int s = 0; s < NB; ++s)
      for (int c = 0; c <= cmask; ++c)
        x_lead_pack |= (unsigned)((xb + ((a.shift_pack >> (2 * (4 * s + c))) & 3)) & (EPV - 1)) << (2 * (4 * s + c));
    const int  x_st_off = xyl * PXP + K * (xci - 1); // P/Q store offset inside a field

    // ---- Y+Z phase identity: lane = x node of the tile, warp = (cell row yc, dst block j)
    const int  ywarp = SPLIT ? (warp >= C::XW ? warp - C::XW : 0) : warp; // index among the Y+Z warps
    const bool y_role = SPLIT ? warp >= C::XW : warp < C::YW;              // takes part in the Y+Z barrier protocol
    const int  yc = ywarp % CY, jz = (ywarp / CY) % NB;
    const int  xo = lane <= TX ? lane : TX;
    const int  xg = x0 + xo;
    const int  cyg = cy0 + yc;                      // cell whose K node rows K*cyg .. K*cyg + K - 1 this thread owns
    const bool yz_warp = y_role && cyg <= a.n[1];
    const bool has_below = cyg >= 1, has_above = cyg < a.n[1];
    const int  yz_ld_off = (K * yc) * PXP + xo;
    // stores: node nn of the chunk is written iff st_mask bit nn; its value is forced to 0 iff con_mask bit nn
    unsigned   st_mask = 0, con_mask = 0, wy_mask = 0;
    T          wx = T(1);
    {
      const bool x_ok  = lane <= TX && xg < np0 && (lane < TX || xg == np0 - 1);
      const bool x_con = ((dm & 1u) && xg == 0) || ((dm & 2u) && xg == np0 - 1);
#pragma unroll
      for (int nn = 0; nn < K; ++nn)
        {
          const int yg = K * cyg + nn;
          if (x_ok && yg < np1) st_mask |= 1u << nn;
          if (x_con || ((dm & 4u) && yg == 0) || ((dm & 8u) && yg == np1 - 1)) con_mask |= 1u << nn;
        }
      // mode 2 on a partitioned mesh: every rank sharing a node adds rhs / multiplicity, the exchange then sums to rhs + A src
      if (GEN && a.iface)
        {
          if (((a.iface & 1u) && xg == 0) || ((a.iface & 2u) && xg == np0 - 1)) wx = T(0.5);
#pragma unroll
          for (int nn = 0; nn <= K; ++nn)
            {
              const int yg = K * cyg + nn;
              if (((a.iface & 4u) && yg == 0) || ((a.iface & 8u) && yg == np1 - 1)) wy_mask |= 1u << nn;
            }
        }
      // the top node row of a mesh whose cell count is a multiple of CY is not owned by any chunk when it is a Dirichlet row
      // (the launcher then saves the extra tile row): the last chunk writes its zeros
      if (a.zero_top_row && x_ok && cyg == a.n[1] - 1) st_mask |= 1u << K;
    }
    T *const dst_base = a.dst[jz] + ((long long)xg + (long long)np0 * (K * cyg));
    const long long plane_stride = (long long)np0 * np1;
    T          acc[K][N1]; // [node of the chunk][plane of the current cell layer]
#pragma unroll
    for (int nn = 0; nn < K; ++nn)
#pragma unroll
      for (int i = 0; i < N1; ++i) acc[nn][i] = T(0);

    // Write NP consecutive node planes zb .. zb+NP-1 (local plane indices 0 .. NP-1 of the accumulators) of the chunk's nodes.
    //   mode 0  dst = A src        mode 1  dst += A src        mode 2  dst = rhs + A src   (constrained rows: 0 / dst / rhs)
    // The plane between two z chunks of the launch (first_shared / last plane with shared_top) gets both partial sums by
    // atomic adds onto zeros (modes 0, 2; in mode 2 the upper chunk adds rhs as well) or onto dst (mode 1).
    // All reads (old dst or rhs) are issued first, as one batch: a load-add-store chain per value would expose the memory
    // latency NP * K times per layer.
    auto store_planes = [&](int zb, auto nptag, bool first_shared) {
      constexpr int NP = decltype(nptag)::value;
      const int     mode = GEN ? a.mode : 0;
      T            *p0   = dst_base + plane_stride * zb;
      const T      *r0p  = (mode == 2 ? a.rhs[jz] : a.dst[jz]) + (dst_base - a.dst[jz]) + plane_stride * zb;
      const bool    acc0 = GEN && a.first_plane_acc && zb == K * a.zlo; // z-slab launches: plane K*zlo holds the partial sum of the slab below
      if (mode == 0 && !acc0)
        {
          // plain assignment: nothing to read
#pragma unroll
          for (int i = 0; i < NP; ++i)
            {
              const int  z     = zb + i;
              const bool z_con = ((dm & 16u) && z == 0) || ((dm & 32u) && z == np2 - 1);
              T         *p     = p0 + plane_stride * i;
              if (i == 0 && first_shared)
                {
                  if (!z_con)
                    {
#pragma unroll
                      for (int nn = 0; nn < K; ++nn)
                        if (((st_mask & ~con_mask) >> nn) & 1u) atomicAdd(p + nn * np0, acc[nn][i]);
                    }
                }
              else if (!z_con && con_mask == 0 && !((st_mask >> K) & 1u))
                {
#pragma unroll
                  for (int nn = 0; nn < K; ++nn)
                    if ((st_mask >> nn) & 1u) p[nn * np0] = acc[nn][i];
                }
              else
                {
#pragma unroll
                  for (int nn = 0; nn < K; ++nn)
                    if ((st_mask >> nn) & 1u) p[nn * np0] = (z_con || ((con_mask >> nn) & 1u)) ? T(0) : acc[nn][i];
                  if ((st_mask >> K) & 1u) p[K * np0] = T(0);
                }
            }
          return;
        }
      if (!GEN) return;
      T old[NP][K + 1];
      if (mode != 0 || acc0)
        {
#pragma unroll
          for (int i = 0; i < NP; ++i)
#pragma unroll
            for (int nn = 0; nn <= K; ++nn)
              {
                const bool need = ((st_mask >> nn) & 1u) && (mode != 0 || i == 0) && !(mode == 1 && i == 0 && first_shared);
                old[i][nn]      = need ? r0p[plane_stride * i + nn * np0] : T(0);
              }
        }
      else
        {
#pragma unroll
          for (int i = 0; i < NP; ++i)
#pragma unroll
            for (int nn = 0; nn <= K; ++nn) old[i][nn] = T(0);
        }
      if (mode == 2 && a.iface)
        {
#pragma unroll
          for (int i = 0; i < NP; ++i)
            {
              const int z  = zb + i;
              const T   wz = (((a.iface & 16u) && z == 0) || ((a.iface & 32u) && z == np2 - 1)) ? T(0.5) : T(1);
#pragma unroll
              for (int nn = 0; nn <= K; ++nn) old[i][nn] *= wx * wz * (((wy_mask >> nn) & 1u) ? T(0.5) : T(1));
            }
        }
#pragma unroll
      for (int i = 0; i < NP; ++i)
        {
          const int  z      = zb + i;
          const bool z_con  = ((dm & 16u) && z == 0) || ((dm & 32u) && z == np2 - 1);
          const bool shared = i == 0 && first_shared;
          T         *p      = p0 + plane_stride * i;
#pragma unroll
          for (int nn = 0; nn < K; ++nn)
            {
              if (!((st_mask >> nn) & 1u)) continue;
              const bool con = z_con || ((con_mask >> nn) & 1u);
              const T    v   = con ? T(0) : acc[nn][i];
              if (shared)
                {
                  if (mode == 2)
                    atomicAdd(p + nn * np0, v + old[i][nn]); // this chunk is the upper one of the plane: it brings rhs along
                  else if (!con)
                    atomicAdd(p + nn * np0, v);
                }
              else if (mode == 1 || (acc0 && i == 0))
                {
                  if (!con) p[nn * np0] = old[i][nn] + v;
                }
              else
                p[nn * np0] = old[i][nn] + v; // mode 0: old = 0
            }
          // the Dirichlet row on top of the mesh that no chunk owns (see zero_top_row): 0 / untouched / rhs
          if (((st_mask >> K) & 1u) && mode != 1 && !(acc0 && i == 0) && !shared) p[K * np0] = old[i][K];
          if (((st_mask >> K) & 1u) && mode == 2 && shared) atomicAdd(p + K * np0, old[i][K]);
        }
    };
    // the top plane of the chunk: if a chunk follows, it is shared with it and this chunk is the LOWER one
    auto store_top = [&](int z, bool shared_top) {
      const int  mode  = a.mode;
      const bool z_con = ((dm & 16u) && z == 0) || ((dm & 32u) && z == np2 - 1);
      T         *p     = dst_base + plane_stride * z;
      if (shared_top)
        {
          if (!z_con)
            {
#pragma unroll
              for (int nn = 0; nn < K; ++nn)
                if (((st_mask & ~con_mask) >> nn) & 1u) atomicAdd(p + nn * np0, acc[nn][0]);
            }
          return;
        }
      store_planes(z, std::integral_constant<int, 1>(), false);
      (void)mode;
    };

    // ---- the two phases of one node plane q (local index of this CTA's march)
    // X: source tile of stage q % S -> P, Q of buffer q % NPQ (X warps only)
    auto x_phase = [&](int q) {
      const int  zp = K * start + q;
      const bool plane_con = ((dm & 16u) && zp == 0) || ((dm & 32u) && zp == np2 - 1);
      const int  r0 = r0_first + q * np1;
      T         *st  = tiles + (q % S) * stage_elems;
      T         *pqb = pq + (q % NPQ) * C::PQ_BUF;
        {
          const bool valid = x_valid0 && !plane_con;
          // tile row xyl sits in class c at index idx of that class' box
          const int g   = r0 + xyl;
          const int c   = g & cmask;
          const int idx = (g - (r0 + ((c - r0) & cmask))) >> cshift;
          T         P[NB][N1], Q[NB][N1];
#pragma unroll
          for (int s = 0; s < NB; ++s)
            {
              // the box of this class starts at the 16-byte boundary at or below the first node of the tile row
              const int lead = (x_lead_pack >> (2 * (4 * s + c))) & 3;
              // an invalid lane (row outside the mesh, constrained row or plane, no such cell) reads zeros: FP32 through a
              // pointer to a zero row, FP64 by selects (measured: the dependent pointer select costs more there)
              const T *row = st + ((s << cshift) + c) * sub_elems + idx * WXP + lead + K * xci;
              if (sizeof(T) == 4 && !valid) row = zero_row;
              T u[N1], av[N1], bv[N1];
#pragma unroll
              for (int i = 0; i < N1; ++i) u[i] = (sizeof(T) == 4 || valid) ? row[i] : T(0);
              if (x_zero_u0) u[0] = T(0);
              if (x_zero_uK) u[K] = T(0);
              brick_eo_apply2<T, N1>(u, a.Mxe, a.Mxo, a.Kxe, a.Kxo, av, bv);
#pragma unroll
              for (int j = 0; j < NB; ++j)
                {
                  const T be = a.beta[j * NB + s], al = a.alpha[j * NB + s];
#pragma unroll
                  for (int i = 0; i < N1; ++i)
                    {
                      if (s == 0)
                        {
                          P[j][i] = be * av[i];
                          Q[j][i] = al * av[i];
                        }
                      else
                        {
                          P[j][i] += be * av[i];
                          Q[j][i] += al * av[i];
                        }
                      P[j][i] += al * bv[i];
                    }
                }
            }
          // the vertex shared with the left cell: add that cell's partial sum (all lanes take part in the shuffles)
#pragma unroll
          for (int j = 0; j < NB; ++j)
            {
              const T pk = __shfl_sync(0xffffffffu, P[j][K], x_src_lane);
              const T qk = __shfl_sync(0xffffffffu, Q[j][K], x_src_lane);
              if (x_src_lane != lane)
                {
                  P[j][0] += pk;
                  Q[j][0] += qk;
                }
            }
          if (x_store)
            {
#pragma unroll
              for (int j = 0; j < NB; ++j)
                {
                  T *pp = pqb + (0 * NB + j) * C::PQ_FIELD + x_st_off;
                  T *qp = pqb + (1 * NB + j) * C::PQ_FIELD + x_st_off;
                  brick_store_run<T, K>(pp, P[j]);
                  brick_store_run<T, K>(qp, Q[j]);
                  if (xci == CX)
                    {
                      pp[K] = P[j][K];
                      qp[K] = Q[j][K];
                    }
                }
            }
        }
    };
    // Y + Z: P, Q of buffer q % NPQ -> accumulators, stores of completed planes (Y+Z warps with an existing chunk only;
    // `release` is called once the P, Q values are in registers)
    auto yz_phase = [&](int q, auto release) {
      const int m     = q % K; // local index of the plane in the current cell layer (0: also node K of the layer below)
      const int layer = start + q / K;
      T        *pqb   = pq + (q % NPQ) * C::PQ_BUF;
        {
          const T *pp = pqb + (0 * NB + jz) * C::PQ_FIELD + yz_ld_off;
          const T *qp = pqb + (1 * NB + jz) * C::PQ_FIELD + yz_ld_off;
          T        p[2 * K + 1], qv[2 * K + 1];
#pragma unroll
          for (int t = 0; t < 2 * K + 1; ++t)
            {
              p[t]  = pp[t * PXP];
              qv[t] = qp[t * PXP];
            }
          release(); // the buffer may be rewritten as soon as every Y+Z warp got here
          T c[K], d[K];
          // node 0 of the chunk: vertex row shared by the cell below (its node K) and the cell above (its node 0).  The cell
          // matrices are centrosymmetric (row K = row 0 reversed): with both cells present the two contributions are one row
          // applied to the sums of the mirrored values
          // (FP32 only: in FP64 the two extra independent chains of the general form hide the longer pipe latency better)
          if (sizeof(T) == 4 && has_below && has_above)
            {
              T sp = p[K] + p[K], sq = qv[K] + qv[K];
              T cc = a.M[0] * sp, dd = a.M[0] * sq;
              cc += a.Ky[0] * sq;
#pragma unroll
              for (int t = 1; t < N1; ++t)
                {
                  sp = p[K + t] + p[K - t];
                  sq = qv[K + t] + qv[K - t];
                  cc += a.M[t] * sp;
                  cc += a.Ky[t] * sq;
                  dd += a.M[t] * sq;
                }
              c[0] = cc;
              d[0] = dd;
            }
          else
            {
              // short independent chains (mass part, stiffness part) instead of one long one: the FP64 pipe latency is
              // what the Y phase waits for
              T cb_ = a.M[K * N1] * p[0], db_ = a.M[K * N1] * qv[0], ca_ = a.M[0] * p[K], da_ = a.M[0] * qv[K];
              T kb_ = a.Ky[K * N1] * qv[0], ka_ = a.Ky[0] * qv[K];
#pragma unroll
              for (int t = 1; t < N1; ++t)
                {
                  cb_ += a.M[K * N1 + t] * p[t];
                  kb_ += a.Ky[K * N1 + t] * qv[t];
                  db_ += a.M[K * N1 + t] * qv[t];
                  ca_ += a.M[t] * p[K + t];
                  ka_ += a.Ky[t] * qv[K + t];
                  da_ += a.M[t] * qv[K + t];
                }
              cb_ += kb_;
              ca_ += ka_;
              c[0] = (has_below ? cb_ : T(0)) + (has_above ? ca_ : T(0));
              d[0] = (has_below ? db_ : T(0)) + (has_above ? da_ : T(0));
            }
#pragma unroll
          for (int nn = 1; nn < K; ++nn)
            {
              T cc = a.M[nn * N1] * p[K], dd = a.M[nn * N1] * qv[K], ck = a.Ky[nn * N1] * qv[K];
#pragma unroll
              for (int t = 1; t < N1; ++t)
                {
                  cc += a.M[nn * N1 + t] * p[K + t];
                  ck += a.Ky[nn * N1 + t] * qv[K + t];
                  dd += a.M[nn * N1 + t] * qv[K + t];
                }
              c[nn] = cc + ck;
              d[nn] = dd;
            }
          auto zacc = [&](auto mtag) {
            constexpr int mm = decltype(mtag)::value;
#pragma unroll
            for (int nn = 0; nn < K; ++nn)
#pragma unroll
              for (int i = 0; i < N1; ++i)
                {
                  acc[nn][i] += a.M[i * N1 + mm] * c[nn];
                  acc[nn][i] += a.Kz[i * N1 + mm] * d[nn];
                }
          };
          if (m == 0)
            {
              if (q > 0)
                {
                  zacc(std::integral_constant<int, K>()); // closes cell layer `layer - 1`
                  {
                    store_planes(K * (layer - 1), std::integral_constant<int, K>(), layer - 1 == cz0 && cz0 > a.zlo);
                  }
#pragma unroll
                  for (int nn = 0; nn < K; ++nn)
                    {
                      acc[nn][0] = acc[nn][K];
#pragma unroll
                      for (int i = 1; i < N1; ++i) acc[nn][i] = T(0);
                    }
                }
              if (q < nq - 1) zacc(std::integral_constant<int, 0>());
            }
          else
            {
              switch (m)
                {
                  case 1: zacc(std::integral_constant<int, (K > 1 ? 1 : 0)>()); break;
                  case 2: zacc(std::integral_constant<int, (K > 2 ? 2 : 0)>()); break;
                  case 3: zacc(std::integral_constant<int, (K > 3 ? 3 : 0)>()); break;
                  case 4: zacc(std::integral_constant<int, (K > 4 ? 4 : 0)>()); break;
                  default: zacc(std::integral_constant<int, (K > 5 ? 5 : 0)>()); break;
                }
            }
        }
    };

    // ---- march through the planes
    if (flow_a && !pipe)
      {
        // one CTA barrier per plane: X(q) | barrier | reload of the stage, Y+Z(q); the P/Q buffers alternate, so the X phase
        // of the next plane never overwrites what a slower warp still reads
        for (int q = 0; q < nq; ++q)
          {
            brick_sync::mbar_wait(&bar_tfull[q % S], (unsigned)((q / S) & 1));
            if (xwarp < C::XW) x_phase(q);
            __syncthreads();
            if (tid == C::NTHREADS - 32 && q + S < nq) issue_tma(q + S);
            if (yz_warp) yz_phase(q, []() {});
          }
      }
    else if (flow_a)
      {
        // no CTA-wide barrier: X runs one plane ahead of Y+Z; the full / empty barriers of the three P/Q buffers and of
        // the source stages leave a plane of slack on either side, so a warp only waits when it is a whole plane ahead
        const bool x_warp = xwarp < C::XW, y_warp = y_role, issuer = tid == C::NTHREADS - 32;
        auto       x_step = [&](int q) {
          brick_sync::mbar_wait(&bar_tfull[q % S], (unsigned)((q / S) & 1));
          if (q >= NPQ) brick_sync::mbar_wait(&bar_pempty[q % NPQ], (unsigned)((q / NPQ - 1) & 1));
          x_phase(q);
          brick_sync::mbar_arrive(&bar_pfull[q % NPQ]);
          brick_sync::mbar_arrive(&bar_tempty[q % S]);
        };
        if (x_warp) x_step(0);
        for (int q = 0; q < nq; ++q)
          {
            if (x_warp && q + 1 < nq) x_step(q + 1);
            if (issuer && q + S < nq)
              {
                brick_sync::mbar_wait(&bar_tempty[q % S], (unsigned)((q / S) & 1)); // plane q read by every X warp
                issue_tma(q + S);
              }
            if (y_warp)
              {
                brick_sync::mbar_wait(&bar_pfull[q % NPQ], (unsigned)((q / NPQ) & 1));
                if (yz_warp)
                  yz_phase(q, [&]() { brick_sync::mbar_arrive(&bar_pempty[q % NPQ]); });
                else
                  brick_sync::mbar_arrive(&bar_pempty[q % NPQ]);
              }
          }
      }
    else
      {
        for (int q = 0; q < nq; ++q)
          {
            load_plain(q);
            __syncthreads();
            if (xwarp < C::XW) x_phase(q);
            __syncthreads();
            if (yz_warp) yz_phase(q, []() {});
          }
      }
    // the top plane of the chunk: shared with the chunk above; at the top of the launch range it is complete (top of the
    // mesh) or the partial sum of the cells below (z-slab launches)
    if (yz_warp) store_top(K * cz1, cz1 < a.zhi);
  }
} // namespace stfem
