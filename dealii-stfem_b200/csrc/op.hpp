// The operator object behind stfem_op_t.
#pragma once
#include "basis_host.hpp"
#include "common.hpp"
#include "dist.cuh"

struct stfem_op
{
  stfem_mesh *mesh = nullptr;
  int         degree = 1, number_type = STFEM_F64, nb_rows = 1, nb_cols = 1, variant = 0;
  int         np[3] = {1, 1, 1};
  long long   N = 0; // spatial dofs per block
  std::unique_ptr<stfem::ShapeHost> shape;
  std::vector<double> Alpha, Beta; // host copies, row-major nb_rows x nb_cols
  std::vector<double> AlphaT, BetaT, AlphaNeg, BetaNeg; // host copies of the transposed / negated device matrices below
  void *d_alpha = nullptr, *d_beta = nullptr, *d_alphaT = nullptr, *d_betaT = nullptr;
  void *d_alpha_neg = nullptr, *d_beta_neg = nullptr; // -Alpha, -Beta: residual r = b - A x in one cell loop
  void *d_metric = nullptr; // general geometry: per cell, per q-point metric (+JxW)
  void *d_metric_plane = nullptr; // the same in the order st_vmult_plane_kernel reads it (3D, degree <= 4)
  void *d_coeff = nullptr;  // per-cell Laplace coefficient
  bool  geom_otf = false;   // general geometry in 3D: metric computed on the fly from the cell vertices (st_vmult_plane.cuh, OTF)
  std::vector<double> h_coeff_cell, h_coeff_q; // host copies (Vanka set-up works in double)
  std::vector<double> h_coeff_cell_ghost, h_coeff_q_ghost; // partitioned meshes: the same over the brick + its ghost cell layer
  std::vector<double> fd_V, fd_lam; // kernel_variant 60: modes of the reference-cell pencil (Kh, Mh), computed on first use
  std::vector<void *> d_scratch; // device staging for the host-buffer entry points
  std::vector<void *> d_part_scratch; // partitioned meshes: increment of an accumulating apply before the halo sum
  void  *d_xface = nullptr, *h_xface = nullptr; // host-buffer entry point on partitioned meshes: packed x faces (device, pinned host)
  size_t xface_bytes = 0;
  stfem::HaloBuffers halo;
  // launch context of the next kernel dispatch (set by op_apply's callers inside this file): cell sub-box, stream
  const int   *box_lo = nullptr, *box_n = nullptr;
  int          n_xbox = 0;           // > 0: a list of boxes in one launch (Cartesian kernel only)
  int          xbox_lo[6][3], xbox_n[6][3];
  cudaStream_t launch_stream = nullptr;
  bool  timing = false;
  float last_ms = 0.f;
};

namespace stfem
{
  int ctx_ensure_aux(stfem_ctx *ctx);
  // dst (+)= A src with explicit time matrices (device pointers in the operator's number type); rhs != nullptr (with
  // zero_dst): dst = rhs + A src in one pass (the residual b - A x with the negated matrices)
  int op_apply(stfem_op *op, void *const *dst, const void *const *src, int nb_src, int nb_dst, const void *alpha,
               const void *beta, bool zero_dst, const void *const *rhs = nullptr);
  // brick kernel (brick.cu): can this application run through it / run it.  The launch honours op->box_lo / box_n in z
  // (cell layers) and op->launch_stream; first_plane_acc: the first node plane of the z range is accumulated
  bool brick_eligible(const stfem_op *op, int nb_src, int nb_dst, const void *alpha, const void *beta);
  // mode 0: dst = A src, 1: dst += A src, 2: dst = rhs + A src
  int  brick_launch(stfem_op *op, void *const *dst, const void *const *src, int nb, const void *alpha, const void *beta, int mode,
                    const void *const *rhs, bool first_plane_acc);
  // diag K, diag M (double, cudaMalloc'ed, caller frees), constrained rows 0; capi_op.cu
  int op_spatial_diagonals(stfem_op *op, double **dK, double **dM);
} // namespace stfem
