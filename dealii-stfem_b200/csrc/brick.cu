// The brick kernel's translation unit (st_vmult_brick.cuh): kept apart from capi_op.cu so that the two compile in parallel.
#include "st_vmult_brick_host.cuh"

namespace stfem
{
  // BRICK kernel (st_vmult_brick.cuh): 3D Cartesian meshes without a coefficient table, square time matrices that are one
  // of the operator's own (so that their host copies can be passed by value), degree 2-4, <= 3 blocks.
  // kernel_variant: 0 default (TMA loads; meshes below 512 cells keep the per-cell kernel), 70 plain loads, 77 split warps,
  // 79 barrier pipeline, 90 = 72 with the per-SM alternation of the X warps (measured slower),
  // 72 = 0 without the size threshold, 80.. forced number of z chunks (variant - 79), 3 = the
  // per-cell kernel of round 1 (st_vmult_cart.cuh) instead.
  bool brick_eligible(const stfem_op *op, int nb_src, int nb_dst, const void *alpha, const void *beta)
  {
    static const bool off = std::getenv("STFEM_NO_BRICK") != nullptr;
    const stfem_mesh *m = op->mesh;
    if (off || m->dim != 3 || !m->cartesian || op->d_metric || op->d_coeff) return false;
    if (!(op->variant == 0 || (op->variant >= 70 && op->variant <= 90))) return false;
    if (nb_src != nb_dst || nb_dst < 1 || nb_dst > 3 || op->degree < 2 || op->degree > 4) return false;
    if (!brick_host_matrix(op, alpha) || !brick_host_matrix(op, beta)) return false;
    if (op->n_xbox > 0) return false;
    if (op->box_lo && (op->box_lo[0] != 0 || op->box_lo[1] != 0 || op->box_n[0] != m->n[0] || op->box_n[1] != m->n[1])) return false;
    // tiny levels are launch-latency bound: the march through z only adds latency there
    if (op->variant == 0 && m->n_cells < 512) return false;
    // one block: only CY warps in the Y+Z phase against ceil((K CY + K + 1) / rows per warp) in the X phase - the per-cell kernel is faster (measured)
    if (op->variant == 0 && nb_dst == 1) return false;
    return true;
  }

  template <typename T>
  static int launch_brick_any(stfem_op *op, void *const *dst, const void *const *src, int nb, const void *alpha, const void *beta, int mode,
                              const void *const *rhs, bool first_plane_acc)
  {
    const std::vector<double> &hA = *brick_host_matrix(op, alpha), &hB = *brick_host_matrix(op, beta);
    const int  zlo = op->box_lo ? op->box_lo[2] : 0, zhi = op->box_lo ? op->box_lo[2] + op->box_n[2] : op->mesh->n[2];
    static const bool no_tma = std::getenv("STFEM_BRICK_NO_TMA") != nullptr;
    const int  tma = (op->variant == 70 || no_tma) ? 0 : (op->variant == 79 ? 2 : 1); // 79: full / empty barrier pipeline instead of CTA barriers
    const int  n_chunks = op->variant >= 80 && op->variant < 90 ? op->variant - 79 : 0; // 0: chosen by the launcher
    // tuning variants of the headline instance (Q4, two blocks): other tile heights / CTAs per SM
    if (op->variant == 77 && op->degree == 4 && nb == 2) // X and Y+Z phases on separate warps, one CTA per SM
      return launch_brick<5, 2, T, 7, 4, 1, true>(op, dst, src, hA, hB, mode, rhs, zlo, zhi, first_plane_acc, tma, n_chunks);
#define STFEM_BRICK_CASE(N1_, NB_, MINB_)                                                                                             \
  if (op->degree + 1 == N1_ && nb == NB_)                                                                                             \
    {                                                                                                                                 \
      return launch_brick<N1_, NB_, T, BrickTile<N1_>::CX, BrickTile<N1_>::CY, MINB_>(op, dst, src, hA, hB, mode, rhs, zlo, zhi, first_plane_acc, tma, \
                                                                                      n_chunks);                                     \
    }
    STFEM_BRICK_CASE(3, 1, 2) STFEM_BRICK_CASE(3, 2, 2) STFEM_BRICK_CASE(3, 3, 1)
    STFEM_BRICK_CASE(4, 1, 2) STFEM_BRICK_CASE(4, 2, 2) STFEM_BRICK_CASE(4, 3, 1)
    STFEM_BRICK_CASE(5, 1, 2) STFEM_BRICK_CASE(5, 2, 2) STFEM_BRICK_CASE(5, 3, 1)
#undef STFEM_BRICK_CASE
    set_error("st_vmult (brick): degree %d with %d blocks is not instantiated", op->degree, nb);
    return STFEM_ERR_UNSUPPORTED;
  }

  int brick_launch(stfem_op *op, void *const *dst, const void *const *src, int nb, const void *alpha, const void *beta, int mode,
                   const void *const *rhs, bool first_plane_acc)
  {
    return op->number_type == STFEM_F64 ? launch_brick_any<double>(op, dst, src, nb, alpha, beta, mode, rhs, first_plane_acc) :
                                          launch_brick_any<float>(op, dst, src, nb, alpha, beta, mode, rhs, first_plane_acc);
  }
} // namespace stfem
