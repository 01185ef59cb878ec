"""Multi-GPU helpers over the C ABI: NCCL bootstrap through torch.distributed (plumbing only) and the
box partition of the structured mesh."""
import ctypes as C

import numpy as np

from . import capi

_vp, _vpp, _dp, _ip = C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_double), C.POINTER(C.c_int)

capi.SYMBOLS.update({
    "stfem_comm_unique_id": (C.c_int, [C.c_char_p]),
    "stfem_ctx_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, C.c_char_p]),
    "stfem_ctx_comm_destroy": (C.c_int, [_vp]),
    "stfem_ctx_rank": (C.c_int, [_vp]),
    "stfem_ctx_n_ranks": (C.c_int, [_vp]),
    "stfem_partition_brick": (C.c_int, [C.c_int, _ip, _dp, _dp, _ip, _ip, _ip, _ip, _dp, _dp, C.POINTER(C.c_uint)]),
    "stfem_mesh_set_partition": (C.c_int, [_vp, _ip, _ip]),
    "stfem_op_halo_add": (C.c_int, [_vp, _vpp, C.c_int]),
})


def proc_grid_for(n_ranks, dim=3):
    """1, 2x1x1, 2x2x1, 2x2x2, ...: split the directions round-robin by factors of 2."""
    g = [1] * dim
    d = 0
    n = n_ranks
    while n > 1:
        assert n % 2 == 0, "number of ranks must be a power of two"
        g[d % dim] *= 2
        n //= 2
        d += 1
    return g


def coords_of(rank, grid):
    c = []
    for g in grid:
        c.append(rank % g)
        rank //= g
    return c


def partition_brick(n_global, lower, upper, proc_grid, coords):
    """stfem_partition_brick: (n_local, cell_offset, local_lower, local_upper, dirichlet mask)."""
    dim = len(n_global)
    ng = (C.c_int * dim)(*n_global)
    pg = (C.c_int * dim)(*proc_grid)
    co = (C.c_int * dim)(*coords)
    nl, off = (C.c_int * dim)(), (C.c_int * dim)()
    lo = np.asarray(lower, np.float64)
    up = np.asarray(upper, np.float64)
    llo, lup = np.zeros(dim), np.zeros(dim)
    mask = C.c_uint()
    capi.check(capi.lib().stfem_partition_brick(dim, ng, capi._dptr(lo), capi._dptr(up), pg, co, nl, off, capi._dptr(llo),
                                                capi._dptr(lup), C.byref(mask)))
    return list(nl), list(off), llo, lup, mask.value


def init_comm(ctx, rank, world, broadcast_bytes):
    """broadcast_bytes(b: bytes|None) -> bytes distributes rank 0's unique id to all ranks."""
    buf = C.create_string_buffer(128)
    if rank == 0:
        capi.check(capi.lib().stfem_comm_unique_id(buf))
    uid = broadcast_bytes(bytes(buf.raw) if rank == 0 else None)
    capi.check(capi.lib().stfem_ctx_comm_init(ctx.h, rank, world, uid))


def set_partition(mesh, proc_grid, coords):
    dim = mesh.dim
    capi.check(capi.lib().stfem_mesh_set_partition(mesh.h, (C.c_int * dim)(*proc_grid), (C.c_int * dim)(*coords)))
