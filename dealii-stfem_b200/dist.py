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
    "stfem_ctx_allreduce": (C.c_int, [_vp, _dp, C.c_int, C.c_int]),
    "stfem_partition_brick": (C.c_int, [C.c_int, _ip, _dp, _dp, _ip, _ip, _ip, _ip, _dp, _dp, C.POINTER(C.c_uint)]),
    "stfem_mesh_set_partition": (C.c_int, [_vp, _ip, _ip]),
    "stfem_op_halo_add": (C.c_int, [_vp, _vpp, C.c_int]),
    "stfem_mesh_set_ghost_vertices": (C.c_int, [_vp, _dp]),
    "stfem_op_set_ghost_coefficients": (C.c_int, [_vp, _dp, _dp]),
    "stfem_halo_emulate_host": (C.c_int, [C.c_int, _ip, _ip, C.c_int, _dp]),
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


def allreduce(ctx, values, op="sum"):
    """Host values reduced over the ranks of the context's communicator (stfem_ctx_allreduce): returns a new float64 array.
    op: "sum", "max" or "min".  A context without a communicator returns the values unchanged."""
    v = np.ascontiguousarray(np.atleast_1d(values), np.float64).copy()
    capi.check(capi.lib().stfem_ctx_allreduce(ctx.h, capi._dptr(v), int(v.size), {"sum": 0, "max": 1, "min": 2}[op]))
    return v


def parity_check(ctx, device, rank, world, refinement=2, vmult_cells=12):
    """Partitioned run against the single-GPU run of the same GLOBAL problem, computed redundantly by every rank on its own
    GPU (a second context without communicator): returns a dict of relative errors / iteration counts and "ok".
      vmult     Q4 x cG(2) operator, vmult_cells^3 cells per rank, FP64                              (tolerance 1e-12)
      vcycle    one STMG V-cycle, Q2 x DG(1), 2 subdivisions, `refinement` refinements, FP32 levels   (2e-3)
      solve     one time step (rhs + FGMRES): iteration counts +-1, solution                          (1e-8)
    Every rank must call it (collective); ctx is the context that owns the communicator."""
    from . import driver, fe_time_host
    grid = proc_grid_for(world, 3)
    coords = coords_of(rank, grid)
    res = {}
    ctx0 = capi.Context(device)

    def brick_of(a, nb, npg, npl, off):
        sl = (slice(None), slice(off[2], off[2] + npl[2]), slice(off[1], off[1] + npl[1]), slice(off[0], off[0] + npl[0]))
        return np.ascontiguousarray(a.reshape(nb, npg[2], npg[1], npg[0])[sl]).reshape(nb, -1)

    # ---- (1) the headline operator on a partitioned brick (halo overlap path, all kernels of the fine level)
    k = 4
    A, B = fe_time_host.get_fe_time_weights("CGP", 2, 2.0 ** -6, 1)[:2]
    nb = A.shape[0]
    ng = [vmult_cells * g for g in grid]
    up = [float(g) for g in grid]
    gm = capi.Mesh(ctx0, ng, upper=up)
    gop = capi.Operator(gm, k, A, B)
    xg = np.random.RandomState(7).uniform(-1, 1, (nb, gop.n))
    dx, dy = gop.new_vector().upload(xg), gop.new_vector()
    gop.vmult(dy, dx)
    Ag = dy.download()
    dx.free(); dy.free(); gop.close(); gm.close()
    n_loc, _, llo, lup, mask = partition_brick(ng, [0.0] * 3, up, grid, coords)
    pm = capi.Mesh(ctx, n_loc, lower=llo, upper=lup, dirichlet_faces=mask)
    set_partition(pm, grid, coords)
    pop = capi.Operator(pm, k, A, B)
    npg, npl = [k * n + 1 for n in ng], [k * n + 1 for n in n_loc]
    off = [k * n_loc[d] * coords[d] for d in range(3)]
    dx, dy = pop.new_vector().upload(brick_of(xg, nb, npg, npl, off)), pop.new_vector()
    pop.vmult(dy, dx)
    res["vmult"] = float(np.abs(dy.download() - brick_of(Ag, nb, npg, npl, off)).max() / np.abs(Ag).max())
    dx.free(); dy.free(); pop.close(); pm.close()

    # ---- (2), (3) V-cycle and a full time step through the product driver
    k, r = 2, 1
    pj = {"timeType": "DG", "problemType": "heat", "feDegree": r, "refinement": refinement, "subdivisions": "2,2,2",
          "mgTimeBeforeSpace": "true", "smoother": "relaxation", "spaceTimeConvergenceTest": "true"}
    p = driver.parse_parameters(pj, 3)
    glob = driver.HeatWaveProblem(ctx0, p, 3, refinement, r, space_degree=k)
    ng = [2 * (1 << refinement)] * 3
    npg = [k * n + 1 for n in ng]
    nb = glob.nb
    xg = np.random.RandomState(7).uniform(-1, 1, (nb, glob.n))
    dx, dy = glob.matrix.new_vector().upload(xg), glob.matrix.new_vector()
    glob.matrix.vmult(dy, dx)
    Ag = dy.download()
    glob.mg.vmult(dy, dx.upload(Ag))
    Vg = dy.download()
    it_g = glob.step(evaluate_error=False)
    sol_g = glob.x.download()
    dx.free(); dy.free()
    part = driver.HeatWaveProblem(ctx, p, 3, refinement, r, space_degree=k, partition=(grid, coords))
    nl = [n // g for n, g in zip(ng, grid)]
    npl = [k * n + 1 for n in nl]
    off = [k * nl[d] * coords[d] for d in range(3)]
    dx, dy = part.matrix.new_vector(), part.matrix.new_vector()
    part.mg.vmult(dy, dx.upload(brick_of(Ag, nb, npg, npl, off)))
    res["vcycle"] = float(np.abs(dy.download() - brick_of(Vg, nb, npg, npl, off)).max() / np.abs(Vg).max())
    it_p = part.step(evaluate_error=False)
    res["iterations_global"], res["iterations_partitioned"] = int(it_g), int(it_p)
    res["solve"] = float(np.abs(part.x.download() - brick_of(sol_g, nb, npg, npl, off)).max() / np.abs(sol_g).max())
    dx.free(); dy.free(); part.close(); glob.close(); ctx0.close()
    ok = res["vmult"] <= 1e-12 and res["vcycle"] <= 2e-3 and abs(it_p - it_g) <= 1 and res["solve"] <= 1e-8
    # every rank must agree: the worst error / flag over the ranks
    worst = allreduce(ctx, [res["vmult"], res["vcycle"], res["solve"], 0.0 if ok else 1.0], "max")
    res["vmult"], res["vcycle"], res["solve"] = float(worst[0]), float(worst[1]), float(worst[2])
    res["ok"] = bool(worst[3] == 0.0)
    res["tolerances"] = {"vmult": 1e-12, "vcycle": 2e-3, "solve": 1e-8, "iterations": 1}
    res["grid"] = list(grid)
    return res
