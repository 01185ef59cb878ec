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


def ghost_layers(proc_grid, coords):
    """(g_lo, g_hi) per direction: 1 where another rank continues the mesh (one ghost cell layer), else 0."""
    dim = len(proc_grid)
    return [1 if coords[a] > 0 else 0 for a in range(dim)], [1 if coords[a] < proc_grid[a] - 1 else 0 for a in range(dim)]


def brick_vertices(vertices_global, n_global, cell_offset, n_local, g_lo=None, g_hi=None):
    """Vertices of one brick of a structured mesh (optionally extended by its ghost layers) out of the global lexicographic
    vertex array ([..., z][y][x][xyz], x fastest): returns [(n_local + g_lo + g_hi + 1) points per direction, dim]."""
    dim = len(n_global)
    g_lo = [0] * dim if g_lo is None else g_lo
    g_hi = [0] * dim if g_hi is None else g_hi
    vg = np.asarray(vertices_global, np.float64).reshape([m + 1 for m in n_global[::-1]] + [dim])
    sl = tuple(slice(cell_offset[a] - g_lo[a], cell_offset[a] + n_local[a] + g_hi[a] + 1) for a in range(dim))[::-1]
    return np.ascontiguousarray(vg[sl]).reshape(-1, dim)


def set_ghost_vertices(mesh, vertices_ext):
    """stfem_mesh_set_ghost_vertices: vertices of the local brick + one cell layer across every shared face."""
    v = np.ascontiguousarray(vertices_ext, np.float64)
    capi.check(capi.lib().stfem_mesh_set_ghost_vertices(mesh.h, capi._dptr(v)))


def set_ghost_coefficients(op, coeff_cell_ext=None, coeff_q_ext=None):
    """stfem_op_set_ghost_coefficients: the operator's Laplace coefficient over the ghost-extended brick."""
    c = None if coeff_cell_ext is None else np.ascontiguousarray(coeff_cell_ext, np.float64)
    q = None if coeff_q_ext is None else np.ascontiguousarray(coeff_q_ext, np.float64)
    capi.check(capi.lib().stfem_op_set_ghost_coefficients(op.h, None if c is None else capi._dptr(c), None if q is None else capi._dptr(q)))


def allreduce(ctx, values, op="sum"):
    """Host values reduced over the ranks of the context's communicator (stfem_ctx_allreduce): returns a new float64 array.
    op: "sum", "max" or "min".  A context without a communicator returns the values unchanged."""
    v = np.ascontiguousarray(np.atleast_1d(values), np.float64).copy()
    capi.check(capi.lib().stfem_ctx_allreduce(ctx.h, capi._dptr(v), int(v.size), {"sum": 0, "max": 1, "min": 2}[op]))
    return v


def parity_check(ctx, device, rank, world, refinement=2, vmult_cells=12, extended=False):
    """Partitioned run against the single-GPU run of the same GLOBAL problem, computed redundantly by every rank on its own
    GPU (a second context without communicator): returns a dict of relative errors / iteration counts and "ok".
      vmult     Q4 x cG(2) operator, vmult_cells^3 cells per rank, FP64                              (tolerance 1e-12)
      vcycle    one STMG V-cycle, Q2 x DG(1), 2 subdivisions, `refinement` refinements, FP32 levels   (2e-3)
      solve     one time step (rhs + FGMRES): iteration counts +-1, solution                          (1e-8)
      practical_*  the reference's practical set-up in small (perturbed mesh, coefficient table, dense cell-patch smoother on
                ghost-layer patches, cut-off initial value): operator + matrix diagonal (1e-12), V-cycle (2e-3), two time
                steps (iterations +-1, solution and point functionals 1e-8)
      extended  (scripts/mgpu_check.py only) the practical set-up once more with the point-Jacobi inner preconditioner and
                Chebyshev smoothing: iterations +-1, solution 1e-8
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
    # the same through the host-buffer entry point (slab pipeline + interface nodes downloaded again after the exchange)
    xh, yh = brick_of(xg, nb, npg, npl, off), np.full((nb, pop.n), 7.5)
    pop.vmult_host(yh, xh)
    res["vmult_host"] = float(np.abs(yh - brick_of(Ag, nb, npg, npl, off)).max() / np.abs(Ag).max())
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
    it_g0 = glob.step(evaluate_error=False)
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
    res["iterations_global"], res["iterations_partitioned"] = int(it_g0), int(it_p)
    res["solve"] = float(np.abs(part.x.download() - brick_of(sol_g, nb, npg, npl, off)).max() / np.abs(sol_g).max())
    dx.free(); dy.free(); part.close(); glob.close()

    # ---- (4) the reference's practical set-up (configs[3] in small): perturbed MappingQ1 mesh, Coefficient<dim> table on K,
    # dense cell-patch smoother (ghost-layer patches), cut-off initial value, point functionals; two partitioned levels
    k, r = 2, 1

    def vertices(n_cells):
        g = [np.linspace(-1.0, 1.0, m + 1) for m in n_cells]
        V = np.stack(np.meshgrid(g[2], g[1], g[0], indexing="ij")[::-1], axis=-1)     # [z][y][x][xyz]
        d = np.random.RandomState(3).uniform(-1, 1, V.shape) * 0.1 * (2.0 / n_cells[0])
        d[0] = d[-1] = 0
        d[:, 0] = d[:, -1] = 0
        d[:, :, 0] = d[:, :, -1] = 0
        return V + d

    pj = {"timeType": "DG", "problemType": "heat", "feDegree": r, "refinement": refinement, "subdivisions": "2,2,2",
          "hyperRectLowerLeft": "-1,-1,-1", "hyperRectUpperRight": "1,1,1", "mgTimeBeforeSpace": "true", "smoother": "relaxation",
          "spaceTimeConvergenceTest": "false", "distortGrid": 0.1, "distortCoeff": "0.6", "extrapolate": "false"}
    p = driver.parse_parameters(pj, 3)
    ng = [2 * (1 << refinement)] * 3
    Vf = vertices(ng).reshape(-1, 3)
    p["sourcePoint"] = [float(c) for c in Vf[np.argmin(np.sum(Vf * Vf, axis=1))]]
    p["agglomerateBelow"] = 2
    glob = driver.HeatWaveProblem(ctx0, p, 3, refinement, r, space_degree=k, vertices_fn=vertices)
    npg = [k * n + 1 for n in ng]
    nb = glob.nb
    xg = np.random.RandomState(11).uniform(-1, 1, (nb, glob.n))
    dx, dy = glob.matrix.new_vector().upload(xg), glob.matrix.new_vector()
    glob.matrix.vmult(dy, dx)
    Ag = dy.download()
    glob.matrix.diagonal(dy)
    Dg = dy.download()
    glob.mg.vmult(dy, dx.upload(Ag))
    Vg = dy.download()
    it_g = [glob.step(evaluate_error=False) for _ in range(2)]
    sol_g = glob.x.download()
    fun_g = np.array(glob.functional_rows[-1][1:])
    dx.free(); dy.free()
    part = driver.HeatWaveProblem(ctx, p, 3, refinement, r, space_degree=k, vertices_fn=vertices, partition=(grid, coords))
    nl = [n // g for n, g in zip(ng, grid)]
    npl = [k * n + 1 for n in nl]
    off = [k * nl[d] * coords[d] for d in range(3)]
    dx, dy = part.matrix.new_vector(), part.matrix.new_vector()
    part.matrix.vmult(dy, dx.upload(brick_of(xg, nb, npg, npl, off)))
    res["practical_vmult"] = float(np.abs(dy.download() - brick_of(Ag, nb, npg, npl, off)).max() / np.abs(Ag).max())
    part.matrix.diagonal(dy)
    res["practical_diagonal"] = float(np.abs(dy.download() - brick_of(Dg, nb, npg, npl, off)).max() / np.abs(Dg).max())
    part.mg.vmult(dy, dx.upload(brick_of(Ag, nb, npg, npl, off)))
    res["practical_vcycle"] = float(np.abs(dy.download() - brick_of(Vg, nb, npg, npl, off)).max() / np.abs(Vg).max())
    it_pp = [part.step(evaluate_error=False) for _ in range(2)]
    res["practical_iterations_global"], res["practical_iterations_partitioned"] = [int(i) for i in it_g], [int(i) for i in it_pp]
    res["practical_solve"] = float(np.abs(part.x.download() - brick_of(sol_g, nb, npg, npl, off)).max() / np.abs(sol_g).max())
    fun_p = np.array(part.functional_rows[-1][1:])
    res["practical_functionals"] = float(np.abs(fun_p - fun_g).max() / max(np.abs(fun_g).max(), 1e-300))
    dx.free(); dy.free(); part.close(); glob.close()
    ok_jacobi = True
    if extended:
        # the same set-up with the point-Jacobi inner preconditioner (diagonal summed over the ranks) and Chebyshev smoothing
        p2 = dict(p)
        p2["innerPreconditioner"], p2["smoother"], p2["smoothingSteps"] = "jacobi", "chebyshev", 2
        glob = driver.HeatWaveProblem(ctx0, p2, 3, refinement, r, space_degree=k, vertices_fn=vertices)
        it_gj = glob.step(evaluate_error=False)
        sol_gj = glob.x.download()
        part = driver.HeatWaveProblem(ctx, p2, 3, refinement, r, space_degree=k, vertices_fn=vertices, partition=(grid, coords))
        it_pj = part.step(evaluate_error=False)
        res["jacobi_iterations_global"], res["jacobi_iterations_partitioned"] = int(it_gj), int(it_pj)
        res["jacobi_solve"] = float(np.abs(part.x.download() - brick_of(sol_gj, nb, npg, npl, off)).max() / np.abs(sol_gj).max())
        part.close(); glob.close()
        ok_jacobi = abs(it_gj - it_pj) <= 1 and res["jacobi_solve"] <= 1e-8
    ctx0.close()
    ok_practical = ok_jacobi and (res["practical_vmult"] <= 1e-12 and res["practical_diagonal"] <= 1e-12 and res["practical_vcycle"] <= 2e-3 and
                    all(abs(a - b) <= 1 for a, b in zip(it_g, it_pp)) and res["practical_solve"] <= 1e-8 and
                    res["practical_functionals"] <= 1e-8)
    ok = res["vmult"] <= 1e-12 and res["vmult_host"] <= 1e-12 and res["vcycle"] <= 2e-3 and abs(it_p - it_g0) <= 1 and res["solve"] <= 1e-8 and ok_practical
    # every rank must agree: the worst error / flag over the ranks
    keys = ["vmult", "vmult_host", "vcycle", "solve", "practical_vmult", "practical_diagonal", "practical_vcycle", "practical_solve",
            "practical_functionals"] + (["jacobi_solve"] if extended else [])
    worst = allreduce(ctx, [res[kk] for kk in keys] + [0.0 if ok else 1.0], "max")
    for kk, w in zip(keys, worst):
        res[kk] = float(w)
    res["ok"] = bool(worst[-1] == 0.0)
    res["tolerances"] = {"vmult": 1e-12, "diagonal": 1e-12, "vcycle": 2e-3, "solve": 1e-8, "functionals": 1e-8, "iterations": 1}
    res["grid"] = list(grid)
    return res
