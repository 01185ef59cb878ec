"""Multi-process logic of the box partition on CPU (gloo, world_size 2 and 4): brick extents from the
product's stfem_partition_brick, direction-by-direction compress-add of interface DoFs and the masked
dot product + all-reduce, emulated with numpy/gloo around the oracle's local operator, must reproduce the
global operator and the global inner product.  This is the algorithm csrc/dist.cuh runs over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _plane(arr, axis, index):
    # arr [nb, (nz,) ny, nx]; coordinate axis -> array axis
    ax = arr.ndim - 1 - axis
    sl = [slice(None)] * arr.ndim
    sl[ax] = index
    return tuple(sl)


def _worker(rank, world, port, dim, n_global, k, results):
    import dealii_stfem_b200 as st
    from oracle import fe_time as ft, spatial as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grid = st.dist.proc_grid_for(world, dim)
    coords = st.dist.coords_of(rank, grid)
    lower, upper = [0.0] * dim, [1.0] * dim
    n_local, off, llo, lup, mask = st.dist.partition_brick(n_global, lower, upper, grid, coords)
    A, B, _, _ = ft.get_fe_time_weights("DG", 1, 0.05, 1)
    nb = A.shape[0]
    # global problem (every rank builds it; small) and global reference result
    gmesh = S.Mesh(dim, n_global, 0)
    gspace = S.Space(gmesh, k)
    gsys = S.SystemMatrix(S.MatrixFreeOperator(gspace, 0, 1), S.MatrixFreeOperator(gspace, 1, 0), A, B)
    gsrc = np.stack([np.random.RandomState(3 + b).uniform(-1, 1, gspace.n_dofs) for b in range(nb)])
    gref = gsys.vmult(gsrc)
    gshape = [nb] + gspace.np[::-1]
    # local brick
    lmesh = S.Mesh(dim, n_local, 0, lower=llo, upper=lup)
    lspace = S.Space(lmesh, k, dirichlet_faces=mask)
    lsys = S.SystemMatrix(S.MatrixFreeOperator(lspace, 0, 1), S.MatrixFreeOperator(lspace, 1, 0), A, B)
    sl = tuple([slice(None)] + [slice(k * off[dim - 1 - ax], k * off[dim - 1 - ax] + lspace.np[dim - 1 - ax]) for ax in range(dim)])
    lsrc = gsrc.reshape(gshape)[sl].reshape(nb, -1)
    lshape = [nb] + lspace.np[::-1]
    ldst = lsys.vmult(lsrc).reshape(lshape)
    # compress-add, direction by direction (csrc/dist.cuh halo_compress_add)
    for d in range(dim):
        recv = {}
        reqs = []
        for s in (0, 1):
            c = list(coords)
            c[d] += -1 if s == 0 else 1
            if c[d] < 0 or c[d] >= grid[d]:
                continue
            nbr = sum(c[a] * int(np.prod(grid[:a])) for a in range(dim))
            send = torch.from_numpy(np.ascontiguousarray(ldst[_plane(ldst, d, 0 if s == 0 else -1)]))
            buf = torch.empty_like(send)
            recv[s] = buf
            reqs.append(dist.isend(send, nbr))
            reqs.append(dist.irecv(buf, nbr))
        for r in reqs:
            r.wait()
        for s, buf in recv.items():
            ldst[_plane(ldst, d, 0 if s == 0 else -1)] += buf.numpy()
    err = np.abs(ldst - gref.reshape(gshape)[sl]).max() / np.abs(gref).max()
    # masked dot + allreduce
    w = np.ones(lshape[1:], bool)
    for d in range(dim):
        if coords[d] < grid[d] - 1:
            w[_plane(w[None], d, -1)[1:]] = False
    x, y = lsrc.reshape(lshape), ldst
    local = torch.tensor([float((x * y * w[None]).sum())], dtype=torch.float64)
    dist.all_reduce(local)
    gdot = float((gsrc * gref).sum())
    derr = abs(local.item() - gdot) / abs(gdot)
    results[rank] = (err, derr)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,dim,n_global,k", [(2, 2, [4, 4], 2), (2, 3, [4, 2, 2], 2), (4, 2, [4, 4], 3), (4, 3, [4, 4, 2], 1)])
def test_box_partition_reproduces_global_operator(world, dim, n_global, k):
    mgr = mp.Manager()
    results = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, dim, n_global, k, results), nprocs=world, join=True)
    assert len(results) == world
    for r in range(world):
        err, derr = results[r]
        assert err < 1e-13, (r, err)
        assert derr < 1e-13, (r, derr)


# ----------------------------------------------------------------------------------------------------------------
# Partitioned multigrid transfers (csrc/mg.cuh restrict_level / prolongate_level): the global restriction must equal
#   sum over ranks of [ local restriction of the local residual with interface entries weighted by 1/multiplicity ]
# placed at the rank's brick of the coarse level and summed (halo sum on partitioned coarse levels, all-reduce of the
# global coarse vector when the coarse level is agglomerated), and the global prolongation restricted to a brick
# must equal the local prolongation of the brick of the coarse vector.
def _transfer_worker(rank, world, port, dim, n_fine, k, results):
    import dealii_stfem_b200 as st
    from oracle import spatial as S, stmg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grid = st.dist.proc_grid_for(world, dim)
    coords = st.dist.coords_of(rank, grid)
    lower, upper = [0.0] * dim, [1.0] * dim
    n_coarse = [n // 2 for n in n_fine]
    # global levels and the global transfer (every rank builds them; small)
    gf = S.Space(S.Mesh(dim, n_fine, 0), k)
    gc = S.Space(S.Mesh(dim, n_coarse, 0), k)
    P = stmg.space_prolongation(gc, gf, np.float64)
    rng = np.random.RandomState(5)
    r_f = rng.uniform(-1, 1, gf.n_dofs) * (~gf.constrained)
    x_c = rng.uniform(-1, 1, gc.n_dofs) * (~gc.constrained)
    R_ref = P.T @ r_f
    Px_ref = P @ x_c
    # local bricks
    nlf, off_f, llo, lup, mask = st.dist.partition_brick(n_fine, lower, upper, grid, coords)
    nlc, off_c, _, _, _ = st.dist.partition_brick(n_coarse, lower, upper, grid, coords)
    lf = S.Space(S.Mesh(dim, nlf, 0, lower=llo, upper=lup), k, dirichlet_faces=mask)
    lc = S.Space(S.Mesh(dim, nlc, 0, lower=llo, upper=lup), k, dirichlet_faces=mask)
    Pl = stmg.space_prolongation(lc, lf, np.float64)

    def brick(arr, space_g, space_l, off):
        sl = tuple(slice(k * off[dim - 1 - ax], k * off[dim - 1 - ax] + space_l.np[dim - 1 - ax]) for ax in range(dim))
        return arr.reshape(space_g.np[::-1])[sl]

    # restriction: weight interface entries, restrict locally, add the bricks into the global coarse vector
    rl = brick(r_f, gf, lf, off_f).copy()
    for d in range(dim):
        ax = dim - 1 - d
        for s, has in ((0, coords[d] > 0), (-1, coords[d] < grid[d] - 1)):
            if has:
                idx = [slice(None)] * dim
                idx[ax] = s
                rl[tuple(idx)] *= 0.5
    Rl = (Pl.T @ rl.reshape(-1)).reshape(lc.np[::-1])
    glob = np.zeros(gc.np[::-1])
    sl_c = tuple(slice(k * off_c[dim - 1 - ax], k * off_c[dim - 1 - ax] + lc.np[dim - 1 - ax]) for ax in range(dim))
    glob[sl_c] += Rl
    t = torch.from_numpy(glob)
    dist.all_reduce(t)                                   # = agglomerated switch; = halo sums on a partitioned level
    err_r = np.abs(t.numpy().reshape(-1) - R_ref).max() / np.abs(R_ref).max()
    # prolongation: brick of the coarse vector, local prolongation == brick of the global prolongation
    xl = brick(x_c, gc, lc, off_c).reshape(-1)
    err_p = np.abs((Pl @ xl) - brick(Px_ref, gf, lf, off_f).reshape(-1)).max() / np.abs(Px_ref).max()
    results[rank] = (err_r, err_p)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,dim,n_fine,k", [(2, 2, [8, 4], 2), (4, 2, [4, 8], 1), (4, 3, [4, 4, 4], 2), (2, 3, [4, 2, 2], 3)])
def test_partitioned_transfers_reproduce_global_transfers(world, dim, n_fine, k):
    mgr = mp.Manager()
    results = mgr.dict()
    port = _free_port()
    mp.spawn(_transfer_worker, args=(world, port, dim, n_fine, k, results), nprocs=world, join=True)
    assert len(results) == world
    for r in range(world):
        err_r, err_p = results[r]
        assert err_r < 1e-13, (r, err_r)
        assert err_p < 1e-13, (r, err_p)


def test_ghost_layer_vertices_are_slices_of_the_global_mesh():
    """Host logic of the ghost cell layer (dist.ghost_layers / brick_vertices, used by the driver for partitioned
    general-geometry meshes): the union of the local bricks is the global mesh, and a brick's ghost-extended vertex array
    contains exactly the first vertex layer of its neighbours."""
    import dealii_stfem_b200 as st
    n, grid = [4, 6, 2], [2, 2, 1]
    g = [np.linspace(0.0, 1.0, m + 1) for m in n]
    V = np.stack(np.meshgrid(g[2], g[1], g[0], indexing="ij")[::-1], axis=-1)          # [z][y][x][xyz]
    V = V + 0.01 * np.random.RandomState(0).uniform(-1, 1, V.shape)
    seen = np.zeros(V.shape[:3], bool)
    for rank in range(4):
        coords = st.dist.coords_of(rank, grid)
        n_loc, off, llo, lup, mask = st.dist.partition_brick(n, [0.0] * 3, [1.0] * 3, grid, coords)
        glo, ghi = st.dist.ghost_layers(grid, coords)
        assert glo == [int(c > 0) for c in coords] and ghi == [int(c < gg - 1) for c, gg in zip(coords, grid)]
        v_loc = st.dist.brick_vertices(V, n, off, n_loc).reshape(n_loc[2] + 1, n_loc[1] + 1, n_loc[0] + 1, 3)
        v_ext = st.dist.brick_vertices(V, n, off, n_loc, glo, ghi)
        ne = [n_loc[a] + glo[a] + ghi[a] for a in range(3)]
        v_ext = v_ext.reshape(ne[2] + 1, ne[1] + 1, ne[0] + 1, 3)
        # the local brick sits inside the extended one at offset g_lo
        inner = v_ext[glo[2]:glo[2] + n_loc[2] + 1, glo[1]:glo[1] + n_loc[1] + 1, glo[0]:glo[0] + n_loc[0] + 1]
        assert np.array_equal(inner, v_loc)
        # ... and both are slices of the global array
        assert np.array_equal(v_loc, V[off[2]:off[2] + n_loc[2] + 1, off[1]:off[1] + n_loc[1] + 1, off[0]:off[0] + n_loc[0] + 1])
        if glo[0]:
            assert np.array_equal(v_ext[glo[2]:glo[2] + n_loc[2] + 1, glo[1]:glo[1] + n_loc[1] + 1, 0],
                                  V[off[2]:off[2] + n_loc[2] + 1, off[1]:off[1] + n_loc[1] + 1, off[0] - 1])
        if ghi[1]:
            assert np.array_equal(v_ext[glo[2]:glo[2] + n_loc[2] + 1, -1, glo[0]:glo[0] + n_loc[0] + 1],
                                  V[off[2]:off[2] + n_loc[2] + 1, off[1] + n_loc[1] + 1, off[0]:off[0] + n_loc[0] + 1])
        seen[off[2]:off[2] + n_loc[2] + 1, off[1]:off[1] + n_loc[1] + 1, off[0]:off[0] + n_loc[0] + 1] = True
    assert seen.all()
