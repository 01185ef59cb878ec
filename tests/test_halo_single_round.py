"""Single-round interface exchange of the box partition (csrc/dist.cuh: HaloPlan, halo_pack_element, halo_unpack_element) on
the CPU: the library's host-only hook runs the exchange among all bricks of a process grid inside one process, with the very
element functions the CUDA kernels execute and memcpy in place of ncclSend / ncclRecv.  Replaces, for the reference, the
compress(add) of deal.II's distributed vectors inside MatrixFree::cell_loop (include/operators.h:1016).
Expected: every copy of an interface node holds the sum of the partial values of all bricks sharing it, added in the order of
the ranks - bit-identical on every rank."""
import ctypes as C

import numpy as np
import pytest

import dealii_stfem_b200 as st


def exchange(dim, grid, npn, nb, data):
    g = (C.c_int * dim)(*grid)
    n = (C.c_int * dim)(*npn)
    st.capi.check(st.capi.lib().stfem_halo_emulate_host(dim, g, n, nb, st.capi._dptr(data)))


@pytest.mark.parametrize("dim,grid,npn,nb", [
    (3, [2, 1, 1], [5, 4, 3], 2),
    (3, [2, 2, 1], [4, 5, 3], 1),
    (3, [2, 2, 2], [5, 5, 5], 2),
    (3, [3, 2, 2], [3, 4, 5], 3),       # interior bricks with neighbours on both sides
    (3, [1, 1, 4], [3, 3, 2], 1),
    (2, [2, 2], [6, 5], 2),
    (2, [3, 1], [4, 4], 1),
])
def test_single_round_exchange_sums_in_rank_order(dim, grid, npn, nb):
    n_ranks = int(np.prod(grid))
    N = int(np.prod(npn))
    rng = np.random.RandomState(11)
    data = rng.uniform(-1, 1, (n_ranks, nb, N))
    before = data.copy()
    exchange(dim, grid, npn, nb, data)
    # reference: global node index of every local node; bricks overlap in one node plane per interface
    g3 = list(grid) + [1] * (3 - dim)
    n3 = list(npn) + [1] * (3 - dim)
    gl = [g3[d] * (n3[d] - 1) + 1 for d in range(3)]
    contrib = {}                                          # global node -> list of (rank, local index)
    for r in range(n_ranks):
        c = [r % g3[0], (r // g3[0]) % g3[1], r // (g3[0] * g3[1])]
        for iz in range(n3[2]):
            for iy in range(n3[1]):
                for ix in range(n3[0]):
                    key = (c[0] * (n3[0] - 1) + ix, c[1] * (n3[1] - 1) + iy, c[2] * (n3[2] - 1) + iz)
                    contrib.setdefault(key, []).append((r, ix + n3[0] * (iy + n3[1] * iz)))
    assert len(contrib) == gl[0] * gl[1] * gl[2]
    shared = 0
    for key, lst in contrib.items():
        lst.sort()
        for b in range(nb):
            total = before[lst[0][0], b, lst[0][1]]
            for r, i in lst[1:]:
                total = total + before[r, b, i]           # the order the kernel adds in
            for r, i in lst:
                assert data[r, b, i] == total, (key, lst)
        shared += len(lst) > 1
    assert shared > 0


def test_exchange_is_the_identity_on_one_rank():
    data = np.arange(24.0).reshape(1, 2, 12)
    ref = data.copy()
    exchange(3, [1, 1, 1], [3, 2, 2], 2, data)
    assert np.array_equal(data, ref)
