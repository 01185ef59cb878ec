"""The brick kernel (csrc/st_vmult_brick.cuh) compiled for the host (tests/cpp/brick_host_emulation.cpp: one std::thread per
CUDA thread, std::barrier for __syncthreads, a rendezvous for __shfl_sync) against the oracle's unfused SystemMatrix::vmult
(oracle/spatial.py, reference include/operators.h:536-559).  CPU only: checks the tile / z-chunk decomposition, the row-class
addressing of the box loads, lane mappings, Dirichlet masks, ragged tiles, accumulate mode and z-slab launches.  The TMA /
mbarrier path itself only runs on the GPU (tests/test_vmult_gpu.py, tests/test_brick_gpu.py)."""
import os
import subprocess

import numpy as np
import pytest

from dealii_stfem_b200 import fe_time_host as fth
from oracle import spatial as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("brick") / "brick_emul")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-Wno-unknown-pragmas", "-o", exe,
                        os.path.join(ROOT, "tests", "cpp", "brick_host_emulation.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def run_emulator(exe, tmp_path, degree, A, B, cells, upper, mask, src, dst0, zlo=0, zhi=None, mode=0, first_plane_acc=0, n_chunks=0,
                 tile=(7, 4), misalign=0, f32=False, flow_b=False, split=False, pipe=False):
    nb = A.shape[0]
    h = [upper[d] / cells[d] for d in range(3)]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    np.concatenate([A.reshape(-1), B.reshape(-1), src.reshape(-1), dst0.reshape(-1)]).astype(np.float64).tofile(fin)
    zhi = cells[2] if zhi is None else zhi
    cmd = [exe, str(degree), str(nb)] + [str(c) for c in cells] + [repr(v) for v in h] + [hex(mask), str(zlo), str(zhi), str(mode),
                                                                                          str(first_plane_acc), str(n_chunks), str(tile[0]),
                                                                                          str(tile[1]), str(misalign), fin, fout]
    if f32:
        cmd.append("f32")
    env = dict(os.environ)
    if split:
        env["BRICK_EMU_SPLIT"] = "1"        # X phase and Y+Z phase on separate warps (kernel template parameter SPLIT)
    if pipe:
        env["BRICK_EMU_PIPE"] = "1"         # full / empty barrier pipeline (kernel_variant 79) instead of one CTA barrier per plane
    if flow_b:
        env["BRICK_EMU_FLOW_B"] = "1"       # plain loads + CTA-wide barriers instead of the mbarrier pipeline of the TMA flow
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    return np.fromfile(fout).reshape(nb, -1), r.stdout


def oracle(degree, ttype, r, cells, upper, mask, seed=42, transpose=False):
    mesh = S.Mesh(3, cells, 0, [0.0, 0.0, 0.0], upper)
    space = S.Space(mesh, degree, dirichlet_faces=mask)
    A, B, _, _ = fth.get_fe_time_weights(ttype, r, 0.05, 1)
    nb = A.shape[0]
    K, M = S.MatrixFreeOperator(space, 0.0, 1.0), S.MatrixFreeOperator(space, 1.0, 0.0)
    src = np.stack([np.random.RandomState(seed + b).uniform(-1, 1, space.n_dofs) for b in range(nb)])
    sysm = S.SystemMatrix(K, M, A, B)
    return space, A, B, src, (sysm.Tvmult(src) if transpose else sysm.vmult(src))


CASES = [
    # degree, ttype, r, cells, upper, mask, tile, n_chunks, misalign
    (4, "CGP", 2, [4, 3, 3], [1.2, 0.8, 1.0], 0x3f, (3, 2), 2, 0),    # 2 x 2 tiles (ragged), 2 z chunks with warm-up layer
    (4, "CGP", 2, [8, 5, 2], [1.0, 1.0, 0.5], 0x3f, (7, 4), 1, 0),    # the product tile of Q4: 2 x 2 tiles, ragged in x and y
    (4, "DG", 1, [3, 2, 4], [1.0, 1.0, 1.0], 0x00, (3, 2), 3, 1),     # no constraints: the last planes of the mesh are stored; odd base
    (4, "DG", 2, [4, 2, 2], [1.0, 0.5, 1.0], 0x15, (3, 2), 0, 0),     # nb = 3, Dirichlet on the lower faces only
    (4, "DG", 0, [3, 3, 2], [1.0, 1.0, 1.0], 0x2a, (3, 2), 2, 0),     # nb = 1, upper faces
    (3, "DG", 1, [5, 3, 3], [1.0, 1.5, 0.5], 0x3f, (4, 2), 2, 0),     # Q3: 5 cells per tile row incl. halo, 6 rows per warp
    (3, "DG", 2, [4, 4, 2], [1.0, 1.0, 1.0], 0x00, (4, 2), 1, 1),
    (2, "CGP", 2, [4, 3, 3], [1.0, 2.0, 0.7], 0x3f, (3, 2), 3, 0),    # Q2
    (2, "DG", 2, [5, 2, 2], [1.0, 1.0, 1.0], 0x0c, (3, 2), 1, 1),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "k%d_%s%d_%s_m%x_t%dx%d_c%d_a%d" % (c[0], c[1], c[2], "x".join(map(str, c[3])), c[5], c[6][0],
                                                                                          c[6][1], c[7], c[8]))
def test_emulated_brick_kernel_matches_oracle(emulator, tmp_path, case):
    degree, ttype, r, cells, upper, mask, tile, n_chunks, misalign = case
    space, A, B, src, ref = oracle(degree, ttype, r, cells, upper, mask)
    garbage = np.full_like(src, 7.5)                  # mode 0 must overwrite every entry, constrained rows with 0
    out, log = run_emulator(emulator, tmp_path, degree, A, B, cells, upper, mask, src, garbage, n_chunks=n_chunks, tile=tile,
                            misalign=misalign)
    assert np.all(np.isfinite(out)), log
    assert np.abs(out - ref).max() <= 1e-13 * np.abs(ref).max(), log
    assert np.all(out[:, space.constrained] == 0)


@pytest.mark.parametrize("case", CASES[:4], ids=lambda c: "k%d_%s%d_%s" % (c[0], c[1], c[2], "x".join(map(str, c[3]))))
def test_emulated_brick_kernel_plain_load_flow(emulator, tmp_path, case):
    """Flow B of the kernel (no tensor maps: all threads copy the boxes, CTA-wide barriers) gives the same result."""
    degree, ttype, r, cells, upper, mask, tile, n_chunks, misalign = case
    space, A, B, src, ref = oracle(degree, ttype, r, cells, upper, mask)
    out, log = run_emulator(emulator, tmp_path, degree, A, B, cells, upper, mask, src, np.full_like(src, 7.5), n_chunks=n_chunks, tile=tile,
                            misalign=misalign, flow_b=True)
    assert "flow B" in log
    assert np.abs(out - ref).max() <= 1e-13 * np.abs(ref).max(), log


@pytest.mark.parametrize("case", [CASES[0], CASES[2], CASES[3], CASES[7]], ids=lambda c: "k%d_%s%d_%s" % (c[0], c[1], c[2], "x".join(map(str, c[3]))))
def test_emulated_brick_kernel_barrier_pipeline(emulator, tmp_path, case):
    """kernel_variant 79: X one plane ahead of Y+Z, coupled by full / empty barriers of three P/Q buffers, no CTA barrier."""
    degree, ttype, r, cells, upper, mask, tile, n_chunks, misalign = case
    space, A, B, src, ref = oracle(degree, ttype, r, cells, upper, mask)
    out, log = run_emulator(emulator, tmp_path, degree, A, B, cells, upper, mask, src, np.full_like(src, 7.5), n_chunks=n_chunks, tile=tile,
                            misalign=misalign, pipe=True)
    assert "A2" in log
    assert np.abs(out - ref).max() <= 1e-13 * np.abs(ref).max(), log


@pytest.mark.parametrize("flow_b", [False, True])
@pytest.mark.parametrize("case", [CASES[0], CASES[1], CASES[3], CASES[5]], ids=lambda c: "k%d_%s%d_%s" % (c[0], c[1], c[2], "x".join(map(str, c[3]))))
def test_emulated_brick_kernel_split_warps(emulator, tmp_path, case, flow_b):
    """Template parameter SPLIT: the X phase and the Y+Z phase on separate warps, coupled only by the full / empty barriers."""
    degree, ttype, r, cells, upper, mask, tile, n_chunks, misalign = case
    space, A, B, src, ref = oracle(degree, ttype, r, cells, upper, mask)
    out, log = run_emulator(emulator, tmp_path, degree, A, B, cells, upper, mask, src, np.full_like(src, 7.5), n_chunks=n_chunks, tile=tile,
                            misalign=misalign, flow_b=flow_b, split=True, pipe=not flow_b)
    assert "split warps" in log
    assert np.abs(out - ref).max() <= 1e-13 * np.abs(ref).max(), log
    assert np.all(out[:, space.constrained] == 0)


def test_emulated_brick_kernel_float_four_row_classes(emulator, tmp_path):
    """FP32: an odd row pitch gives four row classes (shifts 0..3 elements)."""
    degree, cells, upper, mask = 4, [4, 3, 2], [1.0, 1.0, 1.0], 0x3f
    space, A, B, src, ref = oracle(degree, "CGP", 2, cells, upper, mask)
    for mis in (0, 1, 3):
        out, log = run_emulator(emulator, tmp_path, degree, A, B, cells, upper, mask, src, np.zeros_like(src), n_chunks=2, tile=(3, 2),
                                misalign=mis, f32=True)
        assert "4 row classes" in log
        assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max(), log


def test_emulated_brick_kernel_accumulates_and_transposes(emulator, tmp_path):
    """mode 1: dst += A src with constrained rows untouched (vmult_slice_add / residual contract); transposed time matrices."""
    degree, cells, upper, mask = 4, [4, 3, 2], [1.0, 1.0, 1.0], 0x3f
    space, A, B, src, ref = oracle(degree, "CGP", 2, cells, upper, mask, transpose=True)
    dst0 = np.stack([np.random.RandomState(9 + b).uniform(-1, 1, space.n_dofs) for b in range(A.shape[0])])
    out, log = run_emulator(emulator, tmp_path, degree, A.T.copy(), B.T.copy(), cells, upper, mask, src, dst0, mode=1, n_chunks=2, tile=(3, 2))
    assert np.abs(out - (dst0 + ref)).max() <= 1e-13 * np.abs(ref).max(), log
    assert np.all(out[:, space.constrained] == dst0[:, space.constrained])


@pytest.mark.parametrize("mask,tile,cells", [(0x3f, (3, 2), [4, 3, 4]), (0x0c, (3, 2), [5, 2, 3]), (0x00, (7, 4), [8, 5, 2])])
def test_emulated_brick_kernel_rhs_mode(emulator, tmp_path, mask, tile, cells):
    """mode 2: dst = rhs + A src in one pass (the residual b - A x of the multigrid smoothers, include/stmg.h / deal.II
    PreconditionRelaxation), constrained rows = rhs; several z chunks (atomic planes) included."""
    degree, upper = 4, [1.0, 1.0, 1.0]
    space, A, B, src, ref = oracle(degree, "CGP", 2, cells, upper, mask)
    rhs = np.stack([np.random.RandomState(19 + b).uniform(-1, 1, space.n_dofs) for b in range(A.shape[0])])
    out, log = run_emulator(emulator, tmp_path, degree, -A, -B, cells, upper, mask, src, rhs, mode=2, n_chunks=2, tile=tile)
    assert np.abs(out - (rhs - ref)).max() <= 1e-13 * max(np.abs(ref).max(), 1.0), log
    assert np.all(out[:, space.constrained] == rhs[:, space.constrained])


def test_emulated_brick_kernel_z_slabs(emulator, tmp_path):
    """Two launches over z slabs of cells (the pipelined host-buffer entry point): the second accumulates its first plane."""
    degree, cells, upper, mask = 4, [3, 2, 4], [1.0, 1.0, 1.0], 0x3f
    space, A, B, src, ref = oracle(degree, "CGP", 2, cells, upper, mask)
    dst = np.full_like(src, -3.0)
    dst, _ = run_emulator(emulator, tmp_path, degree, A, B, cells, upper, mask, src, dst, zlo=0, zhi=3, tile=(3, 2), n_chunks=2)
    dst, log = run_emulator(emulator, tmp_path, degree, A, B, cells, upper, mask, src, dst, zlo=3, zhi=4, first_plane_acc=1, tile=(3, 2))
    assert np.abs(dst - ref).max() <= 1e-13 * np.abs(ref).max(), log
