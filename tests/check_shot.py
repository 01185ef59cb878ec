"""Oracle half of the short GPU call (see tests/gpu_shot.py): compares gpurun_out/shot_*.npz with the CPU oracle using
the assertions of tests/test_zz_practical_gpu.py.  Prints one line per check; exit code 1 if any failed."""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")

from golden_util import load  # noqa: E402
from oracle import fe_time as ft  # noqa: E402
from oracle import spatial as S  # noqa: E402
from oracle import tp_01  # noqa: E402
from shot_cases import DIAG_CASES, JACOBI_CASES, POINT_CASES, PRACTICAL, practical_2d_case, tp01_params  # noqa: E402

FAILED = []


def check(name, ok, detail=""):
    print("%-52s %s %s" % (name, "ok  " if ok else "FAIL", detail), flush=True)
    if not ok:
        FAILED.append(name)


def get(section):
    f = os.path.join(OUT, "shot_%s.npz" % section)
    if not os.path.exists(f):
        check(section + ": dump present", False, "(section failed or did not run)")
        return None
    return np.load(f)


def compare_run(tag, d, pre, o, k, nts, steps, its_tol, sol_tol=1e-8):
    scale = np.abs(o["x"]).max()
    check(tag + " levels", str(d[pre + "levels"]) == o["levels"], "%s %s" % (d[pre + "levels"], o["levels"]))
    check(tag + " solution", np.abs(d[pre + "x"] - o["x"]).max() <= sol_tol * scale, "%.2e (scale %.3g)" % (np.abs(d[pre + "x"] - o["x"]).max() / scale, scale))
    if o["v"] is not None and pre + "v" in d:
        check(tag + " velocity", np.abs(d[pre + "v"] - o["v"]).max() <= sol_tol * np.abs(o["v"]).max(),
              "%.2e" % (np.abs(d[pre + "v"] - o["v"]).max() / np.abs(o["v"]).max()))
    its, oi = list(d[pre + "its"]), o["iterations_per_solve"]
    check(tag + " iterations", all(abs(a - b) <= its_tol for a, b in zip(its, oi)), "%s vs %s" % (its, oi))
    rows, orows = d[pre + "rows"], np.array(o["functional_rows"])
    if orows.size:
        ok = rows.shape == orows.shape and rows.shape[0] == steps * nts * (k + 1) ** 2
        check(tag + " functional rows shape", ok, "%s %s" % (rows.shape, orows.shape))
        if ok:
            check(tag + " functional times", np.allclose(rows[:, 0], orows[:, 0], rtol=1e-14, atol=0))
            e = np.abs(rows[:, 1:] - orows[:, 1:]).max() / max(np.abs(orows[:, 1:]).max(), 1e-300)
            check(tag + " functional values", e <= 1e-8, "%.2e" % e)


d = get("smoke")
d = get("diag")
if d is not None:
    for i, (dim, degree, distort, ttype, r, coef) in enumerate(DIAG_CASES):
        lo, up = [-1.0] * dim, [1.0] * dim
        sub = [5] * dim if coef else [3] * dim
        mesh = S.Mesh(dim, sub, 1 if coef else 0, lo, up, distort=distort)
        space = S.Space(mesh, degree)
        A, B, _, _ = ft.get_fe_time_weights(ttype, r, 0.05, 1)
        K, M = S.MatrixFreeOperator(space, 0.0, 1.0), S.MatrixFreeOperator(space, 1.0, 0.0)
        if coef:
            K.evaluate_coefficient(S.Coefficient(dim, sub, lo, up, distort_coeff=0.5))
        want = S.SystemMatrix(K, M, A, B).get_matrix_diagonal()
        for nt, tol in ((0, 1e-12), (1, 1e-5)):
            key = "d%d_%d" % (i, nt)
            if key not in d:
                check("diag case %d nt %d present" % (i, nt), False)
                continue
            got = d[key]
            e = np.abs(got - want).max() / np.abs(want).max()
            check("diag case %d %s" % (i, "f64" if nt == 0 else "f32"), e <= tol and np.all(got[:, space.constrained] == 0), "%.2e" % e)

d = get("points")
if d is not None:
    for i, (dim, distort, degree) in enumerate(POINT_CASES):
        lo, up = [-1.0] * dim, [1.0] * dim
        mesh = S.Mesh(dim, [5] * dim, 1, lo, up, distort=distort)
        space = S.Space(mesh, degree)
        u = np.stack([np.random.RandomState(7 + b).uniform(-1, 1, space.n_dofs) for b in range(3)])
        pts = np.array([[0.75, 0.0], [0.013, -0.48], [-1.0, 1.0]] if dim == 2 else
                       [[0.75, 0.0, 0.0], [0.0, 0.0, 0.75], [0.75, 0.1, 0.75], [0.31, -0.77, 0.05], [1.0, 1.0, 1.0]])
        want = tp_01.point_evaluate(space, pts, u)
        e = np.abs(d["p%d" % i] - want).max() / np.abs(want).max()
        check("point evaluation case %d" % i, e <= 1e-13, "%.2e" % e)
        check("point outside rejected case %d" % i, int(d["rc%d" % i]) != 0)

d = get("jacobi")
if d is not None:
    for i, (name, dim, ref, over) in enumerate(JACOBI_CASES):
        pj = dict(tp01_params(name), innerPreconditioner="jacobi", **over)
        p = tp_01.parse_parameters(pj, dim)
        o = tp_01.convergence_test(p, dim, ref, p["feDegree"], mg_dtype=np.float32, max_steps=2, return_state=True)
        pre = "j%d_" % i
        check("jacobi %d no patches" % i, np.all(d[pre + "patches"] == 0))
        check("jacobi %d l2" % i, abs(float(d[pre + "l2"]) - o["l2"]) <= 1e-9 * o["l2"], "%.3e" % (abs(float(d[pre + "l2"]) - o["l2"]) / o["l2"]))
        its, oi = list(d[pre + "its"]), o["iterations_per_solve"]
        check("jacobi %d iterations" % i, all(abs(a - b) <= 1 for a, b in zip(its, oi)), "%s vs %s" % (its, oi))

d = get("practical3d")
if d is not None:
    for problem in ("heat", "wave"):
        pj = dict(PRACTICAL, problemType=problem)
        o = tp_01.convergence_test(tp_01.parse_parameters(pj, 3), 3, 1, 1, mg_dtype=np.float32, max_steps=2, return_state=True)
        compare_run("practical3d " + problem, d, problem + "_", o, 1, 2, 2, 1)

d = get("practical2d")
if d is not None:
    pj, V = practical_2d_case()
    o = tp_01.convergence_test(tp_01.parse_parameters(pj, 2), 2, 2, 2, mg_dtype=np.float32, max_steps=2, return_state=True)
    compare_run("practical2d", d, "", o, 2, 1, 2, 2)

if get("frontend") is not None:
    T = load("tp_01_text")
    got = open(os.path.join(OUT, "shot_frontend.txt")).read().split("\n")
    want = T["tf03"]
    i, j = got.index("Convergence table k=1"), want.index("Convergence table k=1")
    ok = got[i + 1].split() == want[j + 1].split()
    for r in (2, 3):
        g, w = got[i + r].split(), want[j + r].split()
        ok = ok and g[:4] == w[:4] and g[5:] == w[5:]
    check("front end table", ok and got[0] == ":: Number of active cells: 16" and "Iteration count table" in got, repr(got[i + 2]))

if get("facade") is not None:
    txt = open(os.path.join(OUT, "shot_facade.txt")).read()
    n, k = 4, 2
    mesh = S.Mesh(3, [n, n, n], 0)
    space = S.Space(mesh, k)
    A, B, _, _ = ft.get_fe_time_weights("DG", 1, 0.025, 1)
    sysm = S.SystemMatrix(S.MatrixFreeOperator(space, 0.0, 1.0), S.MatrixFreeOperator(space, 1.0, 0.0), A, B)
    md = re.search(r"diagonal_sum (\S+)", txt)
    dsum = sysm.get_matrix_diagonal().sum()
    check("facade diagonal_sum", bool(md) and abs(float(md.group(1)) - dsum) <= 1e-11 * abs(dsum), md.group(1) if md else txt[-300:])
    check("facade rest", "non-square vmult rejected: yes" in txt)

d = get("half") if os.path.exists(os.path.join(OUT, "shot_half.npz")) else None
if d is not None:
    o = tp_01.convergence_test(tp_01.parse_parameters(dict(PRACTICAL, problemType="heat"), 3), 3, 1, 1, mg_dtype=np.float32,
                               max_steps=2, return_state=True)
    scale = np.abs(o["x"]).max()
    e = np.abs(d["half_vanka"].astype(np.float64) - d["level_vanka"]).max() / np.abs(d["level_vanka"]).max()
    check("half: one Vanka application vs float storage", e <= 2e-3, "%.2e" % e)
    check("half: patch bytes halved", np.all(d["half_bytes"] <= 0.55 * d["level_bytes"]), "%s %s" % (d["half_bytes"], d["level_bytes"]))
    check("half: iterations vs float storage", all(abs(a - b) <= 1 for a, b in zip(d["half_its"], d["level_its"])),
          "%s vs %s (oracle %s)" % (list(d["half_its"]), list(d["level_its"]), o["iterations_per_solve"]))
    e = np.abs(d["half_x"] - o["x"]).max() / scale
    check("half: solution vs oracle", e <= 1e-8, "%.2e" % e)
d = np.load(os.path.join(OUT, "shot_c4time.npz")) if os.path.exists(os.path.join(OUT, "shot_c4time.npz")) else None
if d is not None:
    for storage in ("level", "half"):
        setup_s, ms, bytes_, s0, s1, i0, i1, ndof = d[storage]
        print("c4 %-5s: Vanka apply %.3f ms, %.1f MB of inverses -> %.0f GB/s; steps %.1f / %.1f ms, iterations %d / %d, %.3e st-DoF/s"
              % (storage, ms, bytes_ / 1e6, bytes_ / ms / 1e6, s0, s1, i0, i1, ndof / (s1 * 1e-3)))

print("FAILED: %s" % FAILED if FAILED else "all checks passed")
sys.exit(1 if FAILED else 0)
