"""Product-side half of tests/test_zz_practical_gpu.py, for a SHORT GPU call: runs only the library (no oracle), dumps
every result to gpurun_out/shot_<section>.npz; scripts/check_shot.py compares the dumps with the oracle on the CPU.

    gpurun --timeout 60 -- 'timeout 55 python tests/gpu_shot.py > gpurun_out/shot.log 2>&1'
    python tests/check_shot.py

(Written when the round's GPU budget was down to two minutes: the oracle halves of the tests take longer than that.)"""
import ctypes as C
import io
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))           # shot_cases.py (inputs shared with check_shot.py)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

import dealii_stfem_b200 as st  # noqa: E402
from dealii_stfem_b200 import fe_time_host as fth  # noqa: E402
from dealii_stfem_b200 import tp_01 as front  # noqa: E402
from shot_cases import (DIAG_CASES, JACOBI_CASES, POINT_CASES, PRACTICAL, box_mesh_vertices, practical_2d_case,  # noqa: E402
                        tp01_params)

T0 = time.time()


def log(*a):
    print("[%.1fs]" % (time.time() - T0), *a, flush=True)


def section(name):
    def deco(fn):
        try:
            t = time.time()
            data = fn()
            np.savez(os.path.join(OUT, "shot_%s.npz" % name), **data)
            log("section", name, "ok in %.1fs" % (time.time() - t))
        except Exception:
            log("section", name, "FAILED")
            traceback.print_exc()
            sys.stdout.flush()
        return fn
    return deco


ctx = st.Context(0)
only = set(sys.argv[1:])


def want(name):
    return not only or name in only


def run_problem(pj, dim, ref, k, vertices=None, steps=2):
    p = st.parse_parameters(pj, dim)
    prob = st.HeatWaveProblem(ctx, p, dim, ref, k, vertices_fn=(lambda n: vertices) if vertices is not None else None)
    its = [prob.step() for _ in range(steps)]
    d = dict(its=np.array(its), x=prob.x.download(), rows=np.array(prob.functional_rows), levels=np.array("".join(prob.mg_type_level)),
             l2=np.array(prob.row()["l2"]), patches=np.array([prob.mg.level_info(l)["patch_matrices"] for l in range(prob.mg.n_levels)]))
    if prob.wave:
        d["v"] = prob.v.download()
    prob.close()
    return d


if want("smoke"):
    @section("smoke")
    def _():
        import __graft_entry__ as g
        g.smoke()
        return {"ok": np.array(1)}

if want("diag"):
    @section("diag")
    def _():
        out = {}
        for i, (dim, degree, distort, ttype, r, coef) in enumerate(DIAG_CASES):
            lo, up = [-1.0] * dim, [1.0] * dim
            sub = [5] * dim if coef else [3] * dim
            ref = 1 if coef else 0
            n = [s << ref for s in sub]
            V = box_mesh_vertices(dim, sub, ref, lo, up, distort)
            A, B, _, _ = fth.get_fe_time_weights(ttype, r, 0.05, 1)
            kw = {}
            if coef:
                cq = st.problem_host.coefficient_at_qpoints(n, lo, up, degree, sub, lo, up, 0.5, vertices=None if V is None else V.reshape(-1, dim))
                kw = {"laplace_coeff_q": cq} if coef == "q" else {"laplace_coeff_cell": np.ascontiguousarray(cq[:, 0])}
            gm = st.Mesh(ctx, n, lower=lo, upper=up, vertices=None if V is None else V.reshape(-1, dim))
            for nt in (st.F64, st.F32):
                op = st.Operator(gm, degree, A, B, number_type=nt, **kw)
                d = op.new_vector()
                op.diagonal(d)
                out["d%d_%d" % (i, nt)] = d.download()
                d.free(); op.close()
            gm.close()
            log("diag case", i, "done")
        return out

if want("points"):
    @section("points")
    def _():
        out = {}
        for i, (dim, distort, degree) in enumerate(POINT_CASES):
            lo, up = [-1.0] * dim, [1.0] * dim
            n = [10] * dim
            V = box_mesh_vertices(dim, [5] * dim, 1, lo, up, distort)
            ndofs = int(np.prod([degree * m + 1 for m in n]))
            nb = 3
            u = np.stack([np.random.RandomState(7 + b).uniform(-1, 1, ndofs) for b in range(nb)])
            pts = np.array([[0.75, 0.0], [0.013, -0.48], [-1.0, 1.0]] if dim == 2 else
                           [[0.75, 0.0, 0.0], [0.0, 0.0, 0.75], [0.75, 0.1, 0.75], [0.31, -0.77, 0.05], [1.0, 1.0, 1.0]])
            gm = st.Mesh(ctx, n, lower=lo, upper=up, vertices=None if V is None else V.reshape(-1, dim))
            dv = st.DeviceBlockVector(ctx, nb, ndofs, st.F64).upload(u)
            got = np.zeros((nb, len(pts)))
            ptrs = (C.c_void_p * nb)(*[dv.ptrs[b] for b in range(nb)])
            st.capi.check(st.capi.lib().stfem_point_evaluate(gm.h, degree, len(pts), st.capi._dptr(np.ascontiguousarray(pts)), nb, ptrs,
                                                             st.capi._dptr(got)))
            bad = np.ascontiguousarray([[2.0] * dim])
            rc = st.capi.lib().stfem_point_evaluate(gm.h, degree, 1, st.capi._dptr(bad), nb, ptrs, st.capi._dptr(np.zeros((nb, 1))))
            out["p%d" % i] = got
            out["rc%d" % i] = np.array(rc)
            dv.free(); gm.close()
        return out

if want("jacobi"):
    @section("jacobi")
    def _():
        out = {}
        for i, (name, dim, ref, over) in enumerate(JACOBI_CASES):
            pj = dict(tp01_params(name), innerPreconditioner="jacobi", **over)
            p = st.parse_parameters(pj, dim)
            d = run_problem(pj, dim, ref, p["feDegree"])
            for k_, v_ in d.items():
                out["j%d_%s" % (i, k_)] = v_
            log("jacobi case", i, d["its"])
        return out

if want("practical3d"):
    @section("practical3d")
    def _():
        out = {}
        for problem in ("heat", "wave"):
            d = run_problem(dict(PRACTICAL, problemType=problem), 3, 1, 1)
            for k_, v_ in d.items():
                out["%s_%s" % (problem, k_)] = v_
            log("practical3d", problem, d["its"])
        return out

if want("practical2d"):
    @section("practical2d")
    def _():
        pj, V = practical_2d_case()
        d = run_problem(pj, 2, 2, 2, vertices=V)
        log("practical2d", d["its"])
        return d

if want("frontend"):
    @section("frontend")
    def _():
        pj = dict(tp01_params("tf03"), nDegCycles="1", nRefCycles="2")
        out = io.StringIO()
        front.run(pj, 2, out=out, ctx=ctx)
        with open(os.path.join(OUT, "shot_frontend.txt"), "w") as f:
            f.write(out.getvalue())
        return {"ok": np.array(1)}

if want("facade"):
    @section("facade")
    def _():
        import subprocess
        exe = os.path.join(OUT, "facade_demo")
        lib = os.path.join(ROOT, "dealii-stfem_b200")
        r = subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "facade_demo.cpp"),
                            "-o", exe, "-L", lib, "-lstfem_b200", "-Wl,-rpath," + lib], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        r = subprocess.run([exe, os.path.join(OUT, "facade_y.bin"), "4", "2"], capture_output=True, text=True, timeout=60)
        with open(os.path.join(OUT, "shot_facade.txt"), "w") as f:
            f.write(r.stdout + r.stderr)
        os.remove(exe)
        assert r.returncode == 0, r.stdout + r.stderr
        return {"ok": np.array(1)}

if want("half"):
    @section("half")
    def _():
        # FP16 storage of the dense patch inverses against the level-precision storage: one Vanka application on the
        # finest level (random input) and the first two solves of practical01 (3D heat)
        out = {}
        for storage in ("level", "half"):
            pj = dict(PRACTICAL, problemType="heat", vankaStorage=storage)
            p = st.parse_parameters(pj, 3)
            prob = st.HeatWaveProblem(ctx, p, 3, 1, 1)
            lop = prob.level_ops[-1]
            xin = np.stack([np.random.RandomState(11 + b).uniform(-1, 1, lop.n) for b in range(lop.nb_rows)]).astype(np.float32)
            dx, dy = lop.new_vector().upload(xin), lop.new_vector()
            prob.mg.level_apply(prob.mg.n_levels - 1, 0, dy, dx)
            out[storage + "_vanka"] = dy.download()
            out[storage + "_bytes"] = np.array([prob.mg.level_info(l)["patch_bytes"] for l in range(prob.mg.n_levels)])
            dx.free(); dy.free()
            out[storage + "_its"] = np.array([prob.step() for _ in range(2)])
            out[storage + "_x"] = prob.x.download()
            prob.close()
            log("half section:", storage, out[storage + "_its"])
        return out

if want("c4time"):
    @section("c4time")
    def _():
        # BASELINE configs[3], practical set-up (scripts/solve_c4.py), refinement 3: time of one dense Vanka application on
        # the finest level (CUDA events on the library stream, 20 launches) and of the first two time steps
        out = {}
        ref, k, r = 3, 3, 2

        def vertices(n_cells):
            n = [c + 1 for c in n_cells]
            g = [np.linspace(-1.0, 1.0, m) for m in n]
            V = np.stack(np.meshgrid(g[2], g[1], g[0], indexing="ij")[::-1], axis=-1)
            d = np.random.RandomState(1).uniform(-1, 1, V.shape) * 0.15 * (2.0 / n_cells[0])
            d[0] = d[-1] = 0
            d[:, 0] = d[:, -1] = 0
            d[:, :, 0] = d[:, :, -1] = 0
            return V + d

        V = vertices([5 << ref] * 3).reshape(-1, 3)
        src_pt = [float(c) for c in V[np.argmin(np.sum(V * V, axis=1))]]
        for storage in ("level", "half"):
            pj = {"timeType": "DG", "problemType": "heat", "feDegree": r, "refinement": ref, "subdivisions": "5,5,5",
                  "hyperRectLowerLeft": "-1,-1,-1", "hyperRectUpperRight": "1,1,1", "mgTimeBeforeSpace": "true",
                  "spaceTimeConvergenceTest": "false", "distortGrid": 0.15, "distortCoeff": "0.6", "extrapolate": "false",
                  "vankaStorage": storage}
            p = st.parse_parameters(pj, 3)
            p["sourcePoint"] = src_pt
            t = time.time()
            prob = st.HeatWaveProblem(ctx, p, 3, ref, r, space_degree=k, vertices_fn=vertices)
            ctx.synchronize()
            setup_s = time.time() - t
            l = prob.mg.n_levels - 1
            lop = prob.level_ops[-1]
            xin = np.stack([np.random.RandomState(11 + b).uniform(-1, 1, lop.n) for b in range(lop.nb_rows)]).astype(np.float32)
            dx, dy = lop.new_vector().upload(xin), lop.new_vector()
            for _ in range(3):
                prob.mg.level_apply(l, 0, dy, dx)
            ctx.synchronize()
            ctx.timer_start()
            for _ in range(20):
                prob.mg.level_apply(l, 0, dy, dx)
            ms = ctx.timer_stop() / 20.0
            dx.free(); dy.free()
            steps = []
            its = []
            for _ in range(2):
                ctx.synchronize()
                t = time.perf_counter()
                its.append(prob.step(evaluate_error=False))
                ctx.synchronize()
                steps.append((time.perf_counter() - t) * 1e3)
            bytes_ = prob.mg.level_info(l)["patch_bytes"]
            out[storage] = np.array([setup_s, ms, bytes_, steps[0], steps[1], its[0], its[1], prob.n * prob.nb])
            log("c4time %s: setup %.2fs, level_apply(Vanka) %.3f ms for %.1f MB of inverses (%.0f GB/s incl. 2 vector copies + memset), "
                "steps %.1f / %.1f ms, iterations %s" % (storage, setup_s, ms, bytes_ / 1e6, bytes_ / ms / 1e6, steps[0], steps[1], its))
            prob.close()
        return out

ctx.close()
log("done")
