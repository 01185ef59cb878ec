#!/usr/bin/env python3
"""bench.py — headline benchmark of the space-time operator hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c5]

A "step" is one application of the fused space-time operator  dst = (Alpha (x) K + Beta (x) M) src
(reference SystemMatrix::vmult, include/operators.h:536-559) over the whole block vector.
Workload at N=1 = BASELINE.json configs[1] ("c2"): 3D heat, Q4 space x cG(2) time, 96^3 cells
(subdivisions 3, refinement 5), 385^3 spatial DoFs x 2 time blocks = 1.14e8 space-time DoFs, FP64.
For N>1 every rank owns one such brick (weak scaling, box partition of the mesh).
`--config c5` times configs[4] instead: Q4 x DG(1) on a 228^3-cell brick per GPU (1.5e9 space-time DoFs per GPU).

Prints ONE JSON line (rank 0).  `value` = space-time DoFs/s with inputs resident in HBM (CUDA events on
the library's stream, max over ranks); `e2e` = the same metric through the C-ABI entry point with
pinned HOST buffers (H2D + kernel + D2H inside the timed region); `roofline`, `cpu_baseline`,
`clocks`, `gpu_launches` as the contract asks.  torch is used for torch.distributed plumbing only (rendezvous and
the broadcast of the NCCL unique id); every reduction of the timed path goes through the library's own communicator.
Extra objects on the same line: `solve` (STMG-FGMRES time steps of configs[1]), `weak_c5` (configs[4] operator on a
228^3-cell brick per GPU), `parity_multi_gpu` (N>1: partitioned run against the single-GPU run of the same global problem);
at N=1 also `perturbed_mesh` (general-geometry kernel), `wave_c3` (configs[2]: 3D wave, Q3 x cG(2), 208^3 cells) and
`practical_c4` (configs[3]: heterogeneous coefficient, perturbed mesh, dense cell-patch smoother).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "space-time DoFs/s, operator vmult (3D heat, Q4 x cG(2), FP64)"
UNIT = "DoFs/s"
BYTES_PER_DOF = 16           # FP64: read src once + write dst once
# the two operator workloads: (degree, time type, time degree, cells per direction and GPU)
CONFIGS = {"c2": dict(degree=4, ttype="CGP", tdeg=2, cells=96, name="configs[1]: 3D heat, Q4 x cG(2)"),
           "c5": dict(degree=4, ttype="DG", tdeg=1, cells=228, name="configs[4]: 3D heat, Q4 x DG(1)")}


def time_weights(cfg):
    """Alpha, Beta for tau = 2^-6 (tests/tp_01.cc:106-109 with refinement 5)."""
    from dealii_stfem_b200 import fe_time_host
    return fe_time_host.get_fe_time_weights(cfg["ttype"], cfg["tdeg"], 2.0 ** -6, 1)[:2]


def workload_config(cfg, n, world, grid):
    """The `config` object: identical in both arms (the driver compares them)."""
    N = (cfg["degree"] * n + 1) ** 3
    nb = 2
    return {"workload": "%s, %d^3 cells per GPU, %d spatial DoFs x %d time blocks = %.4g space-time DoFs per GPU; one step = "
                        "one fused operator vmult" % (cfg["name"], n, N, nb, N * nb),
            "l2": "inputs+outputs %.0f MB per step, larger than the 126 MB L2" % (2 * N * nb * 8 / 1e6),
            "parallelism": "box partition %s, one brick per GPU" % "x".join(map(str, grid))}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md 'clocks' recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smmax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smmax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1: not what the CPU arm is about)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_run(cfg, n_cells, steps, warmup):
    """The reference's CPU algorithm (unfused 2*nb cell loops + axpys, operators.h:536-559) through
    oracle/cpu_ref.cpp on all host threads.  Returns (DoFs/s, ms per step, threads, DoFs)."""
    from oracle import cpu_ref, fe_time as oft, spatial as osp
    A, B = oft.get_fe_time_weights(cfg["ttype"], cfg["tdeg"], 2.0 ** -6, 1)[:2]
    mesh = osp.Mesh(3, [n_cells] * 3, 0)
    space = osp.Space(mesh, cfg["degree"])
    nb = A.shape[0]
    src = np.sin(0.1 * np.arange(space.n_dofs)[None, :] + np.arange(nb)[:, None])
    threads = host_threads()
    for _ in range(warmup):
        cpu_ref.system_vmult(space, A, B, src, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_ref.system_vmult(space, A, B, src, n_threads=threads)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return nb * space.n_dofs / dt, dt * 1e3, threads, nb * space.n_dofs


def bind_host_to_gpu(dev):
    """CPU affinity of this process := the cores NVML reports as local to CUDA device `dev` (matched by PCI bus id)."""
    try:
        import pynvml
        import torch
        pr = torch.cuda.get_device_properties(dev)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "unchanged (no local cores reported)"
        os.sched_setaffinity(0, cpus)
        return "%d cores local to GPU %s" % (len(cpus), bus)
    except Exception as e:            # measurement nicety only: never fail the bench over it
        return "unchanged (%s)" % type(e).__name__


def ncu_traffic(cfg_name, n_cells):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the headline kernel, read from the committed ncu capture
    of the same command (profiles/): (bytes, provenance) - not measured in this run."""
    for name in ("r02_st_vmult_brick_ncu.json", "r01_st_vmult_cart_ncu.json"):
        p = os.path.join(ROOT, "profiles", name)
        if cfg_name == "c2" and n_cells == CONFIGS["c2"]["cells"] and os.path.exists(p):
            d = json.load(open(p))
            return d.get("dram_bytes_per_launch"), "profiles/%s (%s; ncu --set full capture of the same command, not this run)" % (
                name, d.get("kernel", "kernel"))
    return None, None


def perturbed_leg(st, ctx, cfg, n, steps=20):
    """The same operator on a randomly perturbed mesh (GridTools::distort_random-style, factor 0.15): general-geometry
    kernel (csrc/st_vmult_plane.cuh).  Algorithmic bytes: 16 B per space-time DoF + 200 B per cell (8 vertices
    + coefficient); the precomputed metric the kernel actually streams is reported separately."""
    A, B = time_weights(cfg)
    n1 = n + 1
    g = np.linspace(0.0, 1.0, n1)
    V = np.stack(np.meshgrid(g, g, g, indexing="ij")[::-1], axis=-1)
    d = np.random.RandomState(1).uniform(-1, 1, V.shape) * 0.15 / n
    d[0] = d[-1] = 0
    d[:, 0] = d[:, -1] = 0
    d[:, :, 0] = d[:, :, -1] = 0
    mesh = st.Mesh(ctx, [n, n, n], vertices=(V + d).reshape(-1, 3))
    peak, _ = measured_peak()
    out = {"metric": "space-time DoFs/s, operator vmult on a perturbed mesh (MappingQ1 cells), FP64", "unit": UNIT}
    # default = stored metric where it fits (this mesh: 7 GB); kernel_variant 6 = geometry on the fly from the cell vertices
    for key, variant in (("stored_metric", 0), ("on_the_fly", 6)):
        op = st.Operator(mesh, cfg["degree"], A, B, number_type=st.F64, variant=variant)
        nb = op.nb_rows
        x, y = op.new_vector(), op.new_vector()
        x.upload(np.sin(0.1 * np.arange(op.n)[None, :] + np.arange(nb)[:, None]))
        for _ in range(3):
            op.vmult(y, x)
        ctx.timer_start()
        for _ in range(steps):
            op.vmult(y, x)
        ms = ctx.timer_stop() / steps
        dofs = op.n * nb
        alg = dofs * BYTES_PER_DOF + n ** 3 * 200
        leg = {"value": dofs / (ms * 1e-3), "ms_per_step": ms, "achieved_GBs": alg / (ms * 1e-3) / 1e9,
               "frac": alg / (ms * 1e-3) / 1e9 / peak}
        if key == "stored_metric":
            out.update(leg)
            out.update({"cells": n ** 3, "distortion": 0.15, "algorithmic_bytes_per_launch": alg,
                        "streamed_metric_bytes": n ** 3 * (cfg["degree"] + 1) ** 3 * 8 * 8,
                        "note": "algorithmic = 16 B/DoF + 200 B/cell (SURVEY 8d); value = the default kernel (stored metric); the "
                                "kernel is bound by the FP64 pipe (1.4 ms at the FP64 peak, stored; 2.1 ms on the fly), DESIGN.md 3.2"})
        else:
            out[key] = leg
        x.free(); y.free(); op.close()
    mesh.close()
    return out


def weak_c5_leg(st, ctx, mesh_factory, steps=10):
    """BASELINE configs[4]: Q4 x DG(1) on a 228^3-cell brick per GPU (1.52e9 space-time DoFs, 24 GB of vectors per GPU),
    halo exchange over NVLink inside every step when partitioned.  The source vector is filled on the device
    (interpolated analytic function), no host copy."""
    cfg = CONFIGS["c5"]
    A, B = time_weights(cfg)
    mesh = mesh_factory(cfg["cells"])
    op = st.Operator(mesh, cfg["degree"], A, B, number_type=st.F64)
    nb = op.nb_rows
    x, y = op.new_vector(), op.new_vector()
    for b in range(nb):
        st.capi.check(st.capi.lib().stfem_interpolate(mesh.h, cfg["degree"], 1, 1.0, 0.1 * (b + 1), x.ptrs[b]))
    for _ in range(3):
        op.vmult(y, x)
    ctx.synchronize()
    ctx.timer_start()
    for _ in range(steps):
        op.vmult(y, x)
    ms = ctx.timer_stop() / steps
    ms = float(st.dist.allreduce(ctx, [ms], "max")[0])
    dofs = op.n * nb
    world = st.capi.lib().stfem_ctx_n_ranks(ctx.h)
    peak, _ = measured_peak()
    out = {"metric": "space-time DoFs/s, operator vmult, configs[4] (3D heat, Q4 x DG(1), 228^3 cells per GPU, FP64)",
           "value": dofs * world / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "st_dofs_per_gpu": dofs, "n_gpus": world,
           "achieved_GBs_per_gpu": dofs * BYTES_PER_DOF / (ms * 1e-3) / 1e9, "frac": dofs * BYTES_PER_DOF / (ms * 1e-3) / 1e9 / peak}
    x.free(); y.free(); op.close(); mesh.close()
    return out


def solve_leg(st, ctx, refinement, n_steps=3, grid=None, coords=None):
    """Full STMG-preconditioned FGMRES time steps of configs[1] (3D heat, Q4 x cG(2), float multigrid) through the
    product driver: rhs assembly + solve per step, all on the device (tests/tp_01.cc:646-669)."""
    cfg = CONFIGS["c2"]
    grid = grid or [1, 1, 1]
    world = int(np.prod(grid))
    pj = {"timeType": cfg["ttype"], "problemType": "heat", "feDegree": cfg["tdeg"], "refinement": refinement,
          "subdivisions": ",".join(str(3 * g) for g in grid), "hyperRectUpperRight": ",".join(str(float(g)) for g in grid),
          "mgTimeBeforeSpace": "true", "smoother": "relaxation", "spaceTimeConvergenceTest": "true",
          "agglomerateBelow": os.environ.get("STFEM_AGGLO", "16"),
          "levelKernelVariant": os.environ.get("STFEM_LEVEL_VARIANT", "0")}
    p = st.parse_parameters(pj, 3)
    t0 = time.perf_counter()
    prob = st.HeatWaveProblem(ctx, p, 3, refinement, cfg["tdeg"], space_degree=cfg["degree"],
                              partition=(grid, coords) if world > 1 else None)
    ctx.synchronize()
    setup_s = time.perf_counter() - t0
    its = [prob.step(evaluate_error=False)]         # warm-up step (graph capture, lazy allocations)
    ctx.synchronize()
    l0 = ctx.launches
    ctx.timer_start()
    for _ in range(n_steps):
        its.append(prob.step(evaluate_error=False))
    ms = ctx.timer_stop()
    ms = float(st.dist.allreduce(ctx, [ms], "max")[0])
    launches = ctx.launches - l0
    dofs = prob.n * prob.nb * world          # interface DoFs counted on every rank owning a copy (< 1 %)
    out = {"metric": "space-time DoFs/s, STMG-FGMRES solve (3D heat, Q4 x cG(2), FP64 outer / FP32 multigrid)",
           "value": dofs * n_steps / (ms * 1e-3), "unit": UNIT, "ms_per_solve": ms / n_steps, "timesteps": n_steps,
           "fgmres_iterations_per_solve": its[1:], "st_dofs": dofs, "levels": "".join(prob.mg_type_level),
           "work_per_s": dofs * sum(its[1:]) / (ms * 1e-3), "setup_s": setup_s, "gpu_launches": int(launches),
           "config": "subdivisions 3,3,3 per GPU brick, refinement %d, tau %.4g, relaxation smoother around cell-patch Vanka "
                     "(Kronecker form), variable V-cycle captured as a CUDA graph, reduce 1e-12, partition %s" %
                     (refinement, prob.tau, "x".join(map(str, grid)))}
    prob.close()
    return out


def wave_leg(st, ctx, n_steps=2):
    """BASELINE configs[2]: 3D wave equation, Q3 x cG(2), subdivisions 13 with 4 refinements = 208^3 cells, 625^3 spatial
    DoFs x 2 = 4.88e8 space-time DoFs; operator Alpha (x) K + (B A^-1 B) (x) M (tests/tp_01.cc:142-159), full time steps."""
    pj = {"timeType": "CGP", "problemType": "wave", "feDegree": 2, "refinement": 4, "subdivisions": "13,13,13",
          "mgTimeBeforeSpace": "true", "smoother": "relaxation", "spaceTimeConvergenceTest": "true"}
    p = st.parse_parameters(pj, 3)
    t0 = time.perf_counter()
    prob = st.HeatWaveProblem(ctx, p, 3, 4, 2, space_degree=3)
    ctx.synchronize()
    setup_s = time.perf_counter() - t0
    its = [prob.step(evaluate_error=False)]
    ctx.synchronize()
    ctx.timer_start()
    for _ in range(n_steps):
        its.append(prob.step(evaluate_error=False))
    ms = ctx.timer_stop() / n_steps
    dofs = prob.n * prob.nb
    out = {"metric": "space-time DoFs/s, STMG-FGMRES solve of configs[2] (3D wave, Q3 x cG(2), 208^3 cells)", "unit": UNIT,
           "value": dofs / (ms * 1e-3), "ms_per_solve": ms, "fgmres_iterations_per_solve": its[1:], "st_dofs": dofs,
           "levels": "".join(prob.mg_type_level), "setup_s": setup_s}
    prob.close()
    return out


def practical_leg(st, ctx, refinement=3, n_steps=2):
    """BASELINE configs[3] in the reference's practical set-up (tests/json/practical01.json + run_practical.sh): 3D heat,
    Q3 x DG(2), box [-1,1]^3 with 5 subdivisions, perturbed mesh (0.15), Coefficient<dim> table (distortCoeff 0.6) on K on
    every level, cut-off initial value, zero source; relaxation smoother around the DENSE cell-patch Vanka (one inverse
    per cell).  Reported for the patch inverses in the level precision (float, what the reference stores) and in FP16."""
    k, r = 3, 2

    def vertices(n_cells):
        n = [c + 1 for c in n_cells]
        g = [np.linspace(-1.0, 1.0, m) for m in n]
        V = np.stack(np.meshgrid(g[2], g[1], g[0], indexing="ij")[::-1], axis=-1)
        d = np.random.RandomState(1).uniform(-1, 1, V.shape) * 0.15 * (2.0 / n_cells[0])
        d[0] = d[-1] = 0
        d[:, 0] = d[:, -1] = 0
        d[:, :, 0] = d[:, :, -1] = 0
        return V + d

    V = vertices([5 << refinement] * 3).reshape(-1, 3)
    source = [float(c) for c in V[np.argmin(np.sum(V * V, axis=1))]]     # the displaced vertex next to the origin
    out = {"metric": "space-time DoFs/s, STMG-FGMRES solve of configs[3] (3D heat, Q3 x DG(2), perturbed mesh, heterogeneous "
                     "coefficient, dense cell-patch smoother)", "unit": UNIT}
    for storage in ("level", "half"):
        pj = {"timeType": "DG", "problemType": "heat", "feDegree": r, "refinement": refinement, "subdivisions": "5,5,5",
              "hyperRectLowerLeft": "-1,-1,-1", "hyperRectUpperRight": "1,1,1", "mgTimeBeforeSpace": "true",
              "spaceTimeConvergenceTest": "false", "distortGrid": 0.15, "distortCoeff": "0.6", "extrapolate": "false",
              "vankaStorage": storage}
        p = st.parse_parameters(pj, 3)
        p["sourcePoint"] = source
        t0 = time.perf_counter()
        prob = st.HeatWaveProblem(ctx, p, 3, refinement, r, space_degree=k, vertices_fn=vertices)
        ctx.synchronize()
        setup_s = time.perf_counter() - t0
        level = prob.mg.n_levels - 1
        lop = prob.level_ops[-1]
        dx, dy = lop.new_vector(), lop.new_vector()
        dx.upload(np.sin(0.1 * np.arange(lop.n)[None, :] + np.arange(lop.nb_rows)[:, None]).astype(np.float32))
        for _ in range(3):
            prob.mg.level_apply(level, 0, dy, dx)
        ctx.timer_start()
        for _ in range(10):
            prob.mg.level_apply(level, 0, dy, dx)
        vanka_ms = ctx.timer_stop() / 10.0
        dx.free(); dy.free()
        its = [prob.step(evaluate_error=False)]        # warm-up step
        ctx.synchronize()
        ctx.timer_start()
        for _ in range(n_steps):
            its.append(prob.step(evaluate_error=False))
        ms = ctx.timer_stop() / n_steps
        patch_bytes = prob.mg.level_info(level)["patch_bytes"]
        dofs = prob.n * prob.nb
        out[storage] = {"value": dofs / (ms * 1e-3), "ms_per_solve": ms, "fgmres_iterations_per_solve": its[1:], "st_dofs": dofs,
                        "cells": int(np.prod(prob.n_cells)), "levels": "".join(prob.mg_type_level), "setup_s": setup_s,
                        "vanka_apply_ms": vanka_ms, "patch_inverse_bytes": patch_bytes,
                        "vanka_apply_GBs": patch_bytes / (vanka_ms * 1e-3) / 1e9}
        prob.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="c2 = BASELINE configs[1] (default), c5 = configs[4]")
    ap.add_argument("--cells", type=int, default=0, help="cells per direction per GPU (default: the configuration's)")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-solve", action="store_true", help="skip the STMG-FGMRES solve legs")
    ap.add_argument("--no-perturbed", action="store_true", help="skip the perturbed-mesh operator leg")
    ap.add_argument("--no-practical", action="store_true", help="skip the configs[3] (practical set-up, dense Vanka) leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[2] wave leg and the configs[4] weak-scaling leg")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the partitioned-vs-global parity check")
    ap.add_argument("--solve-refinement", type=int, default=5, help="solve leg: subdivisions 3, this many refinements (5 = 96^3 cells)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = CONFIGS[args.config]
    n = args.cells or cfg["cells"]
    metric = METRIC if args.config == "c2" else "space-time DoFs/s, operator vmult (3D heat, Q4 x DG(1), FP64)"

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only, N=1 semantics)
    if args.impl == "reference":
        if rank != 0:
            return 0
        # the SAME brick as our arm; a step = one application on all host threads (96^3 cells: about 1 s per step)
        val, ms, threads, ndofs = cpu_reference_run(cfg, n, args.steps, args.warmup)
        sample = ("the full %d^3-cell brick of one GPU (%.4g space-time DoFs) per step, %d host threads (all this process may use), "
                  "oracle/cpu_ref.cpp = C++/OpenMP port of the reference's unfused algorithm (the reference needs deal.II and "
                  "cannot be built here); scalar per cell (deal.II vectorises over cells; a cell-batched SIMD variant of this "
                  "port was measured and is not faster on this host: the unfused algorithm is bound by its vector traffic)" % (n, ndofs, threads))
        line = {"impl": "reference", "metric": metric, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(cfg, n, 1, [1, 1, 1]),
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    import dealii_stfem_b200 as st

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)

    # N > 1: every rank stays on the cores next to its GPU, so that its pinned buffers come from the NUMA node the GPU's PCIe
    # root belongs to (torchrun does not bind; unbound, eight ranks crossing the socket interconnect share its bandwidth)
    host_binding = bind_host_to_gpu(dev) if world > 1 else "none (single rank)"

    ctx = st.Context(dev)

    def barrier():
        """Rank barrier through the LIBRARY's communicator (one small all-reduce on the context stream): torch's NCCL
        communicator is never used while the library has work in flight."""
        ctx.synchronize()
        if world > 1:
            st.dist.allreduce(ctx, [0.0], "max")

    A, B = time_weights(cfg)
    grid, coords = [1, 1, 1], [0, 0, 0]
    if world > 1:
        # box partition 2x1x1, 2x2x1, 2x2x2: every rank owns one n^3 brick (unit cube) of the global mesh; the operator
        # sums interface DoFs over NVLink (peer-memory stores, NCCL fallback) inside every vmult
        grid = st.dist.proc_grid_for(world, 3)
        coords = st.dist.coords_of(rank, grid)

        def bcast(b):
            t = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                t = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
            dist.broadcast(t, 0)
            return bytes(t.cpu().numpy().tobytes())

        st.dist.init_comm(ctx, rank, world, bcast)
        torch.cuda.synchronize()

    def make_mesh(cells):
        if world == 1:
            return st.Mesh(ctx, [cells] * 3)
        n_loc, _, llo, lup, mask = st.dist.partition_brick([cells * g for g in grid], [0.0] * 3, [float(g) for g in grid], grid, coords)
        m = st.Mesh(ctx, n_loc, lower=llo, upper=lup, dirichlet_faces=mask)
        st.dist.set_partition(m, grid, coords)
        return m

    parity = None
    if world > 1 and not args.no_parity:
        parity = st.dist.parity_check(ctx, dev, rank, world)

    mesh = make_mesh(n)
    op = st.Operator(mesh, cfg["degree"], A, B, number_type=st.F64, variant=args.variant)
    nb = op.nb_rows
    dofs_rank = op.n * nb
    x, y = op.new_vector(), op.new_vector()
    hx, px = st.capi.pinned_array((nb, op.n), np.float64)
    hy, py = st.capi.pinned_array((nb, op.n), np.float64)
    hx[:] = np.sin(0.1 * np.arange(op.n)[None, :] + np.arange(nb)[:, None])      # SURVEY §8d synthetic input
    x.upload(hx)

    # device-resident timing; clocks are sampled from the warm-up to the end of the soak loop (all the same kernel under
    # load) because the timed region itself is shorter than nvidia-smi's sampling period.  Every loop below has the SAME
    # trip count on every rank: each vmult contains matched send/receive pairs.
    sampler = ClockSampler(dev)
    sampler.start()
    for _ in range(warmup):
        op.vmult(y, x)
    barrier()
    launches0 = ctx.launches
    ctx.timer_start()
    for _ in range(args.steps):
        op.vmult(y, x)
    ms_total = ctx.timer_stop()
    launches = ctx.launches - launches0
    barrier()
    # kernel-only duration (CUDA events around each launch, same stream), for the roofline
    op.set_timing(True)
    kms = []
    for _ in range(min(args.steps, 10)):
        op.vmult(y, x)
        kms.append(op.last_kernel_ms())
    op.set_timing(False)
    kernel_ms = float(np.mean(kms))
    soak = max(50, int(0.5 / max(kernel_ms * 1e-3, 1e-5)))        # about 0.5 s of the same load for the clock sampler
    soak = int(st.dist.allreduce(ctx, [soak], "max")[0])           # identical on all ranks
    for _ in range(soak):
        op.vmult(y, x)
    ctx.synchronize()
    clocks = sampler.stop()

    # end-to-end through the host-buffer entry point
    e2e_steps = max(3, min(args.steps, 5))
    op.vmult_host(hy, hx)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        op.vmult_host(hy, hx)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    checksum = float(np.abs(hy).sum())
    # the floor of that path: the same buffers copied up and down concurrently with no kernel in between (all ranks at once)
    hy2, py2 = st.capi.pinned_array((nb, op.n), np.float64)
    barrier()
    copy_ms = op.host_copy_floor(hy2, hx, reps=3)
    barrier()
    st.capi.free_pinned(py2)

    ms_total, e2e_ms, kernel_ms, copy_ms = (float(v) for v in st.dist.allreduce(ctx, [ms_total, e2e_ms, kernel_ms, copy_ms], "max"))
    ms_step = ms_total / args.steps
    total_dofs = dofs_rank * world
    value = total_dofs / (ms_step * 1e-3)
    peak, peak_src = measured_peak()
    achieved = dofs_rank * BYTES_PER_DOF / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(args.config, n)

    config = workload_config(cfg, n, world, grid)
    line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "run": {"kernel_variant": args.variant, "checksum": checksum,
                    "halo": "interface DoFs summed over NVLink inside every step (stores into peer memory over CUDA IPC + sequence flags; ncclSend/ncclRecv where IPC is unavailable)" if world > 1 else "none"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": dofs_rank * BYTES_PER_DOF,
                         "note": "algorithmic bytes = 16 B per space-time DoF (read src once, write dst once); kernel_ms = CUDA-event "
                                 "time of one operator application on the library stream (one st_vmult_brick_kernel launch, no "
                                 "memset); the kernel is FP64-pipe / latency bound, see DESIGN.md 3.0"},
            "e2e": {"value": total_dofs / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(dofs_rank * 8), "d2h_bytes_per_step": int(dofs_rank * 8),
                    "copy_floor_ms": copy_ms, "host_binding": host_binding,
                    "note": "copy_floor_ms = the same pinned buffers uploaded and downloaded concurrently without a kernel, all "
                            "ranks at once, max over ranks (stfem_op_host_copy_floor): what PCIe + host memory allow for this step"},
            "gpu_launches": int(launches), "clocks": clocks}
    if parity is not None:
        line["parity_multi_gpu"] = parity
    for v in (x, y):
        v.free()
    x = y = None
    st.capi.free_pinned(px)
    st.capi.free_pinned(py)
    op.close()
    mesh.close()

    def extra(name, fn):
        try:                                         # an extra leg must never cost the headline line
            line[name] = fn()
        except Exception as e:
            line[name] = {"error": repr(e)[:300]}

    if world == 1 and not args.no_perturbed and args.config == "c2":
        extra("perturbed_mesh", lambda: perturbed_leg(st, ctx, cfg, n))
    if not args.no_solve and args.config == "c2":
        if world == 1:
            extra("solve", lambda: solve_leg(st, ctx, args.solve_refinement))
        else:                                        # collective: every rank must run it, errors must surface
            line["solve"] = solve_leg(st, ctx, args.solve_refinement, grid=grid, coords=coords)
    if not args.no_extra and args.config == "c2":
        if world == 1:
            extra("weak_c5", lambda: weak_c5_leg(st, ctx, make_mesh))
            if not args.no_solve:
                extra("wave_c3", lambda: wave_leg(st, ctx))
        else:
            line["weak_c5"] = weak_c5_leg(st, ctx, make_mesh)
    if world == 1 and not args.no_practical and not args.no_solve and args.config == "c2":
        extra("practical_c4", lambda: practical_leg(st, ctx))
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_cells = 48 if n >= 96 else n
        val, ms, threads, ndofs = cpu_reference_run(cfg, cpu_cells, 3, 1)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "%d^3-cell brick of the same workload (%.3g space-time DoFs), 3 applications "
                                          "of the reference's unfused algorithm (oracle/cpu_ref.cpp, scalar per cell); the full brick is timed by "
                                          "--impl reference"
                                          % (cpu_cells, ndofs)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    barrier()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
