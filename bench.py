#!/usr/bin/env python3
"""bench.py — headline benchmark of the space-time operator hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one application of the fused space-time operator  dst = (Alpha (x) K + Beta (x) M) src
(reference SystemMatrix::vmult, include/operators.h:536-559) over the whole block vector.
Workload at N=1 = BASELINE.json configs[1]: 3D heat, Q4 space x cG(2) time, 96^3 cells
(subdivisions 3, refinement 5), 385^3 spatial DoFs x 2 time blocks = 1.14e8 space-time DoFs, FP64.
For N>1 every rank owns one such brick (weak scaling, box partition of the mesh).

Prints ONE JSON line (rank 0).  `value` = space-time DoFs/s with inputs resident in HBM (CUDA events on
the library's stream, max over ranks); `e2e` = the same metric through the C-ABI entry point with
pinned HOST buffers (H2D + kernel + D2H inside the timed region); `roofline`, `cpu_baseline`,
`clocks`, `gpu_launches` as the contract asks.  torch is used for torch.distributed plumbing only.
Extra objects on the same line (N=1): `perturbed_mesh` (general-geometry kernel), `solve` (STMG-FGMRES time steps of
configs[1]) and `practical_c4` (configs[3]: heterogeneous coefficient, perturbed mesh, dense cell-patch smoother, with the
patch inverses in float and in FP16).
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
DEGREE, TTYPE, TDEG = 4, "CGP", 2
N_CELLS_FULL = 96            # per direction: subdivisions 3, refinement 5
CPU_SAMPLE_CELLS = 48        # bounded CPU sample of the same workload (1/8 of the cells)
BYTES_PER_DOF = 16           # FP64: read src once + write dst once


def time_weights():
    """Alpha, Beta of cG(2) for tau = 2^-6 (tests/tp_01.cc:106-109 with refinement 5)."""
    from dealii_stfem_b200 import fe_time_host
    return fe_time_host.get_fe_time_weights(TTYPE, TDEG, 2.0 ** -6, 1)[:2]


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


def cpu_reference_run(n_cells, steps, warmup):
    """The reference's CPU algorithm (unfused 2*nb cell loops + axpys, operators.h:536-559) through
    oracle/cpu_ref.cpp on all host threads.  Returns (DoFs/s, ms per step, threads)."""
    from oracle import cpu_ref, fe_time as oft, spatial as osp
    A, B = oft.get_fe_time_weights(TTYPE, TDEG, 2.0 ** -6, 1)[:2]
    mesh = osp.Mesh(3, [n_cells] * 3, 0)
    space = osp.Space(mesh, DEGREE)
    nb = A.shape[0]
    src = np.sin(0.1 * np.arange(space.n_dofs)[None, :] + np.arange(nb)[:, None])
    threads = cpu_ref.max_threads()
    for _ in range(warmup):
        cpu_ref.system_vmult(space, A, B, src)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_ref.system_vmult(space, A, B, src)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return nb * space.n_dofs / dt, dt * 1e3, threads, nb * space.n_dofs


def ncu_traffic(n_cells):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the headline kernel from the committed ncu capture
    (profiles/), valid for the default workload only."""
    p = os.path.join(ROOT, "profiles", "r01_st_vmult_cart_ncu.json")
    if n_cells != N_CELLS_FULL or not os.path.exists(p):
        return None
    return json.load(open(p)).get("dram_bytes_per_launch")


def perturbed_leg(st, ctx, n, steps=20):
    """The same operator on a randomly perturbed mesh (GridTools::distort_random-style, factor 0.15) with a per-q
    coefficient table: general-geometry kernel (csrc/st_vmult_plane.cuh) with the precomputed metric."""
    A, B = time_weights()
    n1 = n + 1
    g = np.linspace(0.0, 1.0, n1)
    V = np.stack(np.meshgrid(g, g, g, indexing="ij")[::-1], axis=-1)
    d = np.random.RandomState(1).uniform(-1, 1, V.shape) * 0.15 / n
    d[0] = d[-1] = 0
    d[:, 0] = d[:, -1] = 0
    d[:, :, 0] = d[:, :, -1] = 0
    mesh = st.Mesh(ctx, [n, n, n], vertices=(V + d).reshape(-1, 3))
    op = st.Operator(mesh, DEGREE, A, B, number_type=st.F64)
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
    # algorithmic bytes: vectors + the metric the kernel streams (8 numbers per quadrature point)
    alg = dofs * BYTES_PER_DOF + n ** 3 * (DEGREE + 1) ** 3 * 8 * 8
    out = {"metric": "space-time DoFs/s, operator vmult on a perturbed mesh (MappingQ1 cells, precomputed metric), FP64",
           "value": dofs / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "cells": n ** 3, "distortion": 0.15,
           "achieved_GBs": alg / (ms * 1e-3) / 1e9, "algorithmic_bytes_per_launch": alg}
    x.free(); y.free(); op.close(); mesh.close()
    return out


def solve_leg(st, ctx, refinement, n_steps=3, grid=None, coords=None, reduce_max=None):
    """Full STMG-preconditioned FGMRES time steps of configs[1] (3D heat, Q4 x cG(2), float multigrid) through the
    product driver: rhs assembly + solve per step, all on the device (tests/tp_01.cc:646-669)."""
    grid = grid or [1, 1, 1]
    world = int(np.prod(grid))
    pj = {"timeType": TTYPE, "problemType": "heat", "feDegree": TDEG, "refinement": refinement,
          "subdivisions": ",".join(str(3 * g) for g in grid), "hyperRectUpperRight": ",".join(str(float(g)) for g in grid),
          "mgTimeBeforeSpace": "true", "smoother": "relaxation", "spaceTimeConvergenceTest": "true",
          "agglomerateBelow": os.environ.get("STFEM_AGGLO", "16"),
          "levelKernelVariant": os.environ.get("STFEM_LEVEL_VARIANT", "0")}
    p = st.parse_parameters(pj, 3)
    t0 = time.perf_counter()
    prob = st.HeatWaveProblem(ctx, p, 3, refinement, TDEG, space_degree=DEGREE,
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
    if reduce_max is not None:
        ms = reduce_max(ms)
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
    ap.add_argument("--cells", type=int, default=N_CELLS_FULL, help="cells per direction per GPU (default 96)")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-solve", action="store_true", help="skip the STMG-FGMRES solve leg")
    ap.add_argument("--no-perturbed", action="store_true", help="skip the perturbed-mesh operator leg")
    ap.add_argument("--no-practical", action="store_true", help="skip the configs[3] (practical set-up, dense Vanka) leg")
    ap.add_argument("--solve-refinement", type=int, default=5, help="solve leg: subdivisions 3, this many refinements (5 = 96^3 cells)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        val, ms, threads, ndofs = cpu_reference_run(CPU_SAMPLE_CELLS, args.steps, args.warmup)
        sample = "%d^3-cell Q4 x cG(2) brick (%.3g space-time DoFs) per step, all host threads" % (CPU_SAMPLE_CELLS, ndofs)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "3D heat Q4 x cG(2) operator vmult; CPU arm runs the reference's unfused "
                                       "algorithm (oracle/cpu_ref.cpp, the reference needs deal.II and cannot be built)",
                           "sample": sample},
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = st.Context(dev)
    A, B = time_weights()
    n = args.cells
    grid, coords = [1, 1, 1], [0, 0, 0]
    if world > 1:
        # box partition 2x1x1, 2x2x1, 2x2x2: every rank owns one n^3 brick (unit cube) of the global mesh; the operator
        # sums interface DoFs over NVLink (ncclSend/ncclRecv) inside every vmult
        grid = st.dist.proc_grid_for(world, 3)
        coords = st.dist.coords_of(rank, grid)

        def bcast(b):
            t = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                t = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
            dist.broadcast(t, 0)
            return bytes(t.cpu().numpy().tobytes())

        st.dist.init_comm(ctx, rank, world, bcast)
        n_loc, _, llo, lup, mask = st.dist.partition_brick([n * g for g in grid], [0.0] * 3, [float(g) for g in grid], grid, coords)
        mesh = st.Mesh(ctx, n_loc, lower=llo, upper=lup, dirichlet_faces=mask)
        st.dist.set_partition(mesh, grid, coords)
    else:
        mesh = st.Mesh(ctx, [n, n, n])
    op = st.Operator(mesh, DEGREE, A, B, number_type=st.F64, variant=args.variant)
    nb = op.nb_rows
    dofs_rank = op.n * nb
    x, y = op.new_vector(), op.new_vector()
    hx, px = st.capi.pinned_array((nb, op.n), np.float64)
    hy, py = st.capi.pinned_array((nb, op.n), np.float64)
    hx[:] = np.sin(0.1 * np.arange(op.n)[None, :] + np.arange(nb)[:, None])      # SURVEY §8d synthetic input
    x.upload(hx)

    # device-resident timing; clocks are sampled from the warm-up to the end of the kernel-event loop (all the same
    # kernel under load) because the timed region itself is shorter than nvidia-smi's sampling period
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
    t_soak = time.perf_counter()
    while time.perf_counter() - t_soak < 0.4:      # keep the same load up until the sampler has a few readings
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

    t = torch.tensor([ms_total, e2e_ms, kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, kernel_ms = (float(v) for v in t.tolist())
    ms_step = ms_total / args.steps
    total_dofs = dofs_rank * world
    value = total_dofs / (ms_step * 1e-3)
    peak, peak_src = measured_peak()
    achieved = dofs_rank * BYTES_PER_DOF / (kernel_ms * 1e-3) / 1e9

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: 3D heat, Q4 x cG(2), %d^3 cells per GPU, %d spatial DoFs x %d time blocks "
                                   "= %.4g space-time DoFs per GPU; one step = one fused operator vmult" % (n, op.n, nb, dofs_rank),
                       "l2": "inputs+outputs %.0f MB per step, larger than the 126 MB L2" % (2 * dofs_rank * 8 / 1e6),
                       "parallelism": "box partition %s, one brick per GPU%s" % ("x".join(map(str, grid)), ", interface DoFs summed "
                                       "over NVLink (ncclSend/ncclRecv) inside every step" if world > 1 else ""),
                       "kernel_variant": args.variant, "checksum": checksum},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(n), "peak_source": peak_src, "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": dofs_rank * BYTES_PER_DOF,
                         "note": "algorithmic bytes = 16 B per space-time DoF; kernel_ms = event time of memset(dst) + "
                                 "st_vmult_cart_kernel; the kernel is FP64-pipe / L1 bound (86 DFMA per DoF), see DESIGN.md 3.1"},
            "e2e": {"value": total_dofs / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(dofs_rank * 8), "d2h_bytes_per_step": int(dofs_rank * 8)},
            "gpu_launches": int(launches), "clocks": clocks}
    if world == 1 and not args.no_perturbed:
        line["perturbed_mesh"] = perturbed_leg(st, ctx, n)
    if not args.no_solve:
        for v in (x, y):
            v.free()
        x = y = None

        def reduce_max(v):
            if world == 1:
                return v
            tt = torch.tensor([v], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())

        line["solve"] = solve_leg(st, ctx, args.solve_refinement, grid=grid, coords=coords, reduce_max=reduce_max)
    if world == 1 and not args.no_practical and not args.no_solve:
        try:
            line["practical_c4"] = practical_leg(st, ctx)
        except Exception as e:                      # an extra leg must never cost the headline line
            line["practical_c4"] = {"error": repr(e)[:300]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, ms, threads, ndofs = cpu_reference_run(CPU_SAMPLE_CELLS, 3, 1)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "%d^3-cell brick of the same workload (%.3g space-time DoFs), 3 applications "
                                          "of the reference's unfused algorithm (oracle/cpu_ref.cpp)" % (CPU_SAMPLE_CELLS, ndofs)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    for v in (x, y):
        if v is not None:
            v.free()
    st.capi.free_pinned(px)
    st.capi.free_pinned(py)
    op.close()
    mesh.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
