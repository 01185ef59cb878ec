"""BASELINE configs[3]: 3D heat with heterogeneous discontinuous coefficients (Coefficient<dim>, reference
include/operators.h:870-965) on a randomly perturbed mesh, DG(2) time, Q3 space, cell-patch (dense Vanka) smoother.
    python scripts/solve_c4.py [refinement] [n_steps]          (PRACTICAL=0: manufactured solution, no coefficient;
                                                                VANKA=half: FP16 patch inverses, 155 -> 86 ms per step)
The reference's practical set-up (tests/json/practical01.json + run_practical.sh: spaceTimeConvergenceTest = false,
box [-1,1]^3, subdivisions 5, distortCoeff 0.6, distortGrid 0.15): coefficient table on K on every level, zero source,
initial value = C-infinity bump of radius 1e-2 (centred on the displaced mesh vertex next to the origin: on a perturbed
mesh the bump around (0,0,0) itself can miss every support point).  The level operators run the general-geometry
kernel, the smoother the dense per-cell patch inverses (the Kronecker form does not apply to distorted cells / variable
coefficients)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dealii_stfem_b200 as st  # noqa: E402

ref = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
k, r = 3, 2
pj = {"timeType": "DG", "problemType": "heat", "feDegree": r, "refinement": ref, "subdivisions": "5,5,5",
      "hyperRectLowerLeft": "-1,-1,-1", "hyperRectUpperRight": "1,1,1", "mgTimeBeforeSpace": "true",
      "smoother": os.environ.get("SMOOTHER", "relaxation"), "spaceTimeConvergenceTest": "true", "distortGrid": 0.15}
pj["vankaStorage"] = os.environ.get("VANKA", "level")       # VANKA=half: FP16 storage of the dense patch inverses
practical = os.environ.get("PRACTICAL", "1") != "0"
if practical:
    pj.update({"spaceTimeConvergenceTest": "false", "distortCoeff": "0.6", "extrapolate": "false"})
p = st.parse_parameters(pj, 3)


def vertices(n_cells):
    n = [c + 1 for c in n_cells]
    g = [np.linspace(-1.0, 1.0, m) for m in n]
    V = np.stack(np.meshgrid(g[2], g[1], g[0], indexing="ij")[::-1], axis=-1)     # [z][y][x][xyz]
    d = np.random.RandomState(1).uniform(-1, 1, V.shape) * 0.15 * (2.0 / n_cells[0])
    d[0] = d[-1] = 0
    d[:, 0] = d[:, -1] = 0
    d[:, :, 0] = d[:, :, -1] = 0
    return V + d


if practical:
    V = vertices([5 << ref] * 3).reshape(-1, 3)
    p["sourcePoint"] = [float(c) for c in V[np.argmin(np.sum(V * V, axis=1))]]
ctx = st.Context(0)
t0 = time.perf_counter()
prob = st.HeatWaveProblem(ctx, p, 3, ref, r, space_degree=k, vertices_fn=vertices)
ctx.synchronize()
print("setup %.2f s  levels %s  N %d  nb %d  cells %d" % (time.perf_counter() - t0, "".join(prob.mg_type_level), prob.n, prob.nb,
                                                         int(np.prod(prob.n_cells))), flush=True)
for l in range(prob.mg.n_levels):
    i = prob.mg.level_info(l)
    print("  level %d: N %d blocks %d patches %d (%.1f MB) lambda %.3f" % (l, i["N"], i["blocks"], i["patch_matrices"], i["patch_bytes"] / 1e6, i["lambda"]))
for s in range(n_steps):
    ctx.synchronize()
    t0 = time.perf_counter()
    it = prob.step(evaluate_error=False)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    print("step %d: %d iterations, %.1f ms, %.3e st-DoFs/s (solve)" % (s, it, dt * 1e3, prob.n * prob.nb / dt), flush=True)
if prob.functional_rows:
    print("point functionals at t = %.5f: %s" % (prob.functional_rows[-1][0], " ".join("%.6e" % v for v in prob.functional_rows[-1][1:])))
prob.close()
ctx.close()
