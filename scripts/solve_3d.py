"""3D heat/wave STMG-FGMRES solve on one GPU through the product driver (HeatWaveProblem over the C ABI).
    python scripts/solve_3d.py [refinement] [space_degree] [time_degree] [CGP|DG] [n_steps] [subdivisions]
Prints per-step wall time (stream-synchronised), iterations and space-time DoFs/s."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dealii_stfem_b200 as st  # noqa: E402

ref = int(sys.argv[1]) if len(sys.argv) > 1 else 3
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
r = int(sys.argv[3]) if len(sys.argv) > 3 else 2
tt = sys.argv[4] if len(sys.argv) > 4 else "CGP"
n_steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
sub = int(sys.argv[6]) if len(sys.argv) > 6 else 3
problem = os.environ.get("PROBLEM", "heat")
pj = {"timeType": tt, "problemType": problem, "feDegree": r, "refinement": ref, "subdivisions": "%d,%d,%d" % (sub, sub, sub),
      "mgTimeBeforeSpace": os.environ.get("TBS", "true"), "smoother": os.environ.get("SMOOTHER", "relaxation"),
      "nTimestepsAtOnce": int(os.environ.get("NTS", "1")), "spaceTimeConvergenceTest": "true"}
p = st.parse_parameters(pj, 3)
ctx = st.Context(0)
t0 = time.perf_counter()
prob = st.HeatWaveProblem(ctx, p, 3, ref, r, space_degree=k)
ctx.synchronize()
print("setup %.2f s  levels %s  N %d  nb %d  tau %.4g  launches %d" % (time.perf_counter() - t0, "".join(prob.mg_type_level), prob.n, prob.nb,
                                                                        prob.tau, ctx.launches), flush=True)
for l in range(prob.mg.n_levels):
    print("  level", l, {k2: (round(v, 4) if isinstance(v, float) else v) for k2, v in prob.mg.level_info(l).items()})
for s in range(n_steps):
    l0 = ctx.launches
    ctx.synchronize()
    t0 = time.perf_counter()
    it = prob.step(evaluate_error=False)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    print("step %d: %d iterations, %.1f ms, %d launches, %.3e st-DoFs/s (solve), %.3e st-DoF-its/s" %
          (s, it, dt * 1e3, ctx.launches - l0, prob.n * prob.nb / dt, prob.n * prob.nb * it / dt), flush=True)
prob.close()
ctx.close()
