"""The two kernels that dominate a V-cycle, on the FINEST multigrid level of configs[1] (3D heat, Q4 x cG(2), 96^3 cells, FP32):
PreconditionVanka::vmult in Kronecker form (k_vanka_fd) and the level operator (st_vmult_brick_kernel<float>), timed with
CUDA events through stfem_mg_level_apply.    python scripts/level_kernels.py [refinement]
(profile with:  PROFILE=1 ncu --profile-from-start off --set full -k regex:k_vanka_fd -s 4 -c 1 ...   /   -k regex:st_vmult_brick ...)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dealii_stfem_b200 as st  # noqa: E402

ref = int(sys.argv[1]) if len(sys.argv) > 1 else 5
pj = {"timeType": "CGP", "problemType": "heat", "feDegree": 2, "refinement": ref, "subdivisions": "3,3,3",
      "mgTimeBeforeSpace": "true", "smoother": "relaxation", "spaceTimeConvergenceTest": "true"}
p = st.parse_parameters(pj, 3)
ctx = st.Context(0)
prob = st.HeatWaveProblem(ctx, p, 3, ref, 2, space_degree=4)
level = prob.mg.n_levels - 1
lop = prob.level_ops[-1]
dx, dy = lop.new_vector(), lop.new_vector()
dx.upload(np.sin(0.1 * np.arange(lop.n)[None, :] + np.arange(lop.nb_rows)[:, None]).astype(np.float32))
dofs = lop.n * lop.nb_rows
if os.environ.get("PROFILE"):
    # ncu --profile-from-start off: only the launches below are candidates (the set-up launches hundreds of small kernels)
    import ctypes
    ctx.synchronize()
    ctypes.CDLL("libcudart.so").cudaProfilerStart()
for what, name, bytes_per_dof in ((0, "PreconditionVanka::vmult (k_vanka_fd, float)", 12), (4, "level operator vmult (brick kernel, float)", 8)):
    for _ in range(3):
        prob.mg.level_apply(level, what, dy, dx)
    ctx.timer_start()
    for _ in range(10):
        prob.mg.level_apply(level, what, dy, dx)
    ms = ctx.timer_stop() / 10.0
    print("%-48s %.3f ms  %.3e st-DoFs/s  %.0f GB/s algorithmic (%d B per st-DoF)"
          % (name, ms, dofs / ms * 1e3, dofs * bytes_per_dof / ms / 1e6, bytes_per_dof), flush=True)
dx.free(); dy.free(); prob.close(); ctx.close()
