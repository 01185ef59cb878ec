"""Time the st_vmult kernel variants on one GPU (device-resident, CUDA events).  Usage:
    python scripts/tune_vmult.py [cells] [degree] [f64|f32] [variants...]"""
import os
import sys

import numpy as np

os.environ.setdefault("STFEM_ALLOW_ABLATION", "1")     # this tuning tool may time the ablation builds (variants 31-34)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dealii_stfem_b200 as st  # noqa: E402
from dealii_stfem_b200 import fe_time_host as ft  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 96
degree = int(sys.argv[2]) if len(sys.argv) > 2 else 4
nt = st.F64 if (len(sys.argv) <= 3 or sys.argv[3] == "f64") else st.F32
variants = [int(v) for v in sys.argv[4:]] or [0]
ttype, tdeg = os.environ.get("TT", "CGP"), int(os.environ.get("TR", "2"))
A, B = ft.get_fe_time_weights(ttype, tdeg, 2.0 ** -6, int(os.environ.get("NTS", "1")))[:2]
ctx = st.Context(0)
distort = float(os.environ.get("DISTORT", "0"))
if distort > 0:
    # perturbed mesh: interior vertices moved by up to distort * h in every direction (deterministic)
    n1 = cells + 1
    g = np.linspace(0.0, 1.0, n1)
    V = np.stack(np.meshgrid(g, g, g, indexing="ij")[::-1], axis=-1)          # [z][y][x][xyz]
    d = np.random.RandomState(1).uniform(-1, 1, V.shape) * distort / cells
    d[0, :, :, :] = d[-1, :, :, :] = 0; d[:, 0, :, :] = d[:, -1, :, :] = 0; d[:, :, 0, :] = d[:, :, -1, :] = 0
    mesh = st.Mesh(ctx, [cells] * 3, vertices=(V + d).reshape(-1, 3))
else:
    mesh = st.Mesh(ctx, [cells] * 3)
ref = None
for variant in variants:
    op = st.Operator(mesh, degree, A, B, number_type=nt, variant=variant)
    nb = op.nb_rows
    x, y = op.new_vector(), op.new_vector()
    x.upload(np.sin(0.1 * np.arange(op.n)[None, :] + np.arange(nb)[:, None]))
    for _ in range(3):
        op.vmult(y, x)
    ctx.timer_start()
    reps = 10
    for _ in range(reps):
        op.vmult(y, x)
    ms = ctx.timer_stop() / reps
    out = y.download().astype(np.float64)
    if ref is None:
        ref = out
    err = np.abs(out - ref).max() / np.abs(ref).max()
    dofs = op.n * nb
    print("variant %2d  %s  Q%d nb=%d  %d^3 cells  %.3f ms  %.3e DoFs/s  %.1f GB/s(alg)  rel.diff to first %.1e"
          % (variant, "f64" if nt == st.F64 else "f32", degree, nb, cells, ms, dofs / ms * 1e3,
             dofs * (16 if nt == st.F64 else 8) / ms / 1e6, err), flush=True)
    x.free(); y.free(); op.close()
mesh.close(); ctx.close()
