"""Host-buffer entry point (stfem_op_vmult_host) on configs[1]: time per call and the copy floor.  STFEM_HOST_SLABS sets the
number of z slabs of the upload / kernel / download pipeline.    python scripts/e2e_slabs.py [cells]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dealii_stfem_b200 as st  # noqa: E402
from dealii_stfem_b200 import fe_time_host as ft  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 96
A, B = ft.get_fe_time_weights("CGP", 2, 2.0 ** -6, 1)[:2]
ctx = st.Context(0)
mesh = st.Mesh(ctx, [cells] * 3)
op = st.Operator(mesh, 4, A, B)
nb = op.nb_rows
hx, px = st.capi.pinned_array((nb, op.n), np.float64)
hy, py = st.capi.pinned_array((nb, op.n), np.float64)
hx[:] = np.sin(0.1 * np.arange(op.n)[None, :] + np.arange(nb)[:, None])
for _ in range(2):
    op.vmult_host(hy, hx)
ctx.synchronize()
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    op.vmult_host(hy, hx)
ms = (time.perf_counter() - t0) * 1e3 / reps
floor = op.host_copy_floor(hy, hx, reps=3)
print("slabs %s: vmult_host %.2f ms per call, copy floor %.2f ms (%.1f GB/s each way), checksum %.6e"
      % (os.environ.get("STFEM_HOST_SLABS", "16"), ms, floor, op.n * nb * 8 / floor / 1e6, float(np.abs(hy).sum())))
