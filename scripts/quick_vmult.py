"""Quick device timing of st_vmult (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dealii_stfem_b200 as st
from oracle import fe_time as ft

def run(n, k, ttype, r, nt=0, variant=0, reps=5):
    ctx = st.Context(0)
    A, B, _, _ = ft.get_fe_time_weights(ttype, r, 0.01, 1)
    gm = st.Mesh(ctx, [n, n, n])
    op = st.Operator(gm, k, A, B, number_type=nt, variant=variant)
    nb = A.shape[0]
    x = op.new_vector(); y = op.new_vector()
    src = np.sin(0.1 * np.arange(op.n)[None, :] + np.arange(nb)[:, None])
    x.upload(src)
    op.set_timing(True)
    ts = []
    for i in range(reps + 2):
        op.vmult(y, x)
        ts.append(op.last_kernel_ms())
    t = min(ts[2:])
    dofs = op.n * nb
    print("n=%d k=%d %s(%d) nt=%d variant=%d: %.3f ms  %.3e DoF/s  %.1f GB/s(alg)" % (
        n, k, ttype, r, nt, variant, t, dofs / t * 1e3, dofs * (16 if nt == 0 else 8) / t * 1e-6), flush=True)
    x.free(); y.free(); op.close(); gm.close(); ctx.close()

if __name__ == "__main__":
    v = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    run(48, 4, "CGP", 2, 0, v)
    run(96, 4, "CGP", 2, 0, v)
    run(96, 4, "CGP", 2, 1, v)
    run(96, 3, "DG", 2, 0, v)
    run(64, 2, "DG", 1, 0, v)
