"""Multi-GPU parity check, run under torchrun with N = 2, 4 or 8 ranks (one per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py
Every rank also solves the GLOBAL problem alone on its own GPU (unpartitioned) and compares its brick of
  (1) the FP64 operator vmult            (relative 1e-12),
  (2) one STMG V-cycle (float levels)     (relative 2e-3: float Vanka amplification),
  (3) a full time step (rhs + FGMRES)     (iterations +-1, solution 1e-8)
with the partitioned run.  Exit code 0 = all ranks agree."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dealii_stfem_b200 as st  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
ref = int(sys.argv[1]) if len(sys.argv) > 1 else 2
k, r, tt = 2, 1, "DG"
pj = {"timeType": tt, "problemType": "heat", "feDegree": r, "refinement": ref, "subdivisions": "2,2,2", "mgTimeBeforeSpace": "true",
      "smoother": "relaxation", "spaceTimeConvergenceTest": "true", "agglomerateBelow": os.environ.get("AGGLO", "16")}
p = st.parse_parameters(pj, 3)
grid = st.dist.proc_grid_for(world, 3)
coords = st.dist.coords_of(rank, grid)


def bcast(b):
    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        t = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())


ok = True


def check(name, err, tol):
    global ok
    good = bool(err <= tol)
    ok = ok and good
    print("rank %d: %-28s %.3e (tol %.1e) %s" % (rank, name, err, tol, "ok" if good else "FAIL"), flush=True)


# ---- global problem on this GPU
ctx0 = st.Context(local)
glob = st.HeatWaveProblem(ctx0, p, 3, ref, r, space_degree=k)
ng = [2 * (1 << ref)] * 3
npg = [k * n + 1 for n in ng]
nb = glob.nb
rng = np.random.RandomState(7)
xg = rng.uniform(-1, 1, (nb, glob.n))
dx, dy = glob.matrix.new_vector().upload(xg), glob.matrix.new_vector()
glob.matrix.vmult(dy, dx)
Ag = dy.download()
# V-cycle on a residual-like vector (zero on the boundary)
glob.mg.vmult(dy, dx.upload(Ag))
Vg = dy.download()
it_g = glob.step(evaluate_error=False)
sol_g = glob.x.download()
dx.free(); dy.free()

# ---- partitioned problem
ctx = st.Context(local)
st.dist.init_comm(ctx, rank, world, bcast)
part = st.HeatWaveProblem(ctx, p, 3, ref, r, space_degree=k, partition=(grid, coords))
nl = [n // g for n, g in zip(ng, grid)]
npl = [k * n + 1 for n in nl]
off = [k * nl[d] * coords[d] for d in range(3)]
sl = (slice(None), slice(off[2], off[2] + npl[2]), slice(off[1], off[1] + npl[1]), slice(off[0], off[0] + npl[0]))


def brick(a):
    return np.ascontiguousarray(a.reshape(nb, npg[2], npg[1], npg[0])[sl]).reshape(nb, -1)


assert part.n == npl[0] * npl[1] * npl[2]
dx, dy = part.matrix.new_vector().upload(brick(xg)), part.matrix.new_vector()
part.matrix.vmult(dy, dx)
check("vmult (FP64)", np.abs(dy.download() - brick(Ag)).max() / np.abs(Ag).max(), 1e-12)
part.mg.vmult(dy, dx.upload(brick(Ag)))
check("V-cycle (FP32 levels)", np.abs(dy.download() - brick(Vg)).max() / np.abs(Vg).max(), 2e-3)
it_p = part.step(evaluate_error=False)
check("time step: iterations", abs(it_p - it_g), 1)
check("time step: solution", np.abs(part.x.download() - brick(sol_g)).max() / np.abs(sol_g).max(), 1e-8)
if rank == 0:
    print("levels %s, %d ranks as %s, global N %d, local N %d, iterations global %d / partitioned %d" %
          ("".join(part.mg_type_level), world, grid, glob.n, part.n, it_g, it_p), flush=True)
dx.free(); dy.free()
t = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(t)
part.close(); glob.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 0 else 1)
