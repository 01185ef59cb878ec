"""Multi-GPU parity check, run under torchrun with N = 2, 4 or 8 ranks (one per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py
Every rank also solves the GLOBAL problem alone on its own GPU (unpartitioned) and compares its brick of
  (1) the FP64 operator vmult            (relative 1e-12),
  (2) one STMG V-cycle (float levels)     (relative 2e-3: float Vanka amplification),
  (3) a full time step (rhs + FGMRES)     (iterations +-1, solution 1e-8)
with the partitioned run (dealii_stfem_b200.dist.parity_check, the same check bench.py --gpus N prints as
"parity_multi_gpu").  torch.distributed only distributes the NCCL unique id.  Exit code 0 = all ranks agree."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dealii_stfem_b200 as st  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
ref = int(sys.argv[1]) if len(sys.argv) > 1 else 2


def bcast(b):
    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        t = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())


ctx = st.Context(local)
st.dist.init_comm(ctx, rank, world, bcast)
torch.cuda.synchronize()
res = st.dist.parity_check(ctx, local, rank, world, refinement=ref, extended=True)
if rank == 0:
    print("parity_multi_gpu " + json.dumps(res), flush=True)
print("rank %d: %s" % (rank, "ok" if res["ok"] else "FAIL"), flush=True)
ctx.synchronize()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if res["ok"] else 1)
