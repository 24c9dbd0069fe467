"""tau / total rate / events per call in slab mode over WORLD_SIZE ranks (torchrun) -- the same numbers must come out per rank for any number of ranks"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from mulskips_b200 import Mskc
from mulskips_b200.workloads import rates_for
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = rates_for("surface"); w["PtransZig"] = 0.8
ctx = Mskc(4608 * world, 4608, 96, w["listcrystal"], w["listcov"], device=local, slab=(rank, world))
ctx.set_rates(w["PtransE"], w["PtransD"], w["PtransAbs"], w["PtransDes"], w["PtransZig"])
if world > 1:
    ids = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_connect(ids[0])
ctx.init_flat(24, walls=False, two_species=True)
t_prev = 0.0
for k in range(8):
    r = ctx.run_parallel(32, events_per_tile=4.0, seed=100 + k + rank)
    if rank == 0:
        print(f"world={world} call {k}: events(rank0)={r.events_done} R_total={r.total_rate:.6g} R/rank={r.total_rate / world:.6g} n_active/rank={r.n_active / world:.0f} tau={r.tau:.6g} dt/sweep={(r.time - t_prev) / 32:.6g} ms={r.kernel_ms:.2f}", flush=True)
    t_prev = r.time
ctx.close()
