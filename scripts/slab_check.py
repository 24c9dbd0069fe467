"""Multi-GPU slab check (run under torchrun on N GPUs): halo consistency across ranks + throughput.
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/slab_check.py [--bx 1152] [--by 1152]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
from mulskips_b200 import Mskc
from mulskips_b200.workloads import rates_for

ap = argparse.ArgumentParser()
ap.add_argument("--bx", type=int, default=1152); ap.add_argument("--by", type=int, default=1152); ap.add_argument("--bz", type=int, default=96)
ap.add_argument("--sweeps", type=int, default=32); ap.add_argument("--ept", type=float, default=4.0); ap.add_argument("--rounds", type=int, default=3)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = rates_for("surface"); w["PtransZig"] = 0.8
ctx = Mskc(a.bx * world, a.by, a.bz, w["listcrystal"], w["listcov"], device=local, slab=(rank, world))
ctx.set_rates(w["PtransE"], w["PtransD"], w["PtransAbs"], w["PtransDes"], w["PtransZig"])
ids = [ctx.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
ctx.comm_connect(ids[0])
ctx.init_flat(24, walls=False, two_species=True)
ctx.run_parallel(3, events_per_tile=a.ept, seed=1)
for rnd in range(a.rounds):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    r = ctx.run_parallel(a.sweeps, events_per_tile=a.ept, seed=2 + rnd)
    dist.barrier(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ev = torch.tensor([float(r.events_done)], dtype=torch.float64, device="cuda"); dist.all_reduce(ev)
    if rank == 0:
        print(f"round {rnd}: {world} ranks, {ev.item():.0f} events in {dt * 1e3:.1f} ms -> {ev.item() / dt:.4g} ev/s (t={r.time:.4g}, active/rank={r.n_active}, launches={r.launches})", flush=True)
# seam consistency: my right halo column must equal the right neighbour's first owned column, and vice versa
L = torch.from_numpy(ctx.download_lattice()).cuda()
own = a.bx
mine = torch.stack([L[:, :, 24:48], L[:, :, own:own + 24]]).contiguous()              # [first owned, last owned]
halo = torch.stack([L[:, :, own + 24:own + 48], L[:, :, 0:24]]).contiguous()           # [right halo, left halo]
got = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(got, mine)
right, left = (rank + 1) % world, (rank - 1) % world
ok = bool(torch.equal(halo[0], got[right][0])) and bool(torch.equal(halo[1], got[left][1]))
bad = int((halo[0] != got[right][0]).sum() + (halo[1] != got[left][1]).sum())
flag = torch.tensor([0 if ok else 1], device="cuda"); dist.all_reduce(flag)
if rank == 0:
    print("SEAMS", "CONSISTENT" if flag.item() == 0 else f"INCONSISTENT on {int(flag.item())} rank(s)", flush=True)
if not ok:
    print(f"rank {rank}: {bad} differing words", flush=True)
ctx.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 0 else 1)
