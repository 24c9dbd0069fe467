#!/usr/bin/env python3
"""bench.py -- KMC events/s of the MulSKIPS event loop on B200 (and the CPU reference arm).

One "step" = one pass of the hot path over one batch: `--sweeps` synchronous-sublattice sweeps of the
whole slab (4 colour steps each).  Workload (config.workload): flat 3C-SiC(001) slab, the reference's
own PVD rate table (regression-tests start-surface.dat) with PtransZig=0.8 (hexagonal/stacking-fault
branch enabled, as in start-cube-zigzag.dat), 4608x4608x96 super-lattice cells (2.0e9 sites; the 4608x4608 surface of the 10^9-site C5 slab with head-room for roughness) PER GPU,
built on the device with the geometry of initial/SetSiC3C.f.  N>1: one slab per rank (weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "kmc_events_per_s"
UNIT = "events/s"
B_PER_SITE_UPDATE = 96     # algorithmic bytes per (event,rate) re-evaluation
B_PER_EVENT = 448          # algorithmic bytes per applied event


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def workload(args):
    from mulskips_b200.workloads import rates_for
    w = rates_for("surface")
    w["PtransZig"] = 0.8
    return w


# ------------------------------------------------------------------------------------------ reference arm
def oracle_sample(args, faithful_cost, events, box=(480, 480, 96)):
    """The CPU path on a bounded sample of the same workload: same rate table, same flat F builder
    (Len1=24), a 480x480x96 patch, `events` events of the serial n-fold-way loop with ran3 ('R')."""
    from cases import make_start
    from oracle_binding import Oracle
    with tempfile.TemporaryDirectory() as td:
        start = make_start("surface", os.path.join(td, "start.dat"), lens=(24, 0, 0), maxiter=10**9, outmolmol=10**9, runtype="R", ptranszig=0.8)
        orc = Oracle(*box, start)
        orc.set_run(faithful_cost=int(faithful_cost))
        orc.setup()
        t0 = time.perf_counter()
        n = orc.run(events)
        dt = time.perf_counter() - t0
        su = orc.quirks()["n_getprob"]
        orc.close()
    return n, su, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle_binding import build_oracle
    build_oracle()
    ev_per_step = 1_000_000
    for _ in range(args.warmup and 1):
        oracle_sample(args, True, 200_000)
    tot_n, tot_t, tot_su = 0, 0.0, 0
    for _ in range(args.steps):
        n, su, dt = oracle_sample(args, True, ev_per_step)
        tot_n += n; tot_t += dt; tot_su += su
    v = tot_n / tot_t
    n2, su2, dt2 = oracle_sample(args, False, ev_per_step)
    sample = f"C restatement of mulskips.e (oracle/, -O2, 1 thread, per-call rate-table copy as in Get_Prob.f:151-152), 480x480x96 patch of the same SiC slab, {ev_per_step} events per step, ran3"
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64",
           "data": "synthetic", "config": config_dict(args, "cpu"),
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample, "site_updates_per_s": tot_su / tot_t,
                            "value_without_table_copy": n2 / dt2, "host_cores_available": os.cpu_count(),
                            "note": "the reference loop is strictly serial (no threads/MPI); gfortran is not in the image so oracle/_ref cannot be built"},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def config_dict(args, where):
    bx, by, bz = args.box
    return {"workload": f"3C-SiC(001) flat slab PVD, start-surface.dat rate table, PtransZig=0.8, {bx}x{by}x{bz} super-lattice cells ({bx * by * bz / 1e9:.2f}e9 sites) per GPU, F builder Len1={args.len1}",
            "box_per_gpu": [bx, by, bz], "sweeps_per_step": args.sweeps, "events_per_tile": args.ept, "mode": "synchronous-sublattice sampler (24-unit tiles, 2x2 colours, Philox4x32-10)",
            "parallelism": ("one periodic box on one GPU" if args.gpus == 1 else
                            f"one periodic box of {bx * args.gpus}x{by} cells cut into {args.gpus} x-slabs (one per GPU), 24-unit halo tile columns exchanged after every colour step (NCCL send/recv), total rate all-reduced per sweep"), "l2": "inputs (state+bonds+slot, 14 B/cell) >> 126 MB L2; no flush", "where": where}


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the KMC event loop has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mulskips_b200 import Mskc
    import ctypes as C

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w = workload(args)
    bx, by, bz = args.box
    if world > 1:
        # one global periodic box of world*bx x by cells, cut into x-slabs: halo-tile exchange after every colour step
        # (NCCL send/recv inside libmskc) and one all-reduce of the total rate per sweep
        ctx = Mskc(bx * world, by, bz, w["listcrystal"], w["listcov"], device=local, slab=(rank, world))
        ids = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_connect(ids[0])
    else:
        ctx = Mskc(bx, by, bz, w["listcrystal"], w["listcov"], device=local)
    ctx.set_rates(w["PtransE"], w["PtransD"], w["PtransAbs"], w["PtransDes"], w["PtransZig"])
    lx = ctx.LenX                      # local LenX (owned block + 2 halo tile columns on a slab)

    def fresh():
        ctx.init_flat(args.len1, walls=False, two_species=True)

    # ---- device-resident throughput ("value") -----------------------------------------------------
    fresh()
    seed = 1000 + rank
    for i in range(args.warmup):
        ctx.run_parallel(args.sweeps, events_per_tile=args.ept, seed=seed + i)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    ev = su = launches = 0; kms = 0.0
    for i in range(args.steps):
        r = ctx.run_parallel(args.sweeps, events_per_tile=args.ept, seed=seed + 100 + i)
        ev += r.events_done; su += r.site_updates; launches += r.launches; kms += r.kernel_ms
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop() if rank == 0 else None
    dt = t1 - t0
    n_active = r.n_active

    # ---- roofline of the dominant kernel: every colour-step launch timed with CUDA events ---------
    rr = ctx.run_parallel(min(args.sweeps, 16), events_per_tile=args.ept, seed=seed + 999, time_kernels=True)
    n_launch = rr.sampler_launches
    alg_bytes = B_PER_SITE_UPDATE * rr.site_updates + B_PER_EVENT * rr.events_done
    peak, peak_src = peaks()
    achieved = alg_bytes / (rr.sweep_kernel_ms * 1e-3) / 1e9 if rr.sweep_kernel_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "colour_step_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "colour_step_kernel<8, 0>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes / max(n_launch, 1), "launch_ms": rr.sweep_kernel_ms / max(n_launch, 1),
                "share_of_step": rr.sweep_kernel_ms / rr.kernel_ms if rr.kernel_ms else None,
                "note": "latency / instruction-issue bound (27 % of the issue slots, 12.9 of 32 lanes per instruction, ncu in profiles/): the surface working set "
                        "(~300 MB) is scattered 4-16 B records, so DRAM traffic per launch (64 B granules, L2 hit 52 %) is ~1.7x the algorithmic bytes while "
                        "HBM itself idles at 7 % -- the fraction of HBM peak is small by construction (DESIGN.md section 3.1)"}

    # ---- end to end through the C ABI with HOST buffers ------------------------------------------
    # One step of what the host program does between two output frames of a PVD/CVD run: (H2D) the step's inputs --
    # the four rate tables (pymulskips rewrites them per run / per coupled cycle), from pinned memory through
    # mskc_set_rates; the sweeps; (D2H) the step's result -- the frame: every cell WriteMolMolXYZFile.f prints
    # (surface atoms, coverage, non-epitaxial bulk), compacted on the device by mskc_extract_frame into pinned memory.
    u32p, f64p = C.POINTER(C.c_uint32), C.POINTER(C.c_double)
    tab = torch.from_numpy(np.concatenate([w["PtransE"].ravel(), w["PtransD"].ravel(), w["PtransAbs"].ravel(), w["PtransDes"].ravel()]).astype(np.float64)).pin_memory()
    tp = tab.data_ptr()
    fcap = 16 * 1024 * 1024
    fcell, fword = torch.empty(fcap, dtype=torch.int32).pin_memory(), torch.empty(fcap, dtype=torch.int32).pin_memory()
    fn, fms = C.c_int64(), C.c_float()
    fresh()
    frame_ms = []

    def e2e_step(i):
        ctx._ck(ctx.lib.mskc_set_rates(ctx.h, C.cast(tp, f64p), C.cast(tp + 8 * 12288, f64p), C.cast(tp + 8 * (12288 + 9), f64p), C.cast(tp + 8 * (12288 + 21), f64p), w["PtransZig"]))
        rr_ = ctx.run_parallel(args.sweeps, events_per_tile=args.ept, seed=seed + 5000 + i)
        ctx._ck(ctx.lib.mskc_extract_frame(ctx.h, fcap, C.cast(fcell.data_ptr(), u32p), C.cast(fword.data_ptr(), u32p), C.byref(fn), C.byref(fms)))
        frame_ms.append(fms.value)
        return rr_.events_done
    for i in range(min(args.warmup, 2)):
        e2e_step(i)
    frame_ms.clear()
    barrier()
    t2 = time.perf_counter()
    ev2 = 0
    for i in range(args.steps):
        ev2 += e2e_step(10 + i)
    barrier()
    t3 = time.perf_counter()
    nrec = min(int(fn.value), fcap)
    occ = int(((fword[:nrec] >> 23) & 1).sum())          # touch the result
    h2d_step, d2h_step = 8 * (12288 + 9 + 12 + 12), 8 * nrec
    fms_avg = sum(frame_ms) / max(len(frame_ms), 1)
    dense = {"kernel": "frame_extract_kernel", "bound": "hbm", "algorithmic_bytes_per_launch": 4 * (bz + 1) * by * lx, "launch_ms": fms_avg,
             "achieved": 4 * (bz + 1) * by * lx / (fms_avg * 1e-3) / 1e9 if fms_avg > 0 else 0.0, "unit": "GB/s"}
    dense["peak"], _ = peaks(); dense["frac"] = dense["achieved"] / dense["peak"]

    # ---- the laser-annealing coupling cycle, for reference (3 steps): per-cell int32 temperature field up
    # (SetT.f / the socket's SENDFIELD), sweeps, whole lattice down (GETPHASES) -----------------------------------
    la_cycle = None
    if world == 1 and not args.no_la_cycle:
        ncell_T = lx * by * bz
        T_host = torch.full((ncell_T,), 1000, dtype=torch.int32).pin_memory()
        L_host = torch.empty(((bz + 1) * by * lx,), dtype=torch.int32).pin_memory()
        i32p = C.POINTER(C.c_int32)
        Tp, Lp = C.cast(T_host.data_ptr(), i32p), C.cast(L_host.data_ptr(), i32p)
        ev3 = 0
        torch.cuda.synchronize(); t4 = time.perf_counter()
        for i in range(3):
            ctx._ck(ctx.lib.mskc_set_temperature(ctx.h, Tp))
            ev3 += ctx.run_parallel(args.sweeps, events_per_tile=args.ept, seed=seed + 7000 + i).events_done
            ctx._ck(ctx.lib.mskc_download_lattice(ctx.h, Lp))
        torch.cuda.synchronize(); t5 = time.perf_counter()
        la_cycle = {"value": ev3 / (t5 - t4), "unit": UNIT, "h2d_bytes_per_step": 4 * ncell_T, "d2h_bytes_per_step": 4 * (bz + 1) * by * lx, "ms_per_step": 1e3 * (t5 - t4) / 3,
                    "what": "mskc_set_temperature(pinned int32 field) + mskc_run_parallel + mskc_download_lattice(pinned) per step"}
        del T_host, L_host

    # ---- aggregate over ranks -----------------------------------------------------------------------
    vals = torch.tensor([dt, t3 - t2, float(ev), float(ev2), float(su)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dt_max, dt2_max, ev_sum, ev2_sum, su_sum = mx[0].item(), mx[1].item(), sm[2].item(), sm[3].item(), sm[4].item()
    else:
        dt_max, dt2_max, ev_sum, ev2_sum, su_sum = dt, t3 - t2, float(ev), float(ev2), float(su)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle_binding import build_oracle
            build_oracle()
            n, s_u, cdt = oracle_sample(args, True, 3_000_000)
            n2, s_u2, cdt2 = oracle_sample(args, False, 3_000_000)
            cpu = {"value": n / cdt, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "C restatement of mulskips.e (oracle/, 1 thread, per-call 96 KiB rate-table copy as in Get_Prob.f:151-152), 480x480x96 patch of the same slab, 3e6 events, ran3",
                   "site_updates_per_s": s_u / cdt, "value_without_table_copy": n2 / cdt2, "host_cores_available": os.cpu_count()}
        out = {"metric": METRIC, "value": ev_sum / dt_max, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": 1e3 * dt_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64",
               "data": "synthetic", "config": config_dict(args, "gpu"),
               "site_updates_per_s": su_sum / dt_max, "events_per_step": ev_sum / args.steps, "active_sites_per_gpu": int(n_active),
               "device_ms_per_step": kms / args.steps, "gpu_launches": int(launches),
               "e2e": {"value": ev2_sum / dt2_max, "unit": UNIT, "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
                       "ms_per_step": 1e3 * dt2_max / args.steps, "frame_cells": nrec,
                       "what": "per step: mskc_set_rates(pinned tables) + mskc_run_parallel + mskc_extract_frame(device-compacted frame into pinned memory)", "checksum": occ},
               "e2e_la_cycle": la_cycle, "roofline": roofline, "dense_pass": dense, "cpu_baseline": cpu, "clocks": clocks}
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--box", type=int, nargs=3, default=[4608, 4608, 96])
    ap.add_argument("--len1", type=int, default=24)
    ap.add_argument("--sweeps", type=int, default=32)
    ap.add_argument("--ept", type=float, default=4.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-la-cycle", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON: whatever libraries print there (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            if args.impl == "reference":
                run_reference(args)
            else:
                run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(keep, 1)
        os.close(keep)
    lines = [l for l in buf.getvalue().splitlines() if l.strip()]
    for l in lines[:-1]:
        print(l, file=sys.stderr)
    if lines:
        print(lines[-1], flush=True)


if __name__ == "__main__":
    main()
