"""Probe of the sublattice sampler on the bench workload (SiC flat slab): per-team-size timing, tail statistics."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mulskips_b200 import Mskc
from mulskips_b200.workloads import rates_for

ap = argparse.ArgumentParser()
ap.add_argument("--box", type=int, nargs=3, default=[4608, 4608, 96])
ap.add_argument("--len1", type=int, default=24)
ap.add_argument("--case", default="surface")
ap.add_argument("--zig", type=float, default=0.8)
ap.add_argument("--ept", type=float, nargs="+", default=[8])
ap.add_argument("--sweeps", type=int, default=16)
ap.add_argument("--warm", type=int, default=3)
ap.add_argument("--tiles", action="store_true")
a = ap.parse_args()
w = rates_for(a.case); w["PtransZig"] = a.zig
ctx = Mskc(*a.box, w["listcrystal"], w["listcov"])
ctx.set_rates(w["PtransE"], w["PtransD"], w["PtransAbs"], w["PtransDes"], w["PtransZig"])
ctx.init_flat(a.len1, walls=w["walls"], two_species=w["two_species"])
for ept in a.ept:
    ctx.run_parallel(a.warm, events_per_tile=ept, seed=1)
    r = ctx.run_parallel(a.sweeps, events_per_tile=ept, seed=2, time_kernels=True)
    print(f"TEAM={os.environ.get('MSKC_TEAM','def')} ept={ept}: {r.events_done} ev, {r.kernel_ms:.2f} ms, {r.events_done / r.kernel_ms * 1e3:.4g} ev/s, su/ev={r.site_updates / max(r.events_done,1):.2f} "
          f"active={r.n_active} maxev={r.max_tile_events} us/colourstep={r.sweep_kernel_ms * 1e3 / (4 * a.sweeps):.1f} ev/tile/step={r.events_done / (a.sweeps * 4 * (a.box[0] // 48) * (a.box[1] // 48)):.2f}", flush=True)
if a.tiles:
    ts = ctx.tile_stats()
    print("per-tile events percentiles (50,90,99,99.9,100):", np.percentile(ts["events"], [50, 90, 99, 99.9, 100]), "maxev", ts["maxev"].max(), "maxev pct", np.percentile(ts["maxev"], [50, 90, 99, 100]))
ctx.close()
