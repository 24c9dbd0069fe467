"""Quick throughput probe of the sublattice sampler on a flat Si(001) / SiC(001) slab (device-built)."""
import argparse, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from mulskips_b200 import Mskc
from mulskips_b200.workloads import rates_for

ap = argparse.ArgumentParser()
ap.add_argument("--box", type=int, nargs=3, default=[2304, 2304, 240])
ap.add_argument("--len1", type=int, default=120)
ap.add_argument("--case", default="Si-SL")
ap.add_argument("--ept", type=float, nargs="+", default=[2, 4, 8, 16])
ap.add_argument("--sweeps", type=int, default=50)
a = ap.parse_args()
w = rates_for(a.case)
ctx = Mskc(*a.box, w["listcrystal"], w["listcov"])
ctx.set_rates(w["PtransE"], w["PtransD"], w["PtransAbs"], w["PtransDes"], w["PtransZig"])
t0 = time.time(); ctx.init_flat(a.len1, walls=w["walls"], two_species=w["two_species"]); print("init", time.time() - t0, "s")
c, ms = ctx.census(); print("census ms", ms, "under", c.undercoordinated, "bulk", c.bulk, "GB/s", 4 * np.prod(a.box) / ms / 1e6)
for ept in a.ept:
    r = ctx.run_parallel(3, events_per_tile=ept, seed=1)
    r = ctx.run_parallel(a.sweeps, events_per_tile=ept, seed=2, time_kernels=True)
    print(f"ept={ept}: {r.events_done} events, {r.kernel_ms:.2f} ms, {r.events_done / r.kernel_ms * 1e3:.4g} ev/s, {r.site_updates / r.kernel_ms * 1e3:.4g} su/s, "
          f"active={r.n_active} R={r.total_rate:.4g} t={r.time:.4g} launches={r.launches} maxev={r.max_tile_events} us/sweep={r.kernel_ms * 1e3 / a.sweeps:.1f}")
c, ms = ctx.census(); print("census", c.undercoordinated, c.bulk, c.noepi_bulk, c.mean_height)
