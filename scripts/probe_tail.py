import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mulskips_b200 import Mskc
from mulskips_b200.workloads import rates_for
case = sys.argv[1] if len(sys.argv) > 1 else "surface"
w = rates_for(case); 
box = (2304, 2304, 96)
ctx = Mskc(*box, w["listcrystal"], w["listcov"])
ctx.set_rates(w["PtransE"], w["PtransD"], w["PtransAbs"], w["PtransDes"], w["PtransZig"])
ctx.init_flat(24 if case == "surface" else 48, walls=w["walls"], two_species=w["two_species"])
r = ctx.run_parallel(16, events_per_tile=8.0, seed=3)
ts = ctx.tile_stats()
print("events", r.events_done, "tiles", len(ts), "mean ev/tile", ts["events"].mean(), "max", ts["events"].max(), "maxev", ts["maxev"].max())
print("hist of per-tile total events (16 sweeps):", np.percentile(ts["events"], [50, 90, 99, 99.9, 100]))
o = np.argsort(-ts["events"])[:8]
for i in o:
    print(i, ts[i])
