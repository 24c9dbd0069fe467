"""events per active tile per colour step (m) under the present tau rule, as a function of the box size"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mulskips_b200 import Mskc
from mulskips_b200.workloads import rates_for
w = rates_for("surface"); w["PtransZig"] = 0.8
for bx in (96, 240, 480, 1152, 4608):
    ctx = Mskc(bx, bx, 96, w["listcrystal"], w["listcov"])
    ctx.set_rates(w["PtransE"], w["PtransD"], w["PtransAbs"], w["PtransDes"], w["PtransZig"])
    ctx.init_flat(24, walls=False, two_species=True)
    nt = (bx // 24) ** 2
    ctx.run_parallel(96, events_per_tile=float(os.environ.get("EPT", 4)), seed=1)
    ms, taus, dts = [], [], []
    t_prev = None
    for k in range(12):
        r = ctx.run_parallel(32, events_per_tile=float(os.environ.get("EPT", 4)), seed=10 + k)
        ms.append(r.events_done / (32 * nt)); taus.append(r.tau)
        if t_prev is not None: dts.append((r.time - t_prev) / 32)
        t_prev = r.time
    print(f"box {bx}: tiles {nt}, events per tile per colour step m = {np.mean(ms):.2f} (min {min(ms):.2f} max {max(ms):.2f}), mean dt/sweep {np.mean(dts):.4g}, tau at call ends: median {np.median(taus):.4g} min {min(taus):.4g} max {max(taus):.4g}; R/tile median {4 / np.median(taus):.4g}", flush=True)
    ctx.close()
