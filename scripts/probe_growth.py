import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mulskips_b200 import Mskc
from mulskips_b200.workloads import rates_for
w = rates_for("surface"); w["PtransZig"] = float(sys.argv[1]) if len(sys.argv) > 1 else 0.8
box = (1152, 1152, 96)
ctx = Mskc(*box, w["listcrystal"], w["listcov"])
ctx.set_rates(w["PtransE"], w["PtransD"], w["PtransAbs"], w["PtransDes"], w["PtransZig"])
ctx.init_flat(24, walls=False, two_species=True)
nlayer = 2 * (box[0] // 12) ** 2
tot = 0
for i in range(12):
    r = ctx.run_parallel(16, events_per_tile=8.0, seed=10 + i)
    tot += r.events_done
    c, ms = ctx.census()
    L = ctx.download_lattice()
    occ = ((L >> 23) & 1).astype(bool) & ((L & 255) != 113)
    zs = np.nonzero(occ.any(axis=(1, 2)))[0]
    anyw = np.nonzero((L != 0).any(axis=(1, 2)))[0]
    print(f"step {i}: events/site-per-layer {tot / nlayer:.1f} t={r.time:.4g} under={c.undercoordinated} bulk={c.bulk} noepi={c.noepi_bulk} mean_h={c.mean_height:.2f} "
          f"rms={np.sqrt(max(c.sq_height - c.mean_height ** 2, 0)):.2f} max_occ_z={zs.max()} max_nonzero_z={anyw.max()} dep={list(r.count_dep)} evap={r.count_evap}")
