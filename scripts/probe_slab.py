"""Slab (halo-exchange) mode on ONE GPU: the rank is its own neighbour, so the pack / apply kernels and the split colour step run
as between two GPUs.  Prints the device time per colour step; env knobs MSKC_NO_OVERLAP, MSKC_BTEAM, MSKC_SKIP_HALO select variants."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mulskips_b200 import Mskc
from mulskips_b200.workloads import rates_for
ap = argparse.ArgumentParser()
ap.add_argument("--box", type=int, nargs=3, default=[4608, 4608, 96]); ap.add_argument("--sweeps", type=int, default=32); ap.add_argument("--slab", type=int, default=1)
a = ap.parse_args()
w = rates_for("surface"); w["PtransZig"] = 0.8
ctx = Mskc(*a.box, w["listcrystal"], w["listcov"], slab=(0, 1) if a.slab else None)
ctx.set_rates(w["PtransE"], w["PtransD"], w["PtransAbs"], w["PtransDes"], w["PtransZig"])
ctx.init_flat(24, walls=False, two_species=True)
ctx.run_parallel(24, events_per_tile=4.0, seed=1)
for k in range(3):
    r = ctx.run_parallel(a.sweeps, events_per_tile=4.0, seed=2 + k)
    print(f"slab={a.slab} {' '.join(f'{k}={v}' for k, v in os.environ.items() if k.startswith('MSKC_'))}: {r.events_done} ev, {r.kernel_ms:.2f} ms, {r.events_done / r.kernel_ms * 1e3:.4g} ev/s, us/colourstep={r.kernel_ms * 1e3 / (4 * a.sweeps):.1f} launches={r.launches}", flush=True)
if os.environ.get('MSKC_SPLIT_TIMING'):
    ctx.run_parallel(2, events_per_tile=4.0, seed=9)
ctx.close()
