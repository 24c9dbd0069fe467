"""A/B of sampler settings (per-tile event cap) on the same evolving state, one process."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mulskips_b200 import Mskc
from mulskips_b200.workloads import rates_for
ap = argparse.ArgumentParser()
ap.add_argument("--box", type=int, nargs=3, default=[4608, 4608, 96])
ap.add_argument("--ept", type=float, default=4)
ap.add_argument("--sweeps", type=int, default=32)
ap.add_argument("--warm", type=int, default=16)
ap.add_argument("--modes", nargs="+", default=["16", "12", "16", "12"], help="MSKC_EVCAP values, one timed call each")
a = ap.parse_args()
w = rates_for("surface"); w["PtransZig"] = 0.8
ctx = Mskc(*a.box, w["listcrystal"], w["listcov"])
ctx.set_rates(w["PtransE"], w["PtransD"], w["PtransAbs"], w["PtransDes"], w["PtransZig"])
ctx.init_flat(24, walls=w["walls"], two_species=w["two_species"])
ctx.run_parallel(a.warm, events_per_tile=a.ept, seed=1)
for i, m in enumerate(a.modes):
    os.environ["MSKC_EVCAP"] = m
    r = ctx.run_parallel(a.sweeps, events_per_tile=a.ept, seed=2 + i)
    print(f"evcap={m}: {r.events_done} ev, {r.kernel_ms:.2f} ms total, sampler kernels {r.sweep_kernel_ms:.2f} ms in {r.sampler_launches} launches, {r.events_done / r.kernel_ms * 1e3:.4g} ev/s, "
          f"active={r.n_active} maxev={r.max_tile_events} ev/tile/step={r.events_done / (a.sweeps * 4 * (a.box[0] // 48) * (a.box[1] // 48)):.2f} t={r.time:.5g}", flush=True)
ctx.close()
