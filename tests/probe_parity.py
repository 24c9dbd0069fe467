"""Probe of the serial-parity kernel: one reduced regression case, N events in one launch (for ncu / timing).
Lives under tests/ because it takes its initial state from the oracle (test infrastructure)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import make_start
from mulskips_b200 import Mskc
from oracle_binding import Oracle, upload_oracle_state
case = sys.argv[1] if len(sys.argv) > 1 else "surface"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
box, lens = ((96, 96, 240), (48, 0, 0)) if case in ("surface", "Si-SL", "CVD-Si-SiH2") else ((144, 144, 144), (48, 48, 48))
with tempfile.TemporaryDirectory() as td:
    start = make_start(case, os.path.join(td, "start.dat"), lens=lens, maxiter=10**9, outmolmol=10**9, maxtime=1e30, outtime=1e30)
    orc = Oracle(*box, start); orc.setup()
    lc, lv = orc.species()
    ctx = Mskc(*box, lc, lv)
    upload_oracle_state(orc, ctx)
    ctx.run_parity(event_budget=2000)
    r = ctx.run_parity(event_budget=n)
    print(f"{case}: {r.events_done} events, {r.kernel_ms:.1f} ms, {r.events_done / r.kernel_ms * 1e3:.4g} events/s, {r.site_updates / r.events_done:.2f} site-updates/event, NumAdAtom {r.num_adatom}")
