"""Probe of the serial-parity kernel: one regression case at its own box, N events (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from cases import make_start
from oracle_binding import Oracle, upload_oracle_state
from mulskips_b200 import Mskc
case = sys.argv[1] if len(sys.argv) > 1 else "surface"
nev = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
box = (480, 480, 480)
start = make_start(case, "/tmp/s.dat", maxiter=10**9, outmolmol=10**9, maxtime=1e30, outtime=1e30)
orc = Oracle(*box, start); orc.setup()
lc, lv = orc.species()
ctx = Mskc(*box, lc, lv)
upload_oracle_state(orc, ctx)
ctx.run_parity(event_budget=2000)
r = ctx.run_parity(event_budget=nev)
print(case, f"{r.events_done / (r.kernel_ms * 1e-3):.4g} events/s", flush=True)
