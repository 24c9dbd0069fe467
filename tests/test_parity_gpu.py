"""Serial-parity mode on the GPU against the CPU oracle: same seeded input, bit-exact lattice words,
event sequence end state and KMC time (integer work bit-exact, Time compared as raw bits)."""
import os
import struct

import numpy as np
import pytest

from cases import make_start
from oracle_binding import Oracle, upload_oracle_state

pytestmark = pytest.mark.gpu

# (case, box, lens, events): reduced boxes of the reference regression cases
SMALL = [
    ("Si-SL", (96, 96, 240), (216, 0, 0), 50000),
    ("surface", (96, 96, 240), (48, 0, 0), 300000),
    ("cube", (144, 144, 144), (48, 48, 48), 300000),
    ("cube-zigzag", (144, 144, 144), (48, 48, 48), 300000),
    ("sphere", (144, 144, 144), (60, 0, 0), 300000),
    ("CVD-Si-SiH2", (96, 96, 96), (48, 0, 0), 400000),
    # examples/ex5-CVD-IO/CVD-SiGeBdoped: Si/Ge/B crystal species + Cl/H coverage, 288 evaporation classes, InitSt SG
    ("SiGeBdoped", (96, 96, 240), (48, 0, 0), 300000),
]


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


@pytest.mark.parametrize("case,box,lens,nev", SMALL, ids=[c[0] for c in SMALL])
def test_parity_small(tmp_path, case, box, lens, nev):
    from mulskips_b200 import Mskc
    start = make_start(case, str(tmp_path / "start.dat"), lens=lens, maxiter=10**9, outmolmol=10**9, maxtime=1e30, outtime=1e30)
    orc = Oracle(*box, start)
    orc.setup()
    lc, lv = orc.species()
    ctx = Mskc(*box, lc, lv)
    upload_oracle_state(orc, ctx)
    assert ctx.state_hash() == orc.hash()
    done = 0
    for chunk in (1, 7, 1000, nev - 1008):
        r = ctx.run_parity(event_budget=chunk)
        n = orc.run(chunk)
        assert n == chunk and r.events_done == chunk
        done += chunk
        s = orc.sizes()
        assert (r.num_adatom, r.num_atoms, r.num_voids) == (s["NumAdAtom"], s["NumAtoms"], s["NumVoids"]), f"after {done} events"
        assert bits(r.time) == bits(orc.time), f"after {done} events"
        assert tuple(r.last_site) == orc.last_site()
        assert r.iter == s["Iter"]
    L = ctx.download_lattice()
    assert np.array_equal(L, orc.lattcoo())
    assert ctx.state_hash() == orc.hash()
    assert bits(r.tree_root) == bits(orc.tree_root)
    # lists in reference order
    d, g = orc.lists(), ctx.download_lists()
    for k in ("mc_xyz", "mc_nn", "mc_event", "void_xyz", "void_nn"):
        assert np.array_equal(d[k], g[k]), k
    assert np.array_equal(d["mc_rate"].view(np.uint64), g["mc_rate"].view(np.uint64))
    cc, cv = orc.counts()
    c = ctx.counters()
    assert list(c.CountCrystal) == list(cc) and list(c.CountCov) == list(cv)
    assert r.site_updates > 0
    ctx.close(); orc.close()


@pytest.mark.parametrize("variant", ["ramp", "fixedT", "ramp-T"])
def test_parity_tabulated_field(tmp_path, variant):
    """InitSt ST: trench builder with a temperature ramp along z, the four rate tables interpolated
    between tabulated temperatures at every site's own temperature (InterpProbFromField.f) -- on the device a
    per-kelvin table built by mskc_set_rate_field.  'fixedT' is the Ttop == Tbottom variant (one interpolation,
    SetTrenchSi.f:65-67); 'ramp-T' uses the reference's test-mode random stream (RunType T)."""
    from cases import make_st_case
    from mulskips_b200 import Mskc
    box = (96, 48, 144)
    kw = dict(ttop=1000.0, tbottom=1000.0) if variant == "fixedT" else {}
    start = make_st_case(str(tmp_path), runtype="T" if variant.endswith("-T") else "R", **kw)
    orc = Oracle(*box, start)
    orc.setup()
    lc, lv = orc.species()
    ctx = Mskc(*box, lc, lv)
    upload_oracle_state(orc, ctx)
    assert ctx.state_hash() == orc.hash()
    for chunk in (1, 999, 199000):
        r = ctx.run_parity(event_budget=chunk)
        assert orc.run(chunk) == chunk and r.events_done == chunk
        s = orc.sizes()
        assert (r.num_adatom, r.num_atoms, r.num_voids) == (s["NumAdAtom"], s["NumAtoms"], s["NumVoids"])
        assert bits(r.time) == bits(orc.time)
        assert tuple(r.last_site) == orc.last_site()
    assert np.array_equal(ctx.download_lattice(), orc.lattcoo())
    assert bits(r.tree_root) == bits(orc.tree_root)
    d, g = orc.lists(), ctx.download_lists()
    for k in ("mc_xyz", "mc_nn", "mc_event", "void_xyz", "void_nn"):
        assert np.array_equal(d[k], g[k]), k
    assert np.array_equal(d["mc_rate"].view(np.uint64), g["mc_rate"].view(np.uint64))
    cc, cv = orc.counts()
    c = ctx.counters()
    assert list(c.CountCrystal) == list(cc) and list(c.CountCov) == list(cv)
    assert cv[0] > 100 and cc[0] != 0          # adsorption and growth both happened
    ctx.close(); orc.close()


def test_parity_laser_annealing(tmp_path):
    """InitSt LA (config C4): temperature in the lattice word, rates from the host-built exp/erf tables
    (Get_Prob.f:127-149), on a synthetic nucleus that re-solidifies until the reference's
    'few MC particles: stop MC' exit.  Bit-exact lattice, Time bits, list sizes, stop reason."""
    from cases import make_la_case
    from mulskips_b200 import Mskc
    box = (96, 96, 96)
    start = make_la_case(str(tmp_path), box=box, lennuc=18.0)
    orc = Oracle(*box, start)
    orc.setup()
    lc, lv = orc.species()
    ctx = Mskc(*box, lc, lv)
    upload_oracle_state(orc, ctx)
    assert ctx.state_hash() == orc.hash()
    total = 0
    while True:
        n = orc.run(3000)
        r = ctx.run_parity(event_budget=3000)
        total += n
        assert r.events_done == n
        s = orc.sizes()
        assert (r.num_adatom, r.num_atoms, r.num_voids) == (s["NumAdAtom"], s["NumAtoms"], s["NumVoids"])
        assert bits(r.time) == bits(orc.time)
        if s["done"]:
            assert r.stop_reason == 4            # MSKC_STOP_FEWMC
            break
    assert total > 5000
    assert np.array_equal(ctx.download_lattice(), orc.lattcoo())
    # SetT.f through the ABI: re-uploading the temperature field leaves the lattice unchanged, a shifted field changes bits 9..21 only
    L = orc.lattcoo().copy()
    T = ((L[:-1] >> 9) & 8191).astype(np.int32)
    ctx.set_temperature(np.where(T == 0, 10, T))
    assert np.array_equal(ctx.download_lattice(), L)
    ctx.set_temperature(np.where(T == 0, 10, T + 1))
    L2 = ctx.download_lattice()
    assert np.array_equal(L2[:-1] & ~(8191 << 9), L[:-1] & ~(8191 << 9))
    assert np.array_equal((L2[:-1] >> 9) & 8191, np.where(T == 0, 0, T + 1))
    ctx.close(); orc.close()


@pytest.mark.parametrize("species,phase1", [("SiGe", 100000), ("Si", 20000)])
def test_parity_la_fork(tmp_path, species, phase1):
    """The LaserAnnealing fork (mulskips-source-LaserAnnealing, SURVEY 8 row a-LA) in serial-parity mode: species bits
    28-29 and event codes 1-2/3 at the boundary (mskc_set_tree_variant), per-species Tm/P0/sigma/Tflex, solidification
    rates scaled by the liquid fraction CountLiquid(k)/SUM(CountLiquid) re-evaluated before every event, Si<->Ge swap
    when a species is used up, NumAdAtomOcc bookkeeping and the NumOcc0 stop rule (la/main.f:634-726).  Phase 1: a
    melt nucleus re-solidifies (Ge segregates into the liquid); every chunk: list sizes, Time bits, NumAdAtomOcc.
    Phase 2: the stop threshold is lowered on both sides to a few atoms above the current count so that the rule fires
    while the melt is still rich (when the melt is used up the reference divides by SUM(CountLiquid) = 0).  At the end:
    lattice words in the fork's layout, lists, rates as raw bits, species counters."""
    from cases import make_la_tree_case
    from mulskips_b200 import Mskc
    box = (96, 96, 96)
    start = make_la_tree_case(str(tmp_path), box=box, sige=(species == "SiGe"), tbase=1480 if species == "SiGe" else 1655)
    orc = Oracle(*box, start, la_tree=True)
    orc.setup()
    lc, lv = orc.species()
    ctx = Mskc(*box, lc, lv)
    upload_oracle_state(orc, ctx)
    assert np.array_equal(ctx.download_lattice(), orc.lattcoo())       # round trip through the layout conversion
    total = 0
    for chunk in (1, 999, phase1 // 2 - 1000, phase1 // 2):
        n = orc.run(chunk)
        r = ctx.run_parity(event_budget=chunk)
        total += n
        s = orc.sizes()
        assert n == chunk and r.events_done == n, (r.events_done, n, total)
        assert (r.num_adatom, r.num_atoms, r.num_voids) == (s["NumAdAtom"], s["NumAtoms"], s["NumVoids"]), total
        assert bits(r.time) == bits(orc.time), total
        assert ctx.counters().num_adatom_occ == orc.fork()["NumAdAtomOcc"]
    fk = orc.fork()
    assert fk["CountLiquid"][0] > 0
    occ0 = s["NumAtoms"] + fk["NumAdAtomOcc"] + 5
    orc.set_numocc0(occ0)
    T = (orc.lattcoo()[:-1] >> 9) & 8191
    ctx.set_la_fork(fk["Tm"], fk["P0"], fk["sigma"], fk["Tflex"], int(T[T > 0].min()), int(T.max()), fk["InitCountCrystal"], occ0, fk["NumAdAtomOcc"])
    n = orc.run(200000)
    r = ctx.run_parity(event_budget=200000)
    s, fk = orc.sizes(), orc.fork()
    assert s["done"] == 1 and r.stop_reason == 5 and r.events_done == n      # MSKC_STOP_RESOLIDIFIED
    assert s["NumAtoms"] + fk["NumAdAtomOcc"] > occ0
    assert bits(r.time) == bits(orc.time)
    assert np.array_equal(ctx.download_lattice(), orc.lattcoo())
    d, g = orc.lists(), ctx.download_lists()
    for k in ("mc_xyz", "mc_nn", "mc_event", "void_xyz", "void_nn"):
        assert np.array_equal(d[k], g[k]), k
    assert np.array_equal(d["mc_rate"].view(np.uint64), g["mc_rate"].view(np.uint64))
    cc, cv = orc.counts()
    c = ctx.counters()
    assert list(c.CountCrystal) == list(cc) and c.num_adatom_occ == fk["NumAdAtomOcc"]
    if species == "SiGe":
        assert cc[1] > 0
    ctx.close(); orc.close()
