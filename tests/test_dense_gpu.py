"""The dense passes and the device-side builder against the CPU oracle.

* mskc_init_flat (the builder every bench lattice comes from) == the oracle's F / FS / Fw initialisers
  (initial/SetSiC3C.f:39-146, initial/SetSi.f:49-209): every lattice word, the bonds of every listed site, the
  active set with (event class, rate bits), the list sizes.
* mskc_rerate_all (the dense event-rate kernel, initial/SetCAD.f:301-418 + Get_Prob.f:54-331): after N replayed events it
  reproduces the incrementally maintained (class, rate bits) of every active site exactly; after a temperature upload
  it equals a context that lived at the new temperature all along (initial rates from the oracle at that temperature).
* mskc_census_lattice == the counters of the reference's frame writer (inpout/WriteMolMolXYZFile.f:77-119), read off
  the headers of the oracle's I*.xyz / _d.xyz files.
"""
import os

import numpy as np
import pytest

from cases import make_la_case, make_start
from oracle_binding import Oracle, trace_arrays, upload_oracle_state
from test_replay_gpu import expected_active

pytestmark = pytest.mark.gpu


def active_dict(ctx):
    xyz, ev, rate = ctx.download_active()
    return {tuple(int(v) for v in p): (int(e), np.float64(r).view(np.uint64)) for p, e, r in zip(xyz, ev, rate)}


# (oracle case, InitSt, walls, two_species): the two reference builders mskc_init_flat stands for
FLAT = [("surface", "F", False, True), ("Si-SL", "FS", True, False)]


@pytest.mark.parametrize("case,initst,walls,two", FLAT, ids=[f[1] for f in FLAT])
@pytest.mark.parametrize("len1", [48, 120])
def test_init_flat_matches_oracle(tmp_path, case, initst, walls, two, len1):
    from mulskips_b200 import Mskc
    from mulskips_b200.workloads import rates_for
    box = (96, 96, 240)
    start = make_start(case, str(tmp_path / "start.dat"), lens=(len1, 0, 0), maxiter=10**9, outmolmol=10**9, initst=initst)
    orc = Oracle(*box, start)
    orc.setup()
    lc, lv = orc.species()
    ctx = Mskc(*box, lc, lv)
    w = rates_for(case)
    ctx.set_rates(w["PtransE"], w["PtransD"], w["PtransAbs"], w["PtransDes"], w["PtransZig"])
    ctx.init_flat(len1, walls=walls, two_species=two)
    L, Lo = ctx.download_lattice(), orc.lattcoo()
    assert np.array_equal(L, Lo), f"{np.count_nonzero(L != Lo)} lattice words differ"
    exp, d = expected_active(orc)
    got = active_dict(ctx)
    assert set(got) == set(exp)
    assert got == exp
    for key in ("mc", "void", "bulk"):
        x, nn = d[f"{key}_xyz"], d[f"{key}_nn"]
        if len(x):
            assert np.array_equal(ctx.download_bonds(x), nn), key
    c, s = ctx.counters(), orc.sizes()
    assert (c.num_atoms, c.num_voids) == (s["NumAtoms"], s["NumVoids"])
    nsites, bad_flags, bad_coor, _ = ctx.check_flags()
    assert nsites == s["NumAtoms"] + s["NumAdAtom"] + s["NumVoids"] and bad_flags == 0 and bad_coor == 0
    # and the sampler runs from it
    r = ctx.run_parallel(n_sweeps=5, events_per_tile=2.0, seed=11)
    assert r.events_done > 0 and r.n_active > 0
    ctx.close(); orc.close()


@pytest.mark.parametrize("case,box,lens,nev", [("surface", (96, 96, 240), (48, 0, 0), 60000), ("CVD-Si-SiH2", (96, 96, 96), (48, 0, 0), 100000),
                                               ("cube-zigzag", (144, 144, 144), (48, 48, 48), 60000)], ids=["surface", "CVD", "cube-zigzag"])
def test_rerate_all_reproduces_incremental_rates(tmp_path, case, box, lens, nev):
    from mulskips_b200 import Mskc
    start = make_start(case, str(tmp_path / "start.dat"), lens=lens, maxiter=10**9, outmolmol=10**9, maxtime=1e30, outtime=1e30)
    orc = Oracle(*box, start)
    orc.setup()
    lc, lv = orc.species()
    ctx = Mskc(*box, lc, lv)
    upload_oracle_state(orc, ctx)
    orc.set_trace(nev)
    assert orc.run(nev) == nev
    site, code, arm, _ = trace_arrays(orc)
    ctx.replay_events(site, code, arm, tree_root=1.0)
    before = active_dict(ctx)
    exp, d = expected_active(orc)
    assert set(before) == set(exp)
    n, total, ms = ctx.rerate_all()
    after = active_dict(ctx)
    assert n == len(before) and set(after) == set(before)
    Lo = orc.lattcoo()
    word = lambda s: int(Lo[s[2], s[1], s[0]])
    iso = {s for s in before if (word(s) >> 23) & 1 and ((word(s) >> 24) & 7) == 0}        # 100000*Tree(1): root-dependent
    # Rates the reference holds are those of the LAST time it touched a site.  One class of them is history dependent: an
    # empty 1-coordinated site is rated against the coordination its single neighbour had at that moment (Get_Prob.f:
    # 156-184: < 2 -> code 1, rate 1.e-8), and that neighbour changes later without the site being re-rated.  A dense
    # re-rate evaluates the rule on the lattice as it is now; everything else must come out bit-identical.
    t = orc.tables()
    pdep1 = (t["PtransD"][0] + t["PtransD"][1] + t["PtransD"][2]) + ((0.0 + t["PtransAbs"][3]) + t["PtransAbs"][4] + t["PtransAbs"][5])
    first = {tuple(int(v) for v in x): tuple(int(v) for v in nn[0]) for x, nn in zip(d["mc_xyz"], d["mc_nn"])}
    hist = set()
    for s in before:
        w = word(s)
        if not (w >> 23) & 1 and ((w >> 24) & 7) == 1 and (w & 255) == 0:
            hist.add(s)
            nb = (word(first[s]) >> 24) & 7
            want = (100, np.float64(pdep1).view(np.uint64)) if nb >= 2 else (1, np.float64(np.float32(1e-8)).view(np.uint64))
            assert after[s] == want, (s, after[s], want)
    diff = [s for s in before if s not in iso and s not in hist and after[s] != before[s]]
    assert not diff, diff[:5]
    # ... and against the oracle's own rates (the wrong-list quirk of Get_Prob.f:266-269 is the only other difference)
    bad = [s for s in before if s not in iso and s not in hist and after[s] != exp[s]]
    assert len(bad) <= orc.quirks()["wl_cov"], bad[:5]
    assert len(hist) > 0
    assert abs(total - sum(np.uint64(v[1]).view(np.float64) for v in after.values())) <= 1e-9 * total
    assert ctx.check_flags()[1] == 0
    ctx.close(); orc.close()


def test_rerate_all_after_temperature_upload(tmp_path):
    """LA rate law: context A lives at the regression-like map T(x,y,z), replays the oracle's trace, then receives the map
    T+1 (mskc_set_temperature) and is re-rated; context B starts from the oracle built at T+1 (its initial rates are the
    oracle's Get_Prob_Ini at T+1) and replays the same trace.  Same lattice words, same (class, rate bits) everywhere."""
    from mulskips_b200 import Mskc
    box = (96, 96, 96)
    da, db = tmp_path / "a", tmp_path / "b"
    da.mkdir(); db.mkdir()
    start_a = make_la_case(str(da), box=box, lennuc=18.0)
    start_b = make_la_case(str(db), box=box, lennuc=18.0)
    Ta = np.loadtxt(da / "TempStat.dat", dtype=np.int32).reshape(box[2], box[1], box[0])
    with open(db / "TempStat.dat", "w") as f:
        for z in range(box[2]):
            f.write(" ".join(str(v) for v in (Ta[z].ravel() + 1)) + "\n")
    oa, ob = Oracle(*box, start_a), Oracle(*box, start_b)
    oa.setup(); ob.setup()
    lc, lv = oa.species()
    ca, cb = Mskc(*box, lc, lv), Mskc(*box, lc, lv)
    upload_oracle_state(oa, ca, trange=(1600, 1700))
    upload_oracle_state(ob, cb, trange=(1600, 1700))
    oa.set_trace(10**6)
    n = oa.run(10**6)
    assert n > 5000
    site, code, arm, _ = trace_arrays(oa)
    ca.replay_events(site[:n], code[:n], arm[:n])
    cb.replay_events(site[:n], code[:n], arm[:n])
    assert active_dict(ca) != active_dict(cb)                       # the rates do depend on the temperature
    # the new field, [LenZ][LenY*LenX] as SetT.f reads it: what context B's lattice words hold (10 = leave the cell alone)
    Tb = (cb.download_lattice()[:-1] >> 9) & 8191
    assert Tb.max() == Ta.max() + 1
    ca.set_temperature(np.where(Tb > 0, Tb, 10).astype(np.int32))
    nre, total, ms = ca.rerate_all()
    assert np.array_equal(ca.download_lattice(), cb.download_lattice())
    a, b = active_dict(ca), active_dict(cb)
    assert nre == len(a) and set(a) == set(b)
    assert a == b
    for c in (ca, cb):
        c.close()
    oa.close(); ob.close()


@pytest.mark.parametrize("case,box,lens,nev", [("cube-zigzag", (144, 144, 144), (48, 48, 48), 150000), ("CVD-Si-SiH2", (96, 96, 96), (48, 0, 0), 150000),
                                               ("Si-SL", (96, 96, 240), (216, 0, 0), 60000)], ids=["cube-zigzag", "CVD", "Si-SL"])
def test_census_matches_frame_writer(tmp_path, case, box, lens, nev):
    """the counters of WriteMolMolXYZFile.f:77-119 -- atoms with coordination <= 3 and coverage sites in planes 1..LenZ-1
    (the .xyz header counts them plus the vacancies), non-epitaxial 4-coordinated atoms in planes 0..LenZ-1 (the _d.xyz
    header) -- from the oracle's frame files against the device census"""
    from mulskips_b200 import Mskc
    start = make_start(case, str(tmp_path / "start.dat"), lens=lens, maxiter=10**9, outmolmol=10**9, maxtime=1e30, outtime=1e30)
    out = tmp_path / "frames"
    out.mkdir()
    orc = Oracle(*box, start, outdir=str(out), write_files=True)
    orc.setup()
    assert orc.run(nev) == nev
    lc, lv = orc.species()
    ctx = Mskc(*box, lc, lv)
    upload_oracle_state(orc, ctx)
    orc.write_frame(7)
    n_xyz = int(open(out / "I00000007.xyz").readline().split()[0])
    n_d = int(open(out / "I00000007_d.xyz").readline().split()[0])
    cs, ms = ctx.census()
    s = orc.sizes()
    assert cs.undercoordinated + cs.coverage + s["NumVoids"] == n_xyz
    assert cs.noepi_bulk == n_d
    L = orc.lattcoo()
    occ, coor, cov = (L >> 23) & 1, (L >> 24) & 7, L & 255
    assert cs.walls == int(np.count_nonzero(cov == 113))
    assert cs.occupied == int(np.count_nonzero((occ == 1) & (cov != 113)))
    assert cs.bulk == int(np.count_nonzero((occ == 1) & (cov != 113) & (coor == 4)))
    ctx.close(); orc.close()


def test_temperature_u16_and_phases(tmp_path):
    """mskc_set_temperature_u16 writes the same lattice words as the int32 form (initial/SetT.f:36-46: bits 9-21, value 10 = wall,
    untouched), and mskc_download_phases is the socket loop's phase array (la/main.f:806-814) computed on the device"""
    from mulskips_b200 import Mskc
    box = (96, 96, 96)
    start = make_la_case(str(tmp_path), box=box, lennuc=18.0)
    orc = Oracle(*box, start)
    orc.setup()
    orc.run(3000)
    lc, lv = orc.species()
    a, b = Mskc(*box, lc, lv), Mskc(*box, lc, lv)
    upload_oracle_state(orc, a, trange=(1500, 1800))
    upload_oracle_state(orc, b, trange=(1500, 1800))
    rng = np.random.default_rng(5)
    T = (1600 + rng.integers(0, 150, size=(box[2], box[1] * box[0]))).astype(np.int32)
    T[rng.random(T.shape) < 0.05] = 10                               # "wall": the cell keeps its temperature bits
    a.set_temperature(T)
    b.set_temperature_u16(T.astype(np.uint16))
    La, Lb = a.download_lattice(), b.download_lattice()
    assert np.array_equal(La, Lb)
    want = np.where(T.reshape(box[2], box[1], box[0]) == 10, (orc.lattcoo()[:-1] >> 9) & 8191, T.reshape(box[2], box[1], box[0]))
    assert np.array_equal((La[:-1] >> 9) & 8191, want)
    L = La[:-1]
    occ, coor, cov = (L >> 23) & 1, (L >> 24) & 7, L & 255
    ph = np.where(occ == 1, np.where(cov == 113, 10, coor), 0).astype(np.uint8)
    assert np.array_equal(a.download_phases(), ph)
    assert a.last_dense_ms() >= 0.0
    a.close(); b.close(); orc.close()
