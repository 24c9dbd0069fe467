"""The sampler's warp-cooperative transition code (csrc/mskc_coop.cuh) against the CPU oracle.

The oracle runs N events of the reference loop and records its event trace (site, event code, the
IF_Arm_Choice outcome).  The same events are replayed on the GPU through the code path the
synchronous-sublattice sampler uses for every event (mskc_replay_events).  Bit-exact afterwards:
every lattice word, the bond vectors of every listed site (active, vacancy and bulk lists), and the
(event class, rate) of every active site -- the oracle pre-draws the deposited species
(Get_Prob.f:164-179) where the sampler draws it when the event fires, so a pre-drawn code is compared
as class 100 with the same total rate."""
import numpy as np
import pytest

from cases import make_start
from oracle_binding import Oracle, trace_arrays, upload_oracle_state

pytestmark = pytest.mark.gpu

CASES = [
    ("Si-SL", (96, 96, 240), (216, 0, 0), 60000, {}),
    ("surface", (96, 96, 240), (48, 0, 0), 200000, {}),
    ("surface-zig0.5", (96, 96, 240), (48, 0, 0), 200000, dict(ptranszig=0.5, runtype="R")),
    ("cube-zigzag", (144, 144, 144), (48, 48, 48), 200000, {}),
    ("sphere", (144, 144, 144), (60, 0, 0), 100000, {}),
    ("CVD-Si-SiH2", (96, 96, 96), (48, 0, 0), 300000, {}),
    ("SiGeBdoped", (96, 96, 240), (48, 0, 0), 200000, {}),       # 3 crystal + 2 coverage species (ex5-CVD-IO/CVD-SiGeBdoped)
]


def expected_active(orc):
    """{site: (class, rate bits)} from the oracle's MC list, in the sampler's conventions."""
    d = orc.lists()
    out = {}
    for xyz, ev, rate in zip(d["mc_xyz"], d["mc_event"], d["mc_rate"]):
        forced = ev == 1 and rate in (float(np.float32(1e-8)), float(np.float32(1e-6)))
        cls = int(ev) if (ev in (4, 5) or forced) else 100
        out[tuple(int(v) for v in xyz)] = (cls, np.float64(rate).view(np.uint64))
    return out, d


def check_against_oracle(orc, ctx):
    """lattice words, active set with (class, rate bits), bonds of every listed site, counters, census flags"""
    L, Lo = ctx.download_lattice(), orc.lattcoo()
    assert np.array_equal(L, Lo), f"{np.count_nonzero(L != Lo)} lattice words differ"
    # active sites: same set, same class, same rate bits (isolated atoms carry 100000*Tree(1): skipped)
    exp, d = expected_active(orc)
    xyz, ev, rate = ctx.download_active()
    got = {tuple(int(v) for v in p): (int(e), np.float64(r).view(np.uint64)) for p, e, r in zip(xyz, ev, rate)}
    assert set(got) == set(exp), f"active sets differ: {len(set(got) ^ set(exp))} sites"
    q = orc.quirks()
    bad = []
    for s, (cls, rb) in exp.items():
        w = int(Lo[s[2], s[1], s[0]])
        if (w >> 23) & 1 and ((w >> 24) & 7) == 0:
            continue
        if got[s] != (cls, rb):
            bad.append((s, got[s], (cls, rb)))
    # Get_Prob.f:266-269 reads a coverage neighbour on a 4-coordinated empty site from the wrong list; parity mode
    # reproduces that lookup, the sampler counts no such neighbour.  Only rates of the affected sites may differ.
    assert len(bad) <= q["wl_cov"], bad[:5]
    # bonds of every listed site
    for key in ("mc", "void", "bulk"):
        x, nn = d[f"{key}_xyz"], d[f"{key}_nn"]
        if len(x):
            assert np.array_equal(ctx.download_bonds(x), nn), key
    # counters
    cc, cv = orc.counts()
    c = ctx.counters()
    s = orc.sizes()
    assert (c.num_atoms, c.num_voids) == (s["NumAtoms"], s["NumVoids"])
    assert list(c.CountCrystal) == list(cc) and list(c.CountCov) == list(cv)
    # the incrementally maintained neighbour census of every site equals a fresh census of the final lattice
    nsites, bad_flags, bad_coor, _ = ctx.check_flags()
    assert nsites > 0 and bad_flags == 0, (nsites, bad_flags, bad_coor)


@pytest.mark.parametrize("case,box,lens,nev,kw", CASES, ids=[c[0] for c in CASES])
def test_replay(tmp_path, case, box, lens, nev, kw):
    from mulskips_b200 import Mskc
    base = case.split("-zig")[0] if "-zig0" in case else case
    start = make_start(base, str(tmp_path / "start.dat"), lens=lens, maxiter=10**9, outmolmol=10**9, maxtime=1e30, outtime=1e30, **kw)
    orc = Oracle(*box, start)
    orc.setup()
    lc, lv = orc.species()
    ctx = Mskc(*box, lc, lv)
    upload_oracle_state(orc, ctx)
    orc.set_trace(nev)
    assert orc.run(nev) == nev
    site, code, arm, tree1 = trace_arrays(orc)
    assert len(code) == nev
    done = 0
    for chunk in (1, 9, 990, nev - 1000):
        ctx.replay_events(site[done:done + chunk], code[done:done + chunk], arm[done:done + chunk], tree_root=1.0)
        done += chunk
    check_against_oracle(orc, ctx)
    ctx.close(); orc.close()


# the instantiation the sampler launches -- coop_event<8, FEAT>, four teams of a warp in lock step, FEAT = 0 for the
# PVD cases (rates from the incremental census flags), F_COV for the CVD cases (full census + flag cross-check) -- on
# 192 x 192 boxes, where consecutive trace events are often far enough apart to share a lock-step iteration
TEAM_CASES = [
    ("Si-SL", (192, 192, 240), (216, 0, 0), 120000, {}),
    ("surface", (192, 192, 240), (48, 0, 0), 300000, {}),
    ("surface-zig0.5", (192, 192, 240), (48, 0, 0), 300000, dict(ptranszig=0.5, runtype="R")),
    ("cube-zigzag", (192, 192, 144), (48, 48, 48), 200000, {}),
    ("CVD-Si-SiH2", (192, 192, 96), (48, 0, 0), 300000, {}),
    ("SiGeBdoped", (192, 192, 240), (48, 0, 0), 200000, {}),
]


@pytest.mark.parametrize("case,box,lens,nev,kw", TEAM_CASES, ids=[c[0] for c in TEAM_CASES])
def test_replay_teams(tmp_path, case, box, lens, nev, kw):
    from mulskips_b200 import Mskc
    base = case.split("-zig")[0] if "-zig0" in case else case
    start = make_start(base, str(tmp_path / "start.dat"), lens=lens, maxiter=10**9, outmolmol=10**9, maxtime=1e30, outtime=1e30, **kw)
    orc = Oracle(*box, start)
    orc.setup()
    lc, lv = orc.species()
    ctx = Mskc(*box, lc, lv)
    upload_oracle_state(orc, ctx)
    orc.set_trace(nev)
    assert orc.run(nev) == nev
    site, code, arm, tree1 = trace_arrays(orc)
    done, iters, shared = 0, 0, 0
    for chunk in (1, 7, 992, nev - 1000):
        a, b = ctx.replay_teams(site[done:done + chunk], code[done:done + chunk], arm[done:done + chunk], tree_root=1.0)
        iters += a; shared += b
        done += chunk
    assert iters < nev and shared > nev // 50, (iters, shared)        # teams really ran side by side
    check_against_oracle(orc, ctx)
    ctx.close(); orc.close()


@pytest.mark.parametrize("case,box,lens,nev,kw", [TEAM_CASES[2], TEAM_CASES[4]], ids=[TEAM_CASES[2][0], TEAM_CASES[4][0]])
def test_replay_teams_16_lanes(tmp_path, monkeypatch, case, box, lens, nev, kw):
    """the same through two 16-lane teams: the instantiation mskc_run_parallel picks when a colour has 2369..4736 tiles (the
    strong-scaling slabs at two GPUs) -- hexagonal-phase SiC and Si CVD with coverage"""
    monkeypatch.setenv("MSKC_REPLAY_TEAM", "16")
    test_replay_teams(tmp_path, case, box, lens, nev, kw)


def test_replay_laser_annealing(tmp_path):
    """the cooperative transition code with the LA rate law (temperature-indexed tables, T kept in erased words)"""
    from cases import make_la_case
    from mulskips_b200 import Mskc
    box = (96, 96, 96)
    start = make_la_case(str(tmp_path), box=box, lennuc=18.0)
    orc = Oracle(*box, start)
    orc.setup()
    lc, lv = orc.species()
    ctx = Mskc(*box, lc, lv)
    upload_oracle_state(orc, ctx)
    orc.set_trace(10**6)
    n = orc.run(10**6)
    assert n > 5000
    site, code, arm, tree1 = trace_arrays(orc)
    ctx.replay_events(site[:n], code[:n], arm[:n])
    assert np.array_equal(ctx.download_lattice(), orc.lattcoo())
    exp, d = expected_active(orc)
    xyz, ev, rate = ctx.download_active()
    got = {tuple(int(v) for v in p): (int(e), np.float64(r).view(np.uint64)) for p, e, r in zip(xyz, ev, rate)}
    assert got == exp
    ctx.close(); orc.close()


def test_replay_tabulated_field(tmp_path):
    """the cooperative transition code with the 'ST' per-kelvin tables (all four tables temperature dependent, the
    site temperature consulted for coverage sites too, T kept in erased words)"""
    from cases import make_st_case
    from mulskips_b200 import Mskc
    box = (96, 48, 144)
    start = make_st_case(str(tmp_path))
    orc = Oracle(*box, start)
    orc.setup()
    lc, lv = orc.species()
    ctx = Mskc(*box, lc, lv)
    upload_oracle_state(orc, ctx)
    n = 150000
    orc.set_trace(n)
    assert orc.run(n) == n
    site, code, arm, tree1 = trace_arrays(orc)
    ctx.replay_events(site, code, arm, tree_root=1.0)
    Lo = orc.lattcoo()
    assert np.array_equal(ctx.download_lattice(), Lo)
    exp, d = expected_active(orc)
    xyz, ev, rate = ctx.download_active()
    got = {tuple(int(v) for v in p): (int(e), np.float64(r).view(np.uint64)) for p, e, r in zip(xyz, ev, rate)}
    assert set(got) == set(exp)
    bad = [s for s, v in exp.items() if got[s] != v and not ((int(Lo[s[2], s[1], s[0]]) >> 23) & 1 and ((int(Lo[s[2], s[1], s[0]]) >> 24) & 7) == 0)]
    assert len(bad) <= orc.quirks()["wl_cov"], bad[:5]
    # the sampler itself runs on the temperature-dependent tables
    p = ctx.run_parallel(n_sweeps=20, events_per_tile=2.0, seed=3)
    assert p.events_done > 0
    ctx.close(); orc.close()


def test_extract_frame_matches_lattice(tmp_path):
    """mskc_extract_frame (one streaming pass on the device) returns exactly the cells the reference's frame
    writers print, in z,y,x order, with their lattice words."""
    from mulskips_b200 import Mskc
    box = (144, 144, 144)
    start = make_start("cube-zigzag", str(tmp_path / "start.dat"), lens=(48, 48, 48), maxiter=10**9, outmolmol=10**9)
    orc = Oracle(*box, start)
    orc.setup()
    orc.run(150000)
    lc, lv = orc.species()
    ctx = Mskc(*box, lc, lv)
    upload_oracle_state(orc, ctx)
    cells, words, ms = ctx.extract_frame()
    L = orc.lattcoo().reshape(-1).astype(np.uint32)
    occ, coor, cov = (L >> 23) & 1, (L >> 24) & 7, L & 255
    idx = np.arange(L.size)
    z, r = idx // (box[0] * box[1]), idx % (box[0] * box[1]); y, x = r // box[0], r % box[0]
    x1, y1, z1 = x // 3, y // 3, z // 3
    cubic = (x % 3 == 0) & (y % 3 == 0) & (z % 3 == 0) & ((x1 & 1) == (y1 & 1)) & ((x1 & 1) == (z1 & 1)) & np.isin((x1 + y1 + z1) & 3, (0, 3))
    want = (L != 0) & (cov != 113) & (((occ == 1) & (coor <= 3)) | ((occ == 1) & (coor == 4) & ~cubic) | ((occ == 0) & (cov != 0)))
    assert np.array_equal(cells, idx[want].astype(np.uint32))
    assert np.array_equal(words, L[want])
    assert len(cells) > 100
    ctx.close(); orc.close()


def test_replay_la_fork(tmp_path):
    """the cooperative transition code under the LaserAnnealing fork's law: the liquid-fraction factor of the
    solidification rates follows the trace event by event (replay kernel), event codes and lattice words cross the
    boundary in the fork's layout"""
    from cases import make_la_tree_case
    from mulskips_b200 import Mskc
    box = (96, 96, 96)
    start = make_la_tree_case(str(tmp_path), box=box)
    orc = Oracle(*box, start, la_tree=True)
    orc.setup()
    lc, lv = orc.species()
    ctx = Mskc(*box, lc, lv)
    upload_oracle_state(orc, ctx)
    n = 120000
    orc.set_trace(n)
    assert orc.run(n) == n
    site, code, arm, tree1 = trace_arrays(orc)
    assert set(np.unique(code)) <= {1, 2, 3}
    ctx.replay_events(site, code, arm, tree_root=1.0)
    Lo = orc.lattcoo()
    assert np.array_equal(ctx.download_lattice(), Lo)
    cc, _ = orc.counts()
    assert list(ctx.counters().CountCrystal) == list(cc)
    # every listed site: same class; evaporation rates bit-equal.  Deposition sites carry the liquid fraction of the
    # moment they were last rated on both sides, so those are bit-equal too.
    d = orc.lists()
    exp = {}
    for xyz, ev, rate in zip(d["mc_xyz"], d["mc_event"], d["mc_rate"]):
        forced = ev == 1 and rate in (float(np.float32(1e-8)), float(np.float32(1e-6)))
        exp[tuple(int(v) for v in xyz)] = (int(ev) if (ev == 3 or forced) else 100, np.float64(rate).view(np.uint64))
    xyz, ev, rate = ctx.download_active()
    got = {tuple(int(v) for v in p): (int(e), np.float64(r).view(np.uint64)) for p, e, r in zip(xyz, ev, rate)}
    assert set(got) == set(exp)
    bad = [s for s, v in exp.items() if got[s] != v and not ((int(Lo[s[2], s[1], s[0]]) >> 23) & 1 and ((int(Lo[s[2], s[1], s[0]]) >> 24) & 7) == 0)]
    assert not bad, bad[:5]
    # sampler on the fork's law (CountLiquid frozen per call), occupied count from the census
    p = ctx.run_parallel(n_sweeps=10, events_per_tile=2.0, seed=5)
    assert p.events_done > 0
    c = ctx.counters()
    assert c.num_adatom_occ > 0
    ctx.close(); orc.close()
