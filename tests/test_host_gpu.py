"""The reference's own regression procedure (regression-tests/regression-tests.ipynb) on the B200 host program:
copy a start.dat into an empty directory, run `mulskips.e | tee log.txt` at the suite's 480^3 box, compare
I00000000.xyz, I00000001.xyz and the numeric log tokens.

mulskips_b200/bin/mulskips.e is the C++ host over the C ABI with the event loop on the GPU in serial-parity
mode.  Default: MaxIter cut to 2e5 events and the CPU oracle's command line run on the same start.dat as the
checker (frames byte-identical, log tokens equal, Time bit-equal) -- the oracle itself is pinned on the full
goldens by tests/test_oracle_golden.py.  MSKC_FULL_GOLDEN=1: the unmodified start files (1e7 events, ~4 min
of GPU each at the serial kernel's 4.6e4 events/s) against the reference goldens' sha256 in
tests/golden/manifest.json; the last full run is recorded in profiles/r01_host_full_goldens.txt."""
import hashlib
import json
import os
import subprocess

import pytest

from cases import GOLDEN, make_cad_case, make_la_case, make_la_tree_case, make_sige_case, make_st_case, make_start, start_path

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "mulskips_b200", "bin", "mulskips.e")
ORACLE = os.path.join(ROOT, "oracle", "mulskips_oracle.e")
MAN = json.load(open(os.path.join(GOLDEN, "manifest.json")))
CASES = ["Si-SL", "surface", "cube", "cube-zigzag", "sphere", "inverted-pyramid", "finfet", "CVD-Si-SiH2"]      # InitSt FS, F, C, C (PtransZig 0.8), S, I, FF, FS + coverage
FULL = os.environ.get("MSKC_FULL_GOLDEN") == "1"
NEV = 200000


def _tokens(log, key):
    for line in open(log):
        if line.strip().startswith(key):
            return line.split()
    return None


def _run(exe, d, args=()):
    # (the serial loop is asked for explicitly wherever files are compared byte for byte with the oracle's: without an
    # option the host program picks the loop by RunType, see test_default_mode_follows_runtype)
    if exe == EXE and "--parallel" not in args and "--auto" not in args:
        args = ["--parity", *args]
    args = [a for a in args if a != "--auto"]
    with open(d / "log.txt", "w") as out, open(d / "err.txt", "w") as err:
        rc = subprocess.run([exe, *args], cwd=d, stdout=out, stderr=err).returncode
    assert rc == 0, open(d / "log.txt").read()[-1500:] + open(d / "err.txt").read()[-1500:]


@pytest.fixture(scope="module")
def built():
    if not (os.path.exists(EXE) and os.path.exists(ORACLE)):
        import __graft_entry__ as g
        g.build()


@pytest.mark.parametrize("case", CASES)
def test_regression_case(built, tmp_path, case):
    g, o = tmp_path / "gpu", tmp_path / "oracle"
    g.mkdir(); o.mkdir()
    if FULL:
        with open(start_path(case)) as f:
            (g / "start.dat").write_text(f.read())
    else:       # CVD-Si-SiH2 exits on KMC time: 0.5 s of it is ~2e5 events
        make_start(case, str(g / "start.dat"), maxiter=NEV, outmolmol=NEV, maxtime=0.5, outtime=0.5)
        (o / "start.dat").write_text((g / "start.dat").read_text())
        _run(ORACLE, o)
    _run(EXE, g)
    if FULL:
        m = MAN[case]
        for k in (0, 1):
            b = open(g / f"I0000000{k}.xyz", "rb").read()
            assert hashlib.sha256(b).hexdigest() == m[f"frame{k}_sha256"], f"{case} frame {k} differs from the reference golden"
        t = _tokens(g / "log.txt", "Iter, Time, Site")
        assert int(t[3]) == m["final_iter"] and [int(v) for v in t[5:8]] == m["final_site"] and float(t[4]) == float(m["final_time"])
        assert [int(v) for v in _tokens(g / "log.txt", "NatNvoidNad")[1:4]] == m["final_lists"]
        return
    for name in ("I00000000.xyz", "I00000001.xyz", "I00000000_d.xyz", "I00000001_d.xyz", "I00000000_v.xyz", "I00000001_v.xyz"):
        assert (g / name).read_bytes() == (o / name).read_bytes(), f"{case}: {name} differs from the oracle's"
    if (o / "I00000001_w.xyz").exists():
        assert (g / "I00000001_w.xyz").read_bytes() == (o / "I00000001_w.xyz").read_bytes()
    assert hashlib.sha256((g / "I00000000.xyz").read_bytes()).hexdigest() == MAN[case]["frame0_sha256"]      # frame 0 is the reference golden itself
    for key in ("Atoms,Voids,AdAtoms", "Iter, Time, Site", "NatNvoidNad", "NumAtWrong:"):
        assert _tokens(g / "log.txt", key) == _tokens(o / "log.txt", key), key
    assert (g / "movie.mos").exists() and (g / "fort.30").exists() and (g / "I00000001.src").exists()


@pytest.mark.parametrize("kind", ["CAD", "LA", "SiGe", "ST", "ST-fixedT", "SiGeBdoped", "LA-fork"])
def test_file_driven_case(built, tmp_path, kind):
    """InitSt SG (SetSiGex.f), InitSt IN (SetCAD.f: geometry file with crystal / wall / empty voxels, H coverage) and InitSt LA (SetT.f + SetCAD.f:
    temperature file, LA rate law, liquid nucleus) through the host program, on synthetic files in the reference's formats at
    a small box, every output file against the oracle's.  (The reference's real 480x240x240 inputs run at full length
    against the reference goldens in test_reference_inputs_full_length below.)"""
    g, o = tmp_path / "gpu", tmp_path / "oracle"
    g.mkdir(); o.mkdir()
    box = (96, 96, 120) if kind == "CAD" else (96, 96, 96)
    if kind == "LA-fork":                            # mulskips-source-LaserAnnealing semantics (--la-tree): SiGe melt nucleus, liquid-fraction coupling
        start = make_la_tree_case(str(tmp_path), box, maxiter=60000)
    elif kind.startswith("ST"):                      # InitSt ST (SetTrenchSi.f + SetProbabilityFromField.f): trench, temperature ramp, tabulated rate blocks
        box = (96, 48, 144)
        start = make_st_case(str(tmp_path), maxiter=60000, **(dict(ttop=1000.0, tbottom=1000.0) if kind.endswith("fixedT") else {}))
    elif kind == "SiGeBdoped":                       # examples/ex5-CVD-IO/CVD-SiGeBdoped/start.dat: 3 crystal + 2 coverage species, 288 evaporation classes
        box = (96, 96, 240)
        start = make_start("SiGeBdoped", str(tmp_path / "start.dat"), lens=(48, 0, 0), maxiter=60000, outmolmol=60000)
    elif kind == "SiGe":                               # InitSt SG (SetSiGex.f): random-alloy substrate, the C5 slab's second material
        start = make_sige_case(str(tmp_path), 48, 0.3, maxiter=60000)
    else:
        start = make_cad_case(str(tmp_path), box, maxiter=60000) if kind == "CAD" else make_la_case(str(tmp_path), box, maxiter=3000)   # (the nucleus of this box regrows in ~4000 events)
    for d in (g, o):
        (d / "start.dat").write_text(open(start).read())
    args = ["--box", *[str(v) for v in box]] + (["--la-tree"] if kind == "LA-fork" else [])
    _run(ORACLE, o, args)
    _run(EXE, g, args)
    names = ["I00000000.xyz", "I00000001.xyz", "I00000000_d.xyz", "I00000001_d.xyz", "I00000000_v.xyz", "I00000001_v.xyz"]
    for name in names:
        assert (g / name).read_bytes() == (o / name).read_bytes(), f"{kind}: {name} differs from the oracle's"
    for key in ("Atoms,Voids,AdAtoms", "Iter, Time, Site", "NatNvoidNad", "Max temperature is:", "Number of occupied sites"):
        assert _tokens(g / "log.txt", key) == _tokens(o / "log.txt", key), key
    assert len(open(g / "I00000000.xyz").readlines()) > 50
    if kind in ("LA", "LA-fork"):                                 # the Coor (phase) file the FEM side of the LA coupling reads, main.f:407-430
        assert (g / "LattStat.dat").read_bytes() == (o / "LattStat.dat").read_bytes()


@pytest.mark.parametrize("case", ["CAD", "LA"])
def test_reference_inputs_full_length(built, tmp_path, case):
    """BASELINE configs C3 (CVD on a voxelised nanostructure, InitSt IN) and C4 (laser annealing, InitSt LA) with the
    reference's REAL inputs -- LattGeo / TempStat 480x240x240 unpacked from the reference's own tarballs
    (tests/golden/CAD-input.tar.gz, LA-input.tar.gz) -- and its unmodified start files, at full length (3e6 / 1e6 events),
    against the reference goldens: sha256 of I00000000.xyz / I00000001.xyz and the numeric log tokens of log-CAD.txt:17,24-25 /
    log-LA.txt:26,33-34.  MSKC_CAD_NOSKIP=1 selects the plane offset of the build that produced these goldens."""
    import tarfile
    inputs = tmp_path / "inputs"
    inputs.mkdir()
    with tarfile.open(os.path.join(GOLDEN, f"{case}-input.tar.gz")) as tf:
        tf.extractall(inputs)
    g = tmp_path / "gpu"
    g.mkdir()
    paths = [str(inputs / "LattGeo.480x240x240.withzeros.dat")] if case == "CAD" else [str(inputs / "LattGeo.480x240x240.dat"), str(inputs / "TempStat.480x240x240.dat")]
    make_start(case, str(g / "start.dat"), paths=paths)
    env = dict(os.environ, MSKC_CAD_NOSKIP="1")
    with open(g / "log.txt", "w") as out, open(g / "err.txt", "w") as err:
        rc = subprocess.run([EXE, "--box", "480", "240", "240"], cwd=g, stdout=out, stderr=err, env=env).returncode
    assert rc == 0, open(g / "log.txt").read()[-1500:] + open(g / "err.txt").read()[-1500:]
    m = MAN[case]
    for k in (0, 1):
        b = open(g / f"I0000000{k}.xyz", "rb").read()
        assert b.count(b"\n") == m[f"frame{k}_lines"]
        assert hashlib.sha256(b).hexdigest() == m[f"frame{k}_sha256"], f"{case} frame {k} differs from the reference golden"
    assert [int(v) for v in _tokens(g / "log.txt", "Atoms,Voids,AdAtoms")[1:4]] == m["init_lists"]
    t = _tokens(g / "log.txt", "Iter, Time, Site")
    assert int(t[3]) == m["final_iter"] and [int(v) for v in t[5:8]] == m["final_site"] and float(t[4]) == float(m["final_time"])
    assert [int(v) for v in _tokens(g / "log.txt", "NatNvoidNad")[1:4]] == m["final_lists"]


@pytest.mark.parametrize("runtype,mode", [("T", "serial-parity"), ("R", "sublattice-sampler")])
def test_default_mode_follows_runtype(built, tmp_path, runtype, mode):
    """pymulskips launches `mulskips.e` with no arguments (setuprun.py:414).  RunType T (regression: the reference's own
    event sequence) -> serial parity loop; RunType R (production: "truly random") -> the sublattice sampler."""
    make_start("surface", str(tmp_path / "start.dat"), lens=(48, 0, 0), maxiter=30000, outmolmol=30000, runtype=runtype)
    _run(EXE, tmp_path, ["--auto", "--box", "96", "96", "240"])
    assert f"({mode} mode)" in (tmp_path / "err.txt").read_text()
    assert (tmp_path / "I00000001.xyz").exists()


def test_checkpoint_restart(built, tmp_path):
    """SaveFinalState 'T' writes the lattice words (main.f:385-403); RunType 'C' reads them back through the checkpoint
    branch of SetCAD.f:54-79, rebuilds the lists with ideal diamond bonds (the reference's restart is lossy in exactly
    this way: with adsorbed coverage or hexagonal bonds its own Get_Prob_Ini census stops the run, so the case here is
    the laser-annealing one, pure cubic Si) and continues with ran3.  Both legs on the host program, checked against the
    oracle's."""
    box = (96, 96, 96)
    args = ["--box", *[str(v) for v in box]]
    start = make_la_case(str(tmp_path), box, maxiter=1500)
    lines = open(start).read().split("\n")
    i = next(k for k, l in enumerate(lines) if "Save final state" in l)
    lines[i] = "T ! Save final state"
    lines.insert(i + 1, "LattChk.dat")
    first = "\n".join(lines)
    j = next(k for k, l in enumerate(lines) if "RunType" in l)
    lines[j] = "C ! RunType: continuation"
    lines[3] = "LattChk.dat"                       # the geometry path line now names the checkpoint
    second = "\n".join(lines)
    out = {}
    for tag, exe in (("gpu", EXE), ("oracle", ORACLE)):
        a, b = tmp_path / f"{tag}1", tmp_path / f"{tag}2"
        a.mkdir(); b.mkdir()
        (a / "start.dat").write_text(first)
        _run(exe, a, args)
        (b / "start.dat").write_text(second)
        (b / "LattChk.dat").write_bytes((a / "LattChk.dat").read_bytes())
        _run(exe, b, args)
        out[tag] = (a, b)
    assert (out["gpu"][0] / "LattChk.dat").read_bytes() == (out["oracle"][0] / "LattChk.dat").read_bytes()
    for name in ("I00000000.xyz", "I00000001.xyz", "I00000001_v.xyz", "LattChk.dat", "LattStat.dat"):
        assert (out["gpu"][1] / name).read_bytes() == (out["oracle"][1] / name).read_bytes(), name
    for key in ("Atoms,Voids,AdAtoms", "Iter, Time, Site", "NatNvoidNad", "Max temperature is:"):
        assert _tokens(out["gpu"][1] / "log.txt", key) == _tokens(out["oracle"][1] / "log.txt", key), key


def test_socket_driven_laser_annealing(built, tmp_path):
    """The LaserAnnealing fork end to end: the host program as socket client of a server that speaks the
    protocol of pymulskips/sockets.py.  Leg 1: geometry + temperature arrive through the socket, 30 000 events, phases
    (coordination of every occupied cell, 10 = wall) and species go back.  Leg 2: a new, colder temperature map; the
    lattice keeps its listed rates until events touch them (la/main.f:616-631), clock and frame counters restart.
    Checked against the oracle driven through the same two legs: phases, species map and the frames of both legs."""
    import numpy as np
    from la_socket_server import Server
    from oracle_binding import Oracle
    box = (96, 96, 96)
    g, o = tmp_path / "gpu", tmp_path / "oracle"
    g.mkdir(); o.mkdir()
    twin = make_la_tree_case(str(o), box=box, maxiter=30000)
    make_la_tree_case(str(g), box=box, maxiter=30000, sockets=True)
    geo = np.loadtxt(o / "LattGeo.dat", dtype=np.int32)
    t1 = np.loadtxt(o / "TempStat.dat", dtype=np.int32)
    t2 = t1 - 25
    orc = Oracle(*box, twin, la_tree=True, outdir=str(o), write_files=True)
    orc.setup()

    def phases(L):                                   # la/main.f:795-817
        L = L[:-1].astype(np.int64)
        wall = (L & 255) == 113
        occ = ((L >> 23) & 1) == 1
        c = np.where(wall, 10, np.where(occ, (L >> 24) & 7, 0))
        d = np.where(wall, 10, np.where(occ, (L >> 28) & 3, 0))
        return c.reshape(box[2], -1).astype(np.int32), d.reshape(box[2], -1).astype(np.int32)

    srv = Server((box[2], box[0] * box[1]), timeout=240.0)
    out = open(g / "log.txt", "w")
    p = subprocess.Popen([EXE, "--la-tree", "-p", str(srv.port), "--box", *[str(v) for v in box]], cwd=g, stdout=out, stderr=subprocess.STDOUT)
    try:
        srv.accept()
        assert srv.status() == "NEEDINIT"
        srv.init()
        srv.send("geo", geo)
        srv.send("field", t1)
        srv.wait_status("HAVEDATA")
        c1, d1 = srv.get("phases"), srv.get("xge")
        orc.run(10**9); orc.finish()
        ec1, ed1 = phases(orc.lattcoo())
        assert np.array_equal(c1, ec1) and np.array_equal(d1, ed1)
        frame1 = (g / "I00000001.xyz").read_bytes()
        assert frame1 == (o / "I00000001.xyz").read_bytes()
        assert srv.status() == "READY"
        srv.send("field", t2)
        srv.wait_status("HAVEDATA")
        c2, d2 = srv.get("phases"), srv.get("xge")
        orc.next_leg(t2)
        orc.run(10**9); orc.finish()
        ec2, ed2 = phases(orc.lattcoo())
        assert np.array_equal(c2, ec2) and np.array_equal(d2, ed2)
        assert not np.array_equal(c1, c2)
        assert (g / "I00000001.xyz").read_bytes() == (o / "I00000001.xyz").read_bytes() != frame1
        srv.exit()
        assert p.wait(timeout=60) == 0
    finally:
        srv.close()
        if p.poll() is None:
            p.kill()
        out.close()
    orc.close()


MORE_BUILDERS = [("SS", "Si-SL", (60, 0, 0), (144, 144, 144), 30000), ("D", "inverted-pyramid", (30, 93, 0), (96, 96, 144), 1000),
                 ("Z", "Si-SL", (30, 93, 0), (96, 96, 144), 5000), ("T", "cube", (20, 45, 90), (96, 48, 144), 30000),
                 ("Fw", "surface", (48, 96, 0), (96, 96, 144), 30000), ("Fa", "surface", (48, 72, 48), (96, 96, 144), 30000),
                 ("K", "inverted-pyramid", (30, 93, 0), (96, 96, 144), 2000), ("L", "inverted-pyramid", (30, 93, 0), (96, 96, 144), 2000),
                 ("J", "inverted-pyramid", (30, 93, 0), (96, 96, 144), 2000)]


@pytest.mark.parametrize("initst,table,lens,box,nrun", MORE_BUILDERS, ids=[b[0] for b in MORE_BUILDERS])
def test_more_builders(built, tmp_path, initst, table, lens, box, nrun):
    """SetNCSpSi.f (SS), SetInvPC.f (D), SetInvPSi.f (Z), SetTrench.f (T) through the host program with the loop on the
    GPU: every output file and the log tokens equal the oracle's (SS ends on 'few MC particles', the others on MaxIter)."""
    g, o = tmp_path / "gpu", tmp_path / "oracle"
    g.mkdir(); o.mkdir()
    make_start(table, str(g / "start.dat"), lens=lens, maxiter=nrun, outmolmol=nrun, initst=initst)
    (o / "start.dat").write_text((g / "start.dat").read_text())
    args = ["--box", *[str(v) for v in box]]
    _run(ORACLE, o, args)
    _run(EXE, g, args)
    names = sorted(p.name for p in o.glob("I*.xyz"))
    assert "I00000000.xyz" in names and len(names) >= 3
    for name in names:
        assert (g / name).read_bytes() == (o / name).read_bytes(), f"{initst}: {name} differs from the oracle's"
    for key in ("Atoms,Voids,AdAtoms", "Iter, Time, Site", "NatNvoidNad", "few MC particles"):
        assert _tokens(g / "log.txt", key) == _tokens(o / "log.txt", key), key


@pytest.mark.parametrize("kind", ["surface", "LA-fork"])
def test_host_parallel_mode(built, tmp_path, kind):
    """`mulskips.e --parallel`: the production mode of the drop-in (synchronous-sublattice sampler) writes the same set of
    files at the same cadence as the serial loop; with --la-tree the sampler runs on the fork's law (CountLiquid per call,
    occupied count from the census).  Statistical equivalence is tests/test_sampler_stats_gpu.py's subject; here: the run
    completes, frames are well-formed and the crystal evolved."""
    box = (96, 96, 240) if kind == "surface" else (96, 96, 96)
    args = ["--parallel", "--box", *[str(v) for v in box]]
    if kind == "surface":
        make_start("surface", str(tmp_path / "start.dat"), lens=(48, 0, 0), maxiter=40000, outmolmol=20000, runtype="R")
    else:
        make_la_tree_case(str(tmp_path), box, maxiter=40000)
        lines = (tmp_path / "start.dat").read_text().split("\n")
        i = next(k for k, l in enumerate(lines) if "OutMolMol" in l)
        lines[i] = "20000 ! OutMolMol"
        (tmp_path / "start.dat").write_text("\n".join(lines))
        args.append("--la-tree")
    _run(EXE, tmp_path, args)
    frames = sorted(p.name for p in tmp_path.glob("I*[0-9].xyz"))
    assert frames[0] == "I00000000.xyz" and len(frames) >= 2
    n0 = int(open(tmp_path / frames[0]).readline().split()[0])
    n1 = int(open(tmp_path / frames[-1]).readline().split()[0])
    assert n0 > 50 and n1 > 50
    body0 = open(tmp_path / frames[0]).read().split("\n", 2)[2]
    body1 = open(tmp_path / frames[-1]).read().split("\n", 2)[2]
    assert body0 != body1
    t = _tokens(tmp_path / "log.txt", "Iter, Time, Site")
    assert t is not None and int(t[3]) > 20000 and float(t[4]) > 0
    err = open(tmp_path / "err.txt").read()
    assert "sublattice-sampler" in err
    # every frame is well-formed: the header count is the number of atom lines, and no vacancy sits at the absolute origin
    # (the sampler's vacancy list comes from the lattice, not from the serial loop's list)
    for f in frames:
        lines = [l for l in open(tmp_path / f).read().split("\n")[2:] if l.strip()]
        assert int(open(tmp_path / f).readline().split()[0]) == len(lines), f
        assert not any(l.split()[0] == "O" and all(float(v) == 0.0 for v in l.split()[1:4]) for l in lines), f


def test_shape_builder(built, tmp_path):
    """InitSt SH (initial/SetShape.f): a structure given as a list of filled sites, neighbours found on the whole first
    shell, through the host program with the loop on the GPU; every output file and the log tokens equal the oracle's."""
    from cases import make_shape_case
    box = (144, 144, 144)
    g, o = tmp_path / "gpu", tmp_path / "oracle"
    g.mkdir(); o.mkdir()
    start = make_shape_case(str(tmp_path), box=box, maxiter=60000)
    for d in (g, o):
        (d / "start.dat").write_text(open(start).read())
    args = ["--box", *[str(v) for v in box]]
    _run(ORACLE, o, args)
    _run(EXE, g, args)
    names = sorted(p.name for p in o.glob("I*.xyz"))
    assert len(names) >= 6
    for name in names:
        assert (g / name).read_bytes() == (o / name).read_bytes(), f"SH: {name} differs from the oracle's"
    for key in ("Atoms,Voids,AdAtoms", "Iter, Time, Site", "NatNvoidNad"):
        assert _tokens(g / "log.txt", key) == _tokens(o / "log.txt", key), key


@pytest.mark.parametrize("case", ["surface-zig0.8", "CVD-Si-SiH2"])
def test_exact_checkpoint_restart(built, tmp_path, case):
    """SURVEY 8 f2, second half: the reference's checkpoint keeps LattCoo only and its restart re-derives the lists (lossy:
    list order, hexagonal bonds, rates, generator, clock).  Beside it the host program writes `<name>.mskc` -- the complete
    state -- and `--exact-restart` continues the SAME trajectory: a run of N events + checkpoint + restart to 2N must leave
    the files and the log lines of an uninterrupted run of 2N events, bit for bit (hexagonal-phase SiC with the
    numberslist stream; Si CVD with coverage, ran3, exit on KMC time)."""
    box = ["--box", "96", "96", "240"]
    if case == "surface-zig0.8":
        kw1, kw2 = dict(maxiter=30000, outmolmol=15000), dict(maxiter=60000, outmolmol=15000)
        mk = lambda dst, kw: make_start("surface", dst, lens=(48, 0, 0), ptranszig=0.8, **kw)
        late = ["I00000003", "I00000004"]
    else:
        kw1, kw2 = dict(maxtime=0.02, outtime=0.01), dict(maxtime=0.04, outtime=0.01)
        mk = lambda dst, kw: make_start("CVD-Si-SiH2", dst, lens=(120, 0, 0), runtype="R", idum=-4711, **kw)
        late = ["I00000003", "I00000004"]
    a, b1, b2 = tmp_path / "a", tmp_path / "b1", tmp_path / "b2"
    for d in (a, b1, b2):
        d.mkdir()
    mk(str(a / "start.dat"), kw2)
    _run(EXE, a, box)
    mk(str(b1 / "start.dat"), kw1)
    lines = (b1 / "start.dat").read_text().split("\n")
    i = next(k for k, l in enumerate(lines) if "Save final state" in l or "SaveFinalState" in l)
    lines[i] = "T ! Save final state"
    lines.insert(i + 1, "LattChk.dat")
    (b1 / "start.dat").write_text("\n".join(lines))
    _run(EXE, b1, box)
    assert (b1 / "LattChk.dat.mskc").exists()
    mk(str(b2 / "start.dat"), kw2)
    _run(EXE, b2, [*box, "--exact-restart", str(b1 / "LattChk.dat.mskc")])
    la = [l for l in open(a / "log.txt") if l.startswith((" Iter, Time, Site", " NatNvoidNad"))]
    lb = [l for l in open(b2 / "log.txt") if l.startswith((" Iter, Time, Site", " NatNvoidNad"))]
    assert len(lb) >= 2 and la[-len(lb):] == lb, "the continued run's log lines differ from the uninterrupted run's"
    names = sorted(p.name for p in b2.glob("I0*.xyz"))
    assert any(n.startswith(late[0]) for n in names) and any(n.startswith(late[1]) for n in names), names
    for n in names:
        assert (b2 / n).read_bytes() == (a / n).read_bytes(), f"{n} differs from the uninterrupted run's"
