"""Host logic of the drop-in executable without a GPU: mulskips_b200/bin/mulskips.e parses start.dat, builds the
initial geometry with the reference's builders, writes frame 0 and the start-up log lines, and only then asks the
CUDA library for a device (on a box without one it stops there, loudly: no CPU fallback).  Frame 0 and the list
sizes must equal the oracle's for the start files no reference golden covers (InitSt ST, SG with 3+2 species)."""
import os
import subprocess

import pytest

from cases import make_la_tree_case, make_st_case, make_start

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "mulskips_b200", "bin", "mulskips.e")
ORACLE = os.path.join(ROOT, "oracle", "mulskips_oracle.e")


def _tokens(text, key):
    for line in text.splitlines():
        if line.strip().startswith(key):
            return line.split()
    return None


@pytest.mark.parametrize("kind", ["ST-ramp", "ST-fixedT", "SiGeBdoped", "LA-fork-SiGe", "LA-fork-Si"])
def test_frame0_and_startup_log(tmp_path, kind):
    if not (os.path.exists(EXE) and os.path.exists(ORACLE)):
        import __graft_entry__ as g
        g.build()
    g, o = tmp_path / "host", tmp_path / "oracle"
    g.mkdir(); o.mkdir()
    if kind.startswith("ST"):
        box = (96, 48, 144)
        kw = dict(ttop=1000.0, tbottom=1000.0) if kind.endswith("fixedT") else {}
        start = make_st_case(str(tmp_path), maxiter=1, **kw)
    elif kind.startswith("LA-fork"):                 # the LaserAnnealing fork's grammar, layout and start-up bookkeeping (--la-tree)
        box = (96, 96, 96)
        start = make_la_tree_case(str(tmp_path), box=box, maxiter=1, sige=kind.endswith("SiGe"), tbase=1480 if kind.endswith("SiGe") else 1655)
    else:
        box = (96, 96, 240)
        start = make_start("SiGeBdoped", str(tmp_path / "start.dat"), lens=(48, 0, 0), maxiter=1, outmolmol=1)
    args = ["--box", *[str(v) for v in box]] + (["--la-tree"] if kind.startswith("LA-fork") else [])
    logs = {}
    for d, exe in ((g, EXE), (o, ORACLE)):
        (d / "start.dat").write_text(open(start).read())
        r = subprocess.run([exe, *args], cwd=d, capture_output=True, text=True)
        logs[d] = r.stdout
        if exe == ORACLE:
            assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    for name in ("I00000000.xyz", "I00000000_d.xyz", "I00000000_v.xyz"):
        assert (g / name).read_bytes() == (o / name).read_bytes(), f"{kind}: {name}"
    assert len((g / "I00000000.xyz").read_text().splitlines()) > 50
    for key in ("Atoms,Voids,AdAtoms", "maxElines", "Temperature on top", "Reading input probability table", "Initial number of occupied sites (before",
                "Initial CountCrystal (before", "Initial CountLiquid", "Initial number of liquid sites", "CountLiquid", "xLiquid"):
        assert _tokens(logs[g], key) == _tokens(logs[o], key), key


# further builders of main.f:225-295 on the regression tables: (InitSt, table of, Len line, box)
# (these small pitted slabs etch away under the regression tables -- the builders put no wall under z = 0 -- so the
# follow-up run is kept short of the moment the etch front reaches the bottom plane)
MORE_BUILDERS = [("SS", "Si-SL", (60, 0, 0), (144, 144, 144), 30000), ("D", "inverted-pyramid", (30, 93, 0), (96, 96, 144), 1000),
                 ("Z", "Si-SL", (30, 93, 0), (96, 96, 144), 5000), ("T", "cube", (20, 45, 90), (96, 48, 144), 30000),
                 ("Fw", "surface", (48, 96, 0), (96, 96, 144), 30000), ("Fa", "surface", (48, 72, 48), (96, 96, 144), 30000),
                 ("K", "inverted-pyramid", (30, 93, 0), (96, 96, 144), 2000), ("L", "inverted-pyramid", (30, 93, 0), (96, 96, 144), 2000),
                 ("J", "inverted-pyramid", (30, 93, 0), (96, 96, 144), 2000), ("E", "cube", (60, 24, 24), (96, 96, 144), 30000),
                 ("A", "surface", (72, 0, 0), (144, 144, 120), 30000)]


@pytest.mark.parametrize("initst,table,lens,box,nrun", MORE_BUILDERS, ids=[b[0] for b in MORE_BUILDERS])
def test_more_builders_frame0(tmp_path, initst, table, lens, box, nrun):
    """SetNCSpSi.f (SS), SetInvPC.f (D), SetInvPSi.f (Z), SetTrench.f (T), the wall / antiphase variants and SetSiCSF.f (E):
    the host program's initial structure, list
    sizes and frame 0 equal the oracle's; the oracle then runs nrun events on it without tripping a consistency check."""
    if not (os.path.exists(EXE) and os.path.exists(ORACLE)):
        import __graft_entry__ as g
        g.build()
    g, o = tmp_path / "host", tmp_path / "oracle"
    g.mkdir(); o.mkdir()
    start = make_start(table, str(tmp_path / "start.dat"), lens=lens, maxiter=1, outmolmol=1, initst=initst)
    args = ["--box", *[str(v) for v in box]]
    logs = {}
    for d, exe in ((g, EXE), (o, ORACLE)):
        (d / "start.dat").write_text(open(start).read())
        r = subprocess.run([exe, *args], cwd=d, capture_output=True, text=True)
        logs[d] = r.stdout
        if exe == ORACLE:
            assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    long = tmp_path / "long"
    long.mkdir()
    make_start(table, str(long / "start.dat"), lens=lens, maxiter=nrun, outmolmol=nrun, initst=initst)
    r = subprocess.run([ORACLE, *args, "--no-files"], cwd=long, capture_output=True, text=True)
    assert r.returncode == 0 and "Error" not in r.stdout, r.stdout[-1000:] + r.stderr[-500:]
    for name in ("I00000000.xyz", "I00000000_d.xyz", "I00000000_v.xyz"):
        assert (g / name).read_bytes() == (o / name).read_bytes(), f"{initst}: {name}"
    assert len((g / "I00000000.xyz").read_text().splitlines()) > 50
    assert _tokens(logs[g], "Atoms,Voids,AdAtoms") == _tokens(logs[o], "Atoms,Voids,AdAtoms")


def test_socket_handshake_and_first_frame(tmp_path):
    """The laser-annealing socket loop (mulskips-source-LaserAnnealing/main.f:160-345): the driver connects, answers
    STATUS with NEEDINIT / READY, takes the geometry and temperature arrays, and builds the same initial structure as
    the file-driven run (frame 0 equal to the oracle's).  Without a GPU it stops when it asks the CUDA library for a
    device; with one it runs the single event allowed here and reports HAVEDATA."""
    import numpy as np
    from la_socket_server import Server
    box = (96, 96, 96)
    g, o = tmp_path / "host", tmp_path / "oracle"
    g.mkdir(); o.mkdir()
    twin = make_la_tree_case(str(o), box=box, maxiter=1)                      # file-driven twin for the oracle
    start = make_la_tree_case(str(g), box=box, maxiter=1, sockets=True)
    geo = np.loadtxt(o / "LattGeo.dat", dtype=np.int32)                        # [LenZ][LenX*LenY]
    tmp = np.loadtxt(o / "TempStat.dat", dtype=np.int32)
    r = subprocess.run([ORACLE, "--la-tree", "--box", *[str(v) for v in box]], cwd=o, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-800:]
    srv = Server((box[2], box[0] * box[1]), timeout=60.0)
    out = open(g / "log.txt", "w")
    p = subprocess.Popen([EXE, "--la-tree", "-p", str(srv.port), "--box", *[str(v) for v in box]], cwd=g, stdout=out, stderr=subprocess.STDOUT)
    try:
        srv.accept()
        assert srv.status() == "NEEDINIT"
        srv.init()
        assert srv.status() == "READY"
        srv.send("geo", geo)
        assert srv.status() == "READY"
        srv.send("field", tmp)
        try:                                   # GPU present: one event, then the phases are on offer
            if srv.wait_status("HAVEDATA") == "HAVEDATA":
                c = srv.get("phases")
                assert c.shape == (box[2], box[0] * box[1]) and c.max() == 10
            srv.exit()
        except (ConnectionError, OSError):
            pass                               # no GPU: the driver stopped at mskc_create, after frame 0
        p.wait(timeout=120)
    finally:
        srv.close()
        if p.poll() is None:
            p.kill()
        out.close()
    log = (g / "log.txt").read_text()
    assert "@ message to server: NEEDINIT" in log and "Setting input geometry from sockets" in log
    for name in ("I00000000.xyz", "I00000000_v.xyz"):
        assert (g / name).read_bytes() == (o / name).read_bytes(), name
    for key in ("Atoms,Voids,AdAtoms", "Initial number of occupied sites (before", "Initial CountLiquid"):
        assert _tokens(log, key) == _tokens(r.stdout, key), key
