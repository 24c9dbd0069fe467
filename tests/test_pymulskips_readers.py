"""The readers of pymulskips on the drop-in's output (process contract): pymulskips.analyze is imported
from the reference tree (numpy only) and pointed at the run directory the host program wrote.  Runs where the reference
tree is mounted (this container); the GPU box has no /root/reference and skips.  Without a GPU the host program stops
after frame 0, which is what is read here; the unmodified regression start file at the suite's 480^3 box is used so
that frame 0 is also the reference's own golden file, byte for byte."""
import hashlib
import importlib.util
import json
import os
import subprocess
import sys

import pytest

from cases import GOLDEN, start_path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "mulskips_b200", "bin", "mulskips.e")
REF = "/root/reference"
RTIN = os.path.join(REF, "regression-tests", "input-and-references-files")


def _analyze():
    p = os.path.join(REF, "pymulskips", "analyze.py")
    if not os.path.exists(p):
        pytest.skip("reference tree not mounted")
    spec = importlib.util.spec_from_file_location("pymulskips_analyze", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("case", ["Si-SL", "cube"])
def test_analyze_reads_frame0(tmp_path, case):
    an = _analyze()
    if not os.path.exists(EXE):
        import __graft_entry__ as g
        g.build()
    (tmp_path / "start.dat").write_text(open(start_path(case)).read())
    subprocess.run([EXE], cwd=tmp_path, capture_output=True, text=True)        # 480^3 default box; stops at mskc_create without a GPU
    f0 = tmp_path / "I00000000.xyz"
    assert f0.exists()
    man = json.load(open(os.path.join(GOLDEN, "manifest.json")))
    assert hashlib.sha256(f0.read_bytes()).hexdigest() == man[case]["frame0_sha256"]
    golden = os.path.join(RTIN, f"I00000000-{case}.xyz")
    files = an.read_output_files(str(tmp_path), "undercoordinated")
    assert str(f0) in files
    assert an.get_box_sides(str(f0)) == an.get_box_sides(golden)
    h, hg = an.get_surface_height(str(f0)), an.get_surface_height(golden)
    assert h == hg and h > 0
    assert len(an.read_output_files(str(tmp_path), "vacancies")) >= 1 and len(an.read_output_files(str(tmp_path), "defects")) >= 1


def test_reference_socket_server_drives_the_host(tmp_path):
    """The reference's OWN server classes (pymulskips/sockets.py: InterfaceSocket, Driver, Status) against the host
    program as socket client: STATUS -> NEEDINIT, INIT, STATUS -> READY, SENDGEO, SENDFIELD.  The client builds its
    structure from the two arrays (frame 0 appears); without a GPU it then stops at mskc_create, with one it runs the
    single event allowed here and offers the phases."""
    import numpy as np
    from cases import make_la_tree_case
    p = os.path.join(REF, "pymulskips", "sockets.py")
    if not os.path.exists(p):
        pytest.skip("reference tree not mounted")
    spec = importlib.util.spec_from_file_location("pymulskips_sockets", p)
    S = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(S)
    if not os.path.exists(EXE):
        import __graft_entry__ as g
        g.build()
    box = (96, 96, 96)
    files = tmp_path / "files"; run = tmp_path / "run"
    files.mkdir(); run.mkdir()
    make_la_tree_case(str(files), box=box, maxiter=1)                          # only for its LattGeo / TempStat arrays
    make_la_tree_case(str(run), box=box, maxiter=1, sockets=True)
    geo = np.loadtxt(files / "LattGeo.dat", dtype=np.int32)
    tmp = np.loadtxt(files / "TempStat.dat", dtype=np.int32)
    import socket as pysocket
    probe = pysocket.socket(); probe.bind(("127.0.0.1", 0)); port = probe.getsockname()[1]; probe.close()
    iface = S.InterfaceSocket(address="127.0.0.1", port=port)
    iface.open()
    iface.server.settimeout(60)
    out = open(run / "log.txt", "w")
    proc = subprocess.Popen([EXE, "--la-tree", "-h", "127.0.0.1", "-p", str(port), "--box", *[str(v) for v in box]], cwd=run, stdout=out, stderr=subprocess.STDOUT)
    try:
        client, _ = iface.server.accept()
        drv = S.Driver(client)           # (no timeout on `client`: Driver re-wraps the same descriptor as a blocking socket)
        assert drv.get_status() == (S.Status.Up | S.Status.NeedsInit)
        drv.initialise()
        assert drv.get_status() == (S.Status.Up | S.Status.Ready)
        drv.send_data(geo, "geo")
        assert drv.get_status() == (S.Status.Up | S.Status.Ready)
        drv.send_data(tmp, "field")
        try:
            st = drv.get_status()
            if st & S.Status.HasData:                                           # GPU present
                c = drv.get_data(np.zeros(geo.shape, np.int32), "phases")
                assert c.shape == geo.shape and c.max() == 10
                drv.shutdown()
        except Exception:
            pass
        proc.wait(timeout=120)
    finally:
        iface.close()
        if proc.poll() is None:
            proc.kill()
        out.close()
    log = (run / "log.txt").read_text()
    assert "@ message to server: NEEDINIT" in log and "@ message to server: READY" in log
    assert "Setting input geometry from sockets" in log and "Setting temperatures in LattCoo from array received via sockets" in log
    assert (run / "I00000000.xyz").exists() and len((run / "I00000000.xyz").read_text().splitlines()) > 100


def test_execpath_layout_with_reference_setup(tmp_path, monkeypatch):
    """pymulskips needs `execpath` to hold modules/defsystem.f (whose LenX line it rewrites), main.f (whose port line it
    rewrites), a makefile whose `make clean; make | tee make.log` says "Compilation terminated", and mulskips.e beside them
    (pymulskips/setuprun.py:414-420,458-553).  mulskips_b200/execpath is that directory: the reference's OWN
    setup_mulskips_src() is run on a copy of it, then the executable is launched the way run_mulskips() launches it
    (`<execpath>/mulskips.e | tee log.txt` in the run directory) and frame 0 must come out at the rewritten box."""
    import shutil
    p = os.path.join(REF, "pymulskips", "setuprun.py")
    if not os.path.exists(p):
        pytest.skip("reference tree not mounted")
    spec = importlib.util.spec_from_file_location("pymulskips_setuprun", p)
    sr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sr)
    execpath = tmp_path / "execpath"
    shutil.copytree(os.path.join(ROOT, "mulskips_b200", "execpath"), execpath, ignore=shutil.ignore_patterns("mulskips.e", "make.log"))
    monkeypatch.setenv("MSKC_ROOT", os.path.join(ROOT, "mulskips_b200"))      # the copy lives outside the package (see the makefile)
    monkeypatch.setenv("PWD", str(tmp_path))                                  # compile_mulskips() returns to $PWD
    monkeypatch.chdir(tmp_path)
    sr.setup_mulskips_src(str(execpath), 96, 96, 240, socketport=31415)
    assert "LenX=96, LenY=96, LenZ=240" in (execpath / "modules" / "defsystem.f").read_text()
    assert "port = 31415" in (execpath / "main.f").read_text()
    assert "Compilation terminated" in (execpath / "make.log").read_text()
    assert (execpath / "mulskips.e").exists()
    run = tmp_path / "run"
    run.mkdir()
    from cases import make_start
    make_start("surface", str(run / "start.dat"), lens=(48, 0, 0), maxiter=1000, outmolmol=1000)
    subprocess.run(f"{execpath}/mulskips.e | tee log.txt;", shell=True, cwd=run, capture_output=True)
    f0 = run / "I00000000.xyz"
    assert f0.exists()
    an = _analyze()
    sf = 0.38583333333333333
    assert an.get_box_sides(str(f0)) == pytest.approx([96 * sf, 96 * sf, 240 * sf], rel=1e-6)
    assert int(open(f0).readline().split()[0]) == 2 * (96 // 12) * (96 // 12)           # one filled surface layer of the 96 x 96 box
