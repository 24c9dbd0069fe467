"""The host program's builders against the oracle's, WITHOUT a GPU and beyond frame 0: `mulskips.e --dump-state f` writes
what it is about to upload -- the three particle lists in list order with their bonds, event codes and rates (raw bits),
the position of the ran3 stream, the species counters and a hash of LattCoo.  All of it must equal the oracle's state after
its own builder ran on the same start.dat: with equal initial state, lattice and stream the serial-parity loop on the GPU
reproduces the oracle event for event (tests/test_parity_gpu.py)."""
import os
import subprocess

import numpy as np
import pytest

from cases import make_cad_case, make_la_case, make_la_tree_case, make_shape_case, make_sic_cvd_case, make_sige_case, make_st_case, make_start
from oracle_binding import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "mulskips_b200", "bin", "mulskips.e")

# name -> (start.dat maker, box, la_tree)
CASES = {
    "Si-SL": (lambda d: make_start("Si-SL", f"{d}/start.dat", lens=(216, 0, 0)), (96, 96, 240), False),
    "surface": (lambda d: make_start("surface", f"{d}/start.dat", lens=(48, 0, 0)), (96, 96, 240), False),
    "cube-zigzag": (lambda d: make_start("cube-zigzag", f"{d}/start.dat", lens=(48, 48, 48)), (144, 144, 144), False),
    "sphere": (lambda d: make_start("sphere", f"{d}/start.dat", lens=(60, 0, 0)), (144, 144, 144), False),
    "inverted-pyramid": (lambda d: make_start("inverted-pyramid", f"{d}/start.dat", lens=(30, 93, 0)), (96, 96, 144), False),
    "finfet": (lambda d: make_start("finfet", f"{d}/start.dat", lens=(48, 72, 120, 48)), (96, 96, 168), False),
    "CVD-Si-SiH2": (lambda d: make_start("CVD-Si-SiH2", f"{d}/start.dat", lens=(48, 0, 0)), (96, 96, 96), False),
    "SiGeBdoped": (lambda d: make_start("SiGeBdoped", f"{d}/start.dat", lens=(48, 0, 0)), (96, 96, 240), False),
    "SiC-CVD-H (BASELINE config 2)": (lambda d: make_sic_cvd_case(d, lens=(96, 0, 0)), (240, 240, 288), False),
    "SiGe": (lambda d: make_sige_case(d, 48, 0.3), (96, 96, 96), False),
    "CAD": (lambda d: make_cad_case(d, (96, 96, 120)), (96, 96, 120), False),
    "LA": (lambda d: make_la_case(d, (96, 96, 96)), (96, 96, 96), False),
    "LA-fork": (lambda d: make_la_tree_case(d, (96, 96, 96)), (96, 96, 96), True),
    "ST": (lambda d: make_st_case(d), (96, 48, 144), False),
    "SH": (lambda d: make_shape_case(d), (144, 144, 144), False),
    "E": (lambda d: make_start("cube", f"{d}/start.dat", lens=(60, 24, 24), initst="E"), (96, 96, 144), False),
    "A": (lambda d: make_start("surface", f"{d}/start.dat", lens=(72, 0, 0), initst="A"), (144, 144, 120), False),
    "T": (lambda d: make_start("cube", f"{d}/start.dat", lens=(20, 45, 90), initst="T"), (96, 48, 144), False),
    "K": (lambda d: make_start("inverted-pyramid", f"{d}/start.dat", lens=(30, 93, 0), initst="K"), (96, 96, 144), False),
    "Fa": (lambda d: make_start("surface", f"{d}/start.dat", lens=(48, 72, 48), initst="Fa"), (96, 96, 144), False),
}


def _parse(path):
    out = {"mc": [], "bulk": [], "void": []}
    for line in open(path):
        t = line.split()
        if t[0] in ("mc", "bulk", "void"):
            out[t[0]].append(t[1:])
        else:
            out[t[0]] = t[1:]
    return out


@pytest.mark.parametrize("name", list(CASES))
def test_host_state_equals_oracle(tmp_path, name):
    if not os.path.exists(EXE):
        import __graft_entry__ as g
        g.build()
    maker, box, la_tree = CASES[name]
    start = maker(str(tmp_path))
    run = tmp_path / "run"
    run.mkdir()
    (run / "start.dat").write_text(open(start).read())
    args = [EXE, "--box", *[str(v) for v in box], "--dump-state", str(run / "state.txt")] + (["--la-tree"] if la_tree else [])
    subprocess.run(args, cwd=run, capture_output=True, text=True)          # without a GPU it stops at mskc_create, after the dump
    d = _parse(run / "state.txt")
    orc = Oracle(*box, start, la_tree=la_tree)
    orc.setup()
    s, l = orc.sizes(), orc.lists()
    assert [int(v) for v in d["sizes"]] == [s["NumAdAtom"], s["NumAtoms"], s["NumVoids"]]
    L = orc.lattcoo().astype(np.uint32).reshape(-1)
    h = np.uint64(1469598103934665603)
    by = L.view(np.uint8)
    with np.errstate(over="ignore"):                                       # FNV-1a over the little-endian bytes, vectorised per byte is not possible: chunked python loop
        hv = int(h)
        for b in by.tobytes():
            hv = ((hv ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert int(d["lattice_fnv"][0], 16) == hv, "LattCoo differs"
    mc = np.array([[int(v) for v in r[:16]] for r in d["mc"]], dtype=np.int64).reshape(-1, 16)
    assert np.array_equal(mc[:, :3], l["mc_xyz"]) and np.array_equal(mc[:, 3:15].reshape(-1, 4, 3), l["mc_nn"]) and np.array_equal(mc[:, 15], l["mc_event"])
    rates = np.array([int(r[16], 16) for r in d["mc"]], dtype=np.uint64)
    assert np.array_equal(rates, l["mc_rate"].view(np.uint64)), "rates differ"
    for tag, key in (("bulk", "bulk"), ("void", "void")):
        if len(d[tag]):
            a = np.array([[int(v) for v in r] for r in d[tag]], dtype=np.int64)
            assert np.array_equal(a[:, :3], l[f"{key}_xyz"]) and np.array_equal(a[:, 3:].reshape(-1, 4, 3), l[f"{key}_nn"]), tag
    cc, cv = orc.counts()
    assert [int(v) for v in d["counts"]] == list(cc) + list(cv)
    st, counter, idum = orc.rng()
    assert [int(v) for v in d["ran3"]] == [idum] + [int(v) for v in st]
    orc.close()
