"""tests/analyze_port.py (the frame readers the GPU statistics tests use) against the reference's own pymulskips.analyze,
on frames the oracle writes for a CVD case (H coverage, vacancies) and a PVD case.  CPU only; needs /root/reference."""
import importlib.util
import os
import subprocess

import pytest

import analyze_port as ap
from cases import make_sic_cvd_case, make_start

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle", "mulskips_oracle.e")
REF_ANALYZE = "/root/reference/pymulskips/analyze.py"


@pytest.fixture(scope="module")
def an():
    if not os.path.exists(REF_ANALYZE):
        pytest.skip("reference tree not mounted")
    import sys
    import types
    if "matplotlib" not in sys.modules:                       # analyze_growth_rate imports pyplot even with plotting=False; not installed here
        try:
            import matplotlib.pyplot  # noqa: F401
        except ImportError:
            mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    spec = importlib.util.spec_from_file_location("pymulskips_analyze", REF_ANALYZE)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _time_strategy(path, tot, out):
    lines = open(path).read().split("\n")
    i = next(k for k, l in enumerate(lines) if l.split() and l.split()[0] in ("Iter", "Time"))
    lines[i] = "Time"; lines[i + 1] = f"{tot!r} ! TotTime"; lines[i + 2] = f"{out!r} ! OutTime"
    open(path, "w").write("\n".join(lines))


@pytest.mark.parametrize("kind", ["cvd", "pvd"])
def test_port_equals_reference_readers(tmp_path, an, kind):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    if kind == "cvd":
        start = make_sic_cvd_case(str(tmp_path), lens=(48, 0, 0))
    else:
        start = make_start("surface", str(tmp_path / "start.dat"), lens=(48, 0, 0), runtype="R", ptranszig=0.8)
    _time_strategy(start, 4.0e5, 5.0e4)
    subprocess.run([ORACLE, "--box", "96", "96", "240"], cwd=tmp_path, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    files = ap.frames(str(tmp_path))
    assert len(files) >= 8 and sorted(an.read_output_files(str(tmp_path), "undercoordinated")) == files

    class MP:
        miller = 100
    for f in files[1:]:
        assert ap.get_surface_height(f) == an.get_surface_height(f)
        c = an.get_coverage(f, mp=MP())
        mine = ap.get_coverage(f)
        assert mine["H"] == c["H"][0] and mine["Cl"] == c["Cl"][0]
    gr, h, t = ap.growth_rate(str(tmp_path), Nexclude=2)
    # (the reference takes the frames in glob order, whatever the file system returns; the port sorts them by name = by time)
    orig = an.read_output_files
    an.read_output_files = lambda d, what="undercoordinated": sorted(orig(d, what))
    ref = an.analyze_growth_rate(str(tmp_path), plotting=False, Nexclude=2, method="finitediff")
    ref_gr = ref[0] if isinstance(ref, tuple) else ref
    assert gr == pytest.approx(float(ref_gr), rel=1e-12, abs=1e-300)
    vac = an.count_vacancies(str(tmp_path))
    mine = sorted(ap.count_lines(f) for f in ap.frames(str(tmp_path), "vacancies") if not f.endswith("I00000000_v.xyz"))
    assert sorted(int(v) for v in vac["tot"]) == mine
