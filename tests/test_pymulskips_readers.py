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
