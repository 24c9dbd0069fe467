"""bench.py's GPU arm end to end on tiny boxes against the SIMT-emulated library (tests/emu): the legs, the sweep cap,
the strong-scaling point with its head-room guard and the shape of the JSON line, checked where there is no GPU.  The
emulated library is test infrastructure (g++ build of the same CUDA sources); its numbers mean nothing."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu_lib():
    r = subprocess.run(["make", "-C", EMU], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    return os.path.join(EMU, "_build", "libmskc.so")


def run_bench(emu_lib, *extra):
    env = dict(os.environ, MSKC_LIB=emu_lib)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    cmd = [sys.executable, os.path.join(EMU, "bench_under_emu.py"), "--box", "192", "192", "96", "--steps", "2", "--warmup", "1", "--sweeps", "4",
           "--no-cpu-baseline", *extra]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry exactly one line, the JSON"
    return json.loads(lines[0])


def test_bench_line_and_strong_point(emu_lib):
    d = run_bench(emu_lib, "--strong-box", "192", "192", "48")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "strong_1e9", "dense_passes"):
        assert k in d, k
    assert d["metric"] == "kmc_events_per_s" and d["value"] > 0 and d["gpu_launches"] > 0
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert len(d["dense_passes"]) == 8 and not any("error" in p for p in d["dense_passes"])
    assert {p["kernel"] for p in d["dense_passes"]} >= {"frame_extract_kernel<1>", "antisite_kernel", "frame_extract_kernel", "census_kernel"}
    s = d["strong_1e9"]
    assert s["error"] is None and s["value"] > 0 and 1 <= s["steps"] <= 6
    tops = s["highest_atom_plane_before_each_step"]
    assert len(tops) >= s["steps"] + s["warmup"]
    # every step that ran started with at least the head room left
    assert all(t <= 48 - s["head_room_planes"] for t in tops[:s["steps"] + s["warmup"]])


def test_strong_point_stops_before_the_box_is_outgrown(emu_lib):
    """a box with almost no room above the slab: the point reports why it has no value instead of running into the top plane"""
    d = run_bench(emu_lib, "--strong-box", "192", "192", "39", "--no-dense", "--no-la-cycle")
    s = d["strong_1e9"]
    assert s["value"] is None and s["steps"] == 0
    assert "outgrown" in s["error"]
    assert d["value"] > 0          # the headline legs are unaffected
