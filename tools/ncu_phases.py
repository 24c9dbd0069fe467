#!/usr/bin/env python3
"""Executed warp-instructions, stall samples and active lanes of one kernel from an .ncu-rep, grouped by source REGION
(file + line range -> phase of the algorithm).  Same address -> line mapping as tools/ncu_lines.py.
usage: tools/ncu_phases.py rep.ncu-rep lib.so kernel_substring"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, lib, kname = sys.argv[1:4]
# (file, first line, last line, label) -- line numbers of the build that was profiled
REGIONS = [
    ("mskc_parallel.cuh", 78, 95, "Philox4x32-10"),
    ("mskc_parallel.cuh", 96, 116, "list_of / pack_ce"),
    ("mskc_parallel.cuh", 117, 165, "tile lists: put / erase"),
    ("mskc_parallel.cuh", 166, 232, "tile budget / load"),
    ("mskc_parallel.cuh", 233, 300, "selection: chunk sums, scan, walk, decode"),
    ("mskc_parallel.cuh", 301, 340, "tile finish"),
    ("mskc_parallel.cuh", 341, 375, "kernel prologue / loop"),
    ("mskc_coop.cuh", 39, 110, "coop helpers: fld / nbr / flags"),
    ("mskc_coop.cuh", 111, 196, "coop_rate (full census)"),
    ("mskc_coop.cuh", 197, 222, "flag_rate"),
    ("mskc_coop.cuh", 223, 290, "event: species draw, S record"),
    ("mskc_coop.cuh", 291, 336, "event: bond fixing + IF_Arm_Choice"),
    ("mskc_coop.cuh", 337, 436, "event phase A: the four neighbours"),
    ("mskc_coop.cuh", 437, 462, "event: extra cells (second bond / old first neighbour)"),
    ("mskc_coop.cuh", 463, 490, "event: ballots + S request"),
    ("mskc_coop.cuh", 491, 503, "event: rate call"),
    ("mskc_coop.cuh", 504, 525, "event: commit"),
    ("mskc_device.cuh", 1, 170, "device.cuh: word fields / cell index"),
    ("mskc_device.cuh", 171, 320, "device.cuh: geometry + rate tables"),
]
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[hdr_i[0]]; end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
col = {n: i for i, n in enumerate(hdr)}
inst = {}
for r in rows[hdr_i[0] + 1:end]:
    if len(r) < len(hdr): continue
    try: inst[int(r[0], 16)] = r
    except ValueError: pass
base = min(inst)
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
sec = False; cur = None; amap = {}
for line in dis.split("\n"):
    if line.startswith("\t.section\t.text.") or line.startswith(".section .text.") or re.match(r"\s*//-+ \.text\.", line):
        sec = kname in line; continue
    if not sec: continue
    m = re.match(r'\s+//## File "([^"]+)", line (\d+)', line)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", line)
    if m: amap[int(m.group(1), 16)] = cur
def region(l):
    if not l: return "(unmapped)"
    for f, a, b, lab in REGIONS:
        if l[0] == f and a <= l[1] <= b: return lab
    return f"other: {l[0]}"
ex = collections.Counter(); th = collections.Counter(); st = collections.Counter()
tc = "Thread Instructions Executed" if "Thread Instructions Executed" in col else None
for a, r in inst.items():
    lab = region(amap.get(a - base))
    e = int(float(r[col["Instructions Executed"]] or 0))
    ex[lab] += e; st[lab] += int(float(r[col["# Samples"]] or 0))
    if tc: th[lab] += int(float(r[col[tc]] or 0))
te, ts = sum(ex.values()), sum(st.values())
print(f"total warp-instructions {te}, stall samples {ts}")
print(f"{'issued':>7} {'samples':>8} {'lanes':>6}  region")
for lab, e in ex.most_common():
    print(f"{100 * e / te:6.1f}% {100 * st[lab] / max(ts, 1):7.1f}% {th[lab] / e if e and tc else 0:6.1f}  {lab}")
