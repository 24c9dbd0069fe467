#!/usr/bin/env python3
"""Active-lane statistics per source line of one kernel from an .ncu-rep: how many warp-instructions a line issues
and with how many threads on average (SIMT efficiency of the team scheme).
usage: tools/ncu_active.py rep.ncu-rep lib.so kernel_substring [topN]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, lib, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[hdr_i[0]]; end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
col = {n: i for i, n in enumerate(hdr)}
inst = {}
for r in rows[hdr_i[0] + 1:end]:
    if len(r) < len(hdr): continue
    try: a = int(r[0], 16)
    except ValueError: continue
    inst[a] = r
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
ex = collections.Counter(); th = collections.Counter(); tot = tth = 0
hist = collections.Counter()
for a, r in inst.items():
    l = amap.get(a - base)
    e = float(r[col["Instructions Executed"]] or 0); t = float(r[col["Thread Instructions Executed"]] or 0)
    ex[l] += e; th[l] += t; tot += e; tth += t
    if e: hist[min(int(t / e // 4) * 4, 28)] += e
print(f"warp-instructions {tot:.0f}, avg active threads {tth / tot:.2f}")
print("histogram of avg active threads per SASS instruction (share of issued):", {f"{k}-{k+3}": f"{100 * v / tot:.1f}%" for k, v in sorted(hist.items())})
src = {}
def text(l):
    if not l: return ""
    f, n = l
    p = os.path.join(os.path.dirname(os.path.abspath(lib)), "..", "csrc", f)
    if f not in src: src[f] = open(p).read().split("\n") if os.path.exists(p) else []
    return src[f][n - 1].strip()[:110] if 0 < n <= len(src[f]) else ""
print("top lines by issued warp-instructions (share, avg active threads):")
for l, e in ex.most_common(topn):
    print(f"  {100 * e / tot:5.1f}%  {th[l] / e:5.1f}  {l}  {text(l)}")
