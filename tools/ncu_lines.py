#!/usr/bin/env python3
"""Per-source-line executed instructions and stall samples of one kernel from an .ncu-rep (read here, no GPU).
The SASS page of the report gives counts per instruction address; nvdisasm -g of the SAME build gives address -> line.
usage: tools/ncu_lines.py rep.ncu-rep lib.so kernel_substring [topN]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, lib, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
# first kernel instance only
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[hdr_i[0]]
end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
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
sec = False; cur = None; func = "kernel"; amap = {}
for line in dis.split("\n"):
    if line.startswith("\t.section\t.text.") or line.startswith(".section .text.") or re.match(r"\s*//-+ \.text\.", line):
        sec = kname in line
        continue
    if not sec: continue
    m = re.match(r'\s+//## File "([^"]+)", line (\d+)', line)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*(\$[\w$]+|\.L_x_\d+|\w+):\s*$", line)
    if m and "$" in m.group(1): func = m.group(1).split("$")[-1]
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", line)
    if m: amap[int(m.group(1), 16)] = (func, cur)
ex = collections.Counter(); st = collections.Counter(); stall_kind = collections.defaultdict(collections.Counter); fex = collections.Counter(); fst = collections.Counter()
kinds = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
tot_ex = tot_st = 0
for a, r in inst.items():
    f, l = amap.get(a - base, ("?", None))
    e = int(float(r[col["Instructions Executed"]] or 0)); s = int(float(r[col["# Samples"]] or 0))
    ex[l] += e; st[l] += s; fex[f] += e; fst[f] += s; tot_ex += e; tot_st += s
    for k in kinds:
        v = r[col[k]]
        if v: stall_kind[l][k] += int(float(v))
print(f"total warp-instructions {tot_ex}, samples {tot_st}, mapped addresses {len(amap)} / {len(inst)}")
print("by function (executed%, samples%):")
for f, e in fex.most_common(15): print(f"  {100 * e / tot_ex:5.1f}% {100 * fst[f] / max(tot_st, 1):5.1f}%  {f}")
src = {}
def text(l):
    if not l: return ""
    f, n = l
    p = os.path.join(os.path.dirname(os.path.abspath(lib)), "..", "csrc", f)
    if f not in src: src[f] = open(p).read().split("\n") if os.path.exists(p) else []
    return src[f][n - 1].strip()[:100] if 0 < n <= len(src[f]) else ""
print("top lines by executed instructions:")
for l, e in ex.most_common(topn):
    print(f"  {100 * e / tot_ex:5.1f}% ex {100 * st[l] / max(tot_st, 1):5.1f}% smp  {l}  {text(l)}")
print("top lines by stall samples:")
for l, s in st.most_common(topn):
    top = ",".join(f"{k[6:]}:{v}" for k, v in stall_kind[l].most_common(3))
    print(f"  {100 * s / max(tot_st, 1):5.1f}% smp {100 * ex[l] / tot_ex:5.1f}% ex  {l}  [{top}]  {text(l)}")
if os.environ.get("REGIONS"):
    # executed instructions by source region: "file:lo-hi=name,..."
    regs = []
    for spec in os.environ["REGIONS"].split(","):
        loc, name = spec.split("="); f, rng = loc.split(":"); lo, hi = rng.split("-")
        regs.append((f, int(lo), int(hi), name))
    agg = collections.Counter(); aggs = collections.Counter()
    for l, e in ex.items():
        name = "other"
        if l:
            for f, lo, hi, nm in regs:
                if l[0] == f and lo <= l[1] <= hi: name = nm; break
            else: name = l[0]
        agg[name] += e; aggs[name] += st[l]
    print("by region (executed%, samples%):")
    for nm, e in agg.most_common(): print(f"  {100 * e / tot_ex:5.1f}% {100 * aggs[nm] / max(tot_st, 1):5.1f}%  {nm}")
