#!/usr/bin/env python3
"""Build tests/golden/manifest.json from the reference's regression goldens
(regression-tests/input-and-references-files): sha256 + line count of I0000000{0,1}-<case>.xyz and the
numeric tokens of log-<case>.txt that are comparable across builds ('Atoms,Voids,AdAtoms',
'Iter, Time, Site', 'NatNvoidNad').  The start-<case>.dat inputs are copied verbatim next to it.
Usage: make_golden_manifest.py /root/reference tests/golden"""
import hashlib, json, os, shutil, sys
ref, out = sys.argv[1], sys.argv[2]
rtin = os.path.join(ref, "regression-tests", "input-and-references-files")
cases = ["Si-SL", "surface", "cube", "cube-zigzag", "sphere", "inverted-pyramid", "finfet", "CVD-Si-SiH2", "CAD", "LA"]
man = {}
for c in cases:
    shutil.copy(os.path.join(rtin, f"start-{c}.dat"), os.path.join(out, f"start-{c}.dat"))
    e = {}
    for k in (0, 1):
        b = open(os.path.join(rtin, f"I0000000{k}-{c}.xyz"), "rb").read()
        e[f"frame{k}_sha256"] = hashlib.sha256(b).hexdigest()
        e[f"frame{k}_lines"] = b.count(b"\n")
        e[f"frame{k}_head"] = b.split(b"\n")[0].decode()
    for line in open(os.path.join(rtin, f"log-{c}.txt")):
        t = line.split()
        if line.strip().startswith("Atoms,Voids,AdAtoms"):
            e["init_lists"] = [int(v) for v in t[1:4]]
        elif line.strip().startswith("Iter, Time, Site"):
            e["final_iter"], e["final_time"], e["final_site"] = int(t[3]), t[4], [int(v) for v in t[5:8]]
        elif line.strip().startswith("NatNvoidNad"):
            e["final_lists"] = [int(v) for v in t[1:4]]
    man[c] = e
json.dump(man, open(os.path.join(out, "manifest.json"), "w"), indent=1)
print("wrote", len(man), "cases")
