"""Test-case helpers: reference start.dat files (verbatim copies under tests/golden/) and reduced variants."""
from __future__ import annotations

import os

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
CASES = ["Si-SL", "surface", "cube", "cube-zigzag", "sphere", "inverted-pyramid", "finfet", "CVD-Si-SiH2", "CAD", "LA"]
# further reference input files kept as fixtures (no golden output exists for them in the reference):
#   start-SiGeBdoped.dat = examples/ex5-CVD-IO/CVD-SiGeBdoped/start.dat (3 crystal + 2 coverage species, 288 evaporation classes)


def start_path(case):
    return os.path.join(GOLDEN, f"start-{case}.dat")


def make_start(case, dst, lens=None, maxiter=None, outmolmol=None, runtype=None, idum=None, ptranszig=None, maxtime=None, outtime=None, paths=None,
               alloyfraction=None, initst=None):
    """Copy tests/golden/start-<case>.dat to dst with some of the positional lines replaced.
    The grammar is the one of mulskips-source/main.f:83-212."""
    lines = open(start_path(case)).read().split("\n")
    ncov = int(lines[0].split()[1])
    i = 2 + (1 if ncov > 0 else 0)          # InitSt line
    if initst is not None:                  # another builder on the same table (main.f:225-295)
        lines[i] = f"{initst} ! InitSt"
    initst = lines[i].split()[0]
    i += 1
    if initst in ("IN", "SH"):
        if paths:
            lines[i] = paths[0]
        i += 1
    elif initst == "LA":
        if paths:
            lines[i], lines[i + 1] = paths[0], paths[1]
        i += 9
    else:
        if lens is not None:
            lines[i] = " ".join(str(v) for v in lens) + " ! Len"
        i += 1
    if lines[1].split()[:2] == ["14", "32"]:   # SiGe: the Ge-fraction line follows (main.f:142-149)
        if alloyfraction is not None:
            lines[i] = f"{alloyfraction} ! Ge fraction in substrate"
        i += 1
    if ptranszig is not None:
        lines[i] = f"{ptranszig} ! PtransZig"
    i += 2                                  # PtransZig, sf
    strat = lines[i].split()[0]
    i += 1
    if strat == "Iter":
        if maxiter is not None:
            lines[i] = f"{maxiter} ! IterMax"
        if outmolmol is not None:
            lines[i + 1] = f"{outmolmol} ! OutMolMol"
    else:
        if maxtime is not None:
            lines[i] = f"{maxtime!r} ! TotTime"
        if outtime is not None:
            lines[i + 1] = f"{outtime!r} ! OutTime"
    i += 2
    if runtype is not None:
        lines[i] = f"{runtype} ! RunType"
    if idum is not None:
        lines[i + 1] = f"{idum} ! seed"
    with open(dst, "w") as f:
        f.write("\n".join(lines))
    return dst


def make_la_case(dirpath, box=(96, 96, 96), lennuc=12.0, maxiter=10**9):
    """A small synthetic laser-annealing case in the reference's file formats (initial/SetCAD.f, initial/SetT.f):
    LattGeo (all cells crystal candidates; InitSt LA carves a spherical liquid nucleus below the top air gap) and
    TempStat (integer kelvin, 1655..1661 like the regression map), plus start.dat = start-LA.dat with the paths,
    nucleus radius and iteration limits replaced.  Returns the start.dat path."""
    import numpy as np
    lx, ly, lz = box
    geo, tmp = os.path.join(dirpath, "LattGeo.dat"), os.path.join(dirpath, "TempStat.dat")
    with open(geo, "w") as f:
        row = " ".join(["1"] * (lx * ly))
        for _ in range(lz):
            f.write(row + "\n")
    rng = np.random.default_rng(7)
    with open(tmp, "w") as f:
        for z in range(lz):
            f.write(" ".join(str(v) for v in (1655 + rng.integers(0, 7, lx * ly))) + "\n")
    lines = open(start_path("LA")).read().split("\n")
    lines[3], lines[4] = geo, tmp
    lines[10] = f"{lennuc} ! LenNuc"
    i = next(k for k, l in enumerate(lines) if l.split() and l.split()[0] == "Iter")
    lines[i + 1] = f"{maxiter} ! IterMax"
    lines[i + 2] = f"{maxiter} ! OutMolMol"
    lines[i + 5] = "F ! Save final state"
    del lines[i + 6]                      # the checkpoint file name line
    lines[i + 6] = "T ! Save final Coor"
    dst = os.path.join(dirpath, "start.dat")
    with open(dst, "w") as f:
        f.write("\n".join(lines))
    return dst


def make_cad_case(dirpath, box=(96, 96, 120), maxiter=10**9):
    """A small synthetic CVD-on-nanostructure case in the reference's file format (initial/SetCAD.f): a crystal ridge
    (value 1) on a crystal substrate, a wall block (value 10) beside it, empty (0) above; one z plane per record.
    start.dat = start-CAD.dat (Si + H coverage table) with the path and iteration limits replaced."""
    import numpy as np
    lx, ly, lz = box
    geo = np.zeros((lz, ly, lx), np.int32)
    geo[:36] = 1                                    # substrate
    geo[36:60, :, 24:60] = 1                        # ridge along y
    geo[30:72, :, 72:90] = 10                       # wall block (mask)
    path = os.path.join(dirpath, "LattGeo.dat")
    with open(path, "w") as f:
        for z in range(lz):
            f.write(" ".join(str(v) for v in geo[z].ravel()) + "\n")
    return make_start("CAD", os.path.join(dirpath, "start.dat"), maxiter=maxiter, outmolmol=maxiter, paths=[path])


def make_sige_case(dirpath, len1=48, fraction=0.3, maxiter=10**9):
    """Si(1-x)Ge(x) (001) slab, InitSt SG (initial/SetSiGex.f): the two-species PVD table of start-surface.dat read as
    (Si, Ge), walls, random alloy substrate; the Ge-fraction line follows the Len line (main.f:142-149)."""
    lines = open(start_path("surface")).read().split("\n")
    lines[1] = "14 32 ! list of elements of the crystal"
    lines[2] = "SG ! Initstat"
    lines[3] = f"{len1} 0 0 ! Len1 Len2 Len3"
    lines.insert(4, f"{fraction} ! Ge fraction in the substrate")
    i = next(k for k, l in enumerate(lines) if l.split() and l.split()[0] == "Iter")
    lines[i + 1] = f"{maxiter} ! IterMax"
    lines[i + 2] = f"{maxiter} ! OutMolMol"
    dst = os.path.join(dirpath, "start.dat")
    with open(dst, "w") as f:
        f.write("\n".join(lines))
    return dst


def make_st_case(dirpath, lens=(20, 45, 90), angle=54.7, ttop=1080.0, tbottom=920.0, maxiter=10**9, runtype="R",
                 temps=(900, 950, 1000, 1100)):
    """Silicon trench with a temperature ramp along z and tabulated rates, InitSt ST (initial/SetTrenchSi.f,
    mcevent/SetProbabilityFromField.f, mcevent/InterpProbFromField.f): Si + H coverage, the probability-table file
    holds one block per tabulated temperature (the CVD-Si-SiH2 regression table scaled by an Arrhenius-like factor),
    start.dat ends with the table path instead of the rate lines.  ttop == tbottom gives the fixedT variant."""
    import math
    base_e = [("10", 2226.5697760341627), ("11", 981.1791875933076), ("12", 432.3747715110884), ("13", 190.53394670735102),
              ("20", 34.804675968079124), ("21", 15.337324730794487), ("22", 6.7586760501256276),
              ("30", 0.010861139302185713), ("31", 0.004786162082841824)]
    out = [str(len(temps)), "18"]
    for t in temps:
        f = math.exp(-3.0 * (1000.0 / t - 1.0))
        out.append(str(t))
        out += [f"1 {d} {v * f!r}" for d, v in base_e]
        out += [f"1 {c} {85.49860227606791 * (1.5 - 0.5 * f)!r}" for c in (1, 2, 3)]
        out += [f"2 {c} {287.0271459302869 / f!r}" for c in (1, 2, 3)]
        out += [f"2 {c} {40.27023414006161 * f!r}" for c in (1, 2, 3)]
    table = os.path.join(dirpath, "probtable.dat")
    with open(table, "w") as fh:
        fh.write("\n".join(out) + "\n")
    start = "\n".join([
        "1 1 ! NCrystal NCov", "14 ! crystal species", "1 ! coverage species", "ST ! InitSt",
        f"{lens[0]} {lens[1]} {lens[2]} {angle!r} {ttop!r} {tbottom!r} ! Len1 Len2 Len3 angle Ttop Tbottom",
        "1.0 ! PtransZig", "0.4525 ! sf", "Iter", f"{maxiter} ! IterMax", f"{maxiter} ! OutMolMol", f"{runtype} ! RunType",
        "9117116 ! IDUM", "F ! Save final state", "F ! Save final Coor", table, ""])
    dst = os.path.join(dirpath, "start.dat")
    with open(dst, "w") as fh:
        fh.write(start)
    return dst


def make_la_tree_case(dirpath, box=(96, 96, 96), lennuc=18.0, maxiter=10**9, sige=True, tbase=1480, homogeneous="F", fraction=0.3, lensige=25.0,
                      sockets=False):
    """A laser-annealing case for the LaserAnnealing fork (mulskips-source-LaserAnnealing/main.f:380-520 grammar): one
    Tm / P0 / sigma / Tflex value per crystal species, Si(1-x)Ge(x) with the Ge-fraction and SiGe-thickness lines, the
    18-line two-species evaporation table (energies per bond: Si 0.96 eV, Ge 0.70 eV) and 6 deposition lines.  Geometry
    and temperature files in the formats of initial/SetCAD.f / initial/SetT.f (what the socket arrays carry)."""
    import numpy as np
    lx, ly, lz = box
    geo, tmp = os.path.join(dirpath, "LattGeo.dat"), os.path.join(dirpath, "TempStat.dat")
    z0 = lz - int(18.2 / 0.4525)                      # top of the sample: air (0) above, as the FEM side sends it
    with open(geo, "w") as f:
        for z in range(lz):
            f.write(" ".join(["1" if z <= z0 else "0"] * (lx * ly)) + "\n")
    rng = np.random.default_rng(11)
    with open(tmp, "w") as f:
        for z in range(lz):
            f.write(" ".join(str(v) for v in (tbase + rng.integers(0, 7, lx * ly))) + "\n")
    ncry = 2 if sige else 1
    if sockets:          # geometry, temperature and phases travel through the socket (la/main.f:160-345): the three paths are 'None'
        geo = tmp = "None"
    L = [f"{ncry} 0 ! NCrystal NCov", "14 32 ! crystal species" if sige else "14 ! crystal species", "LA ! InitSt", geo, tmp,
         ("1688 1210" if sige else "1688") + " ! Tm", ("1.33e17 2.596e16" if sige else "1.33e17") + " ! P0",
         ("280 210" if sige else "280") + " ! sigma", ("1080 900" if sige else "1080") + " ! Tflex",
         "18.2 ! LenVac", f"{lennuc} ! LenNuc", f"{homogeneous} ! Homogeneous"]
    if sige:
        L += [f"{fraction} ! Ge fraction", f"{lensige} ! LenSiGe"]
    L += ["1.0 ! PtransZig", "0.4525 ! sf", "Iter", f"{maxiter} ! IterMax", f"{maxiter} ! OutMolMol", "R ! RunType", "9117116 ! IDUM",
          "F ! Save final state", "T ! Save final Coor", "None" if sockets else "LattStat.dat"]
    e = {1: 0.96, 2: 0.70}
    if sige:
        for sp in (1, 2):
            for n1 in range(4):
                for n2 in range(4):
                    if 1 <= n1 + n2 <= 3:
                        L.append(f"{sp} {n1}{n2} {0.5 * (e[sp] * (n1 + n2) + e[1] * n1 + e[2] * n2)!r}")
        for sp, sc in ((1, 1.0), (2, 0.72)):
            L += [f"{sp} {c} {v * sc!r}" for c, v in ((1, 1.86), (2, 1.92), (3, 1.98))]
    else:
        L += ["1 1 0.96", "1 2 1.92", "1 3 2.88", "1 1 1.86", "1 2 1.92", "1 3 1.98"]
    dst = os.path.join(dirpath, "start.dat")
    with open(dst, "w") as f:
        f.write("\n".join(L) + "\n")
    return dst


def make_shape_case(dirpath, box=(144, 144, 144), lens=(48, 48, 48), maxiter=10**9, table="cube"):
    """InitSt SH (initial/SetShape.f): a shape file "sym x y z" per filled site.  The shape is the parallelepiped of
    initial/SetNC.f with the 0- and 1-coordinated atoms peeled off (SetShape.f:98 expects that; coordination counted on the
    full first shell, as SetShape does), written in z,y,x order; start.dat = the cube's table with InitSt SH + the path."""
    lx, ly, lz = box
    l1, l2, l3 = lens
    sites = {}
    for z in range(lz // 2 - l3 // 2, lz // 2 + l3 // 2):
        for y in range(ly // 2 - l2 // 2, ly // 2 + l2 // 2):
            for x in range(lx // 2 - l1 // 2, lx // 2 + l1 // 2):
                if x % 3 or y % 3 or z % 3:
                    continue
                x1, y1, z1 = x // 3, y // 3, z // 3
                if (x1 % 2) != (y1 % 2) or (x1 % 2) != (z1 % 2):
                    continue
                m = (x1 + y1 + z1) % 4
                if m in (0, 3):
                    sites[(x, y, z)] = "Si" if m == 0 else "C"
    shell = [(i, j, k) for k in range(-6, 7) for j in range(-6, 7) for i in range(-6, 7) if i * i + j * j + k * k == 27]
    while True:
        weak = [s for s in sites if sum(((s[0] + d[0]) % lx, (s[1] + d[1]) % ly, s[2] + d[2]) in sites for d in shell) <= 1]
        if not weak:
            break
        for s in weak:
            del sites[s]
    shape = os.path.join(dirpath, "shape.xyz")
    with open(shape, "w") as f:
        f.write(f"{len(sites)}\nparallelepiped, 0- and 1-coordinated atoms removed\n")
        for (x, y, z), sym in sorted(sites.items(), key=lambda kv: (kv[0][2], kv[0][1], kv[0][0])):
            f.write(f"{sym} {x} {y} {z}\n")
    dst = os.path.join(dirpath, "start.dat")
    make_start(table, dst, initst="SH", maxiter=maxiter, outmolmol=maxiter)
    lines = open(dst).read().split("\n")
    i = next(k for k, l in enumerate(lines) if l.split() and l.split()[0] == "SH")
    lines[i + 1] = shape                       # the Len line of the cube's file becomes the path line (main.f:116-117)
    with open(dst, "w") as f:
        f.write("\n".join(lines))
    return dst


def make_sic_cvd_case(dirpath, lens=(48, 0, 0), ptranszig=0.8, maxiter=10**9, runtype="R"):
    """BASELINE config 2: 3C-SiC(001) CVD with hydrogen coverage and hexagonal-phase nucleation on a flat F
    slab -- NCrystal 2 (Si, C), NCov 1 (H): the 50-line evaporation table = the 18 values of start-surface.dat for every
    crystal census (n_Si, n_C), halved per adsorbed-H neighbour; deposition = start-surface.dat; absorption 1e-5 and
    desorption 1e-4 per site (synthetic, on the SiC table's time unit: the H rates of start-CVD-Si-SiH2.dat belong to a
    table 1e5 times faster and, scaled, leave the surface H-passivated and etching), so that the crystal grows with a few
    per cent of H coverage."""
    src = open(start_path("surface")).read().split("\n")
    evap = {}
    dep = []
    for l in src:
        t = l.split()
        if len(t) >= 3 and t[0] in ("1", "2") and len(t[1]) == 2 and t[1].isdigit():
            evap[(int(t[0]), int(t[1][0]), int(t[1][1]))] = float(t[2])
        elif len(t) >= 3 and t[0] in ("1", "2") and t[1] in ("1", "2", "3") and "PtransD" in l:
            dep.append(f"{t[0]} {t[1]} {t[2]}")
    assert len(evap) == 18 and len(dep) == 6
    L = ["2 1 ! NCrystal NCov", "14 6 ! crystal species", "1 ! coverage species (H)", "F ! InitSt", f"{lens[0]} {lens[1]} {lens[2]} ! Len1 Len2 Len3",
         f"{ptranszig} ! PtransZig", "0.38583333333333333 ! sf", "Iter", f"{maxiter} ! IterMax", f"{maxiter} ! OutMolMol", f"{runtype} ! RunType",
         "9117116 ! IDUM", "F ! Save final state", "F ! Save final Coor"]
    n = 0
    for sp in (1, 2):
        for (s, n1, n2), v in sorted(evap.items()):
            if s != sp:
                continue
            for nh in range(0, 4 - (n1 + n2) + 1):
                L.append(f"{sp} {n1}{n2}{nh} {v * 0.5 ** nh!r}")
                n += 1
    assert n == 50
    L += dep
    L += [f"3 {c} 1e-05" for c in (1, 2, 3)]
    L += [f"3 {c} 0.0001" for c in (1, 2, 3)]
    dst = os.path.join(dirpath, "start.dat")
    with open(dst, "w") as f:
        f.write("\n".join(L) + "\n")
    return dst
