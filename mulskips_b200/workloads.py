"""Rate tables of the synthetic workloads (bench.py, scripts/), as numpy arrays in the module-array layout
(PtransE(3,0:3 x6) etc., Fortran order, mulskips-source/modules/defsystem.f:60-65).

The values are the literal table tails of the reference's regression start files
(regression-tests/input-and-references-files/start-surface.dat:14-37, start-Si-SL.dat:14-19, start-CVD-Si-SiH2.dat:15-32),
kept here in the start.dat grammar (mcevent/SetProbability.f:21-143: `species digits value`) so that the package needs
nothing outside itself.  parse_start_tables() reads the same grammar from any start.dat a caller names."""
from __future__ import annotations

import numpy as np

# (NCrystal, NCov, crystal Z list, coverage Z list, PtransZig, sf, table tail)
_TABLES = {
    "surface": (2, 0, [14, 6], [], 1.0, 0.38583333333333333, """
1 01 0.000458642486241
1 10 0.00107413320314
1 11 0.000654677677311
1 02 9.03452031205e-05
1 20 0.000717033435226
1 21 0.000142385075632
1 12 2.91989021579e-05
1 03 3.18575387165e-07
1 30 0.000518948689858
2 01 10693067.5485
2 10 336.433275007
2 11 9.14549727245
2 02 327556.955526
2 20 4.99391159317e-06
2 21 2.81716126678e-06
2 12 0.288419678521
2 03 20148.1158068
2 30 2.85295646127e-07
1 1 9.98730082416e-05
1 2 2.49682520604e-05
1 3 0.00399492032966
2 1 0.000998782318633
2 2 1.24847789829e-05
2 3 0.00199756463727
"""),
    "Si-SL": (1, 0, [14], [], 1.0, 0.4525, """
1 1 1.4037680927273852
1 2 0.0019705648581594803
1 3 2.7662160725341417e-06
1 1 0.0009258079026426269
1 2 0.0009258079026426269
1 3 0.0009258079026426269
"""),
    "CVD-Si-SiH2": (1, 1, [14], [1], 1.0, 0.4525, """
1 10 2226.5697760341627
1 11 981.1791875933076
1 12 432.3747715110884
1 13 190.53394670735102
1 20 34.804675968079124
1 21 15.337324730794487
1 22 6.7586760501256276
1 30 0.010861139302185713
1 31 0.004786162082841824
1 1 85.49860227606791
1 2 85.49860227606791
1 3 85.49860227606791
2 1 287.0271459302869
2 2 287.0271459302869
2 3 287.0271459302869
2 1 40.27023414006161
2 2 40.27023414006161
2 3 40.27023414006161
"""),
}
# evaporation lines per (NCrystal, NCov), mcevent/SetProbability.f:47-75
_N_EVAP = {(1, 0): 3, (2, 0): 18, (3, 0): 57, (1, 1): 9, (2, 1): 50, (3, 1): 150, (1, 2): 19, (2, 2): 100, (3, 2): 288, (1, 3): 34, (2, 3): 172, (3, 3): 480}


def idxE(sp, i1=0, i2=0, i3=0, i4=0, i5=0, i6=0):
    return (sp - 1) + 3 * (i1 + 4 * (i2 + 4 * (i3 + 4 * (i4 + 4 * (i5 + 4 * i6)))))


def _tables_from_tail(ncry, ncov, tail):
    """mcevent/SetProbability.f:80-140: digit j <= NCrystal -> table index j, digit NCrystal+k -> table index 3+k"""
    nE = _N_EVAP[(ncry, ncov)]
    assert len(tail) == nE + 3 * ncry + 6 * ncov, (len(tail), nE, ncry, ncov)
    PE, PD, PA, PS = np.zeros(12288), np.zeros(9), np.zeros(12), np.zeros(12)
    nd = ncry + ncov
    for l in tail[:nE]:
        sp, el, v = l.split()[:3]
        sp, el, v = int(sp), int(el), float(v)
        ind = [0] * 6
        for j in range(1, ncry + 1):
            ind[j - 1] = (el // 10 ** (nd - j)) % 10
        for j in range(ncry + 1, nd + 1):
            ind[3 + j - ncry - 1] = (el // 10 ** (nd - j)) % 10
        PE[idxE(sp, *ind)] = v
    k = nE
    for l in tail[k:k + 3 * ncry]:
        sp, el, v = l.split()[:3]
        PD[(int(sp) - 1) + 3 * (int(el) - 1)] = float(v)
    k += 3 * ncry
    for l in tail[k:k + 3 * ncov]:
        sp, el, v = l.split()[:3]
        PA[(int(sp) - ncry - 1) + 3 * int(el)] = float(v)
    k += 3 * ncov
    for l in tail[k:k + 3 * ncov]:
        sp, el, v = l.split()[:3]
        PS[(int(sp) - ncry - 1) + 3 * int(el)] = float(v)
    return PE, PD, PA, PS


def parse_start_tables(path):
    """Rate-table tail of a start.dat (mcevent/SetProbability.f:21-143) + species header (main.f:83-100)."""
    lines = [l for l in open(path).read().split("\n")]
    ncry, ncov = (int(v) for v in lines[0].split()[:2])
    listcrystal = [int(v) for v in lines[1].split()[:ncry]]
    listcov = [int(v) for v in lines[2].split()[:ncov]] if ncov else []
    ntail = _N_EVAP[(ncry, ncov)] + 3 * ncry + 6 * ncov
    body = [l for l in lines if l.strip()]
    PE, PD, PA, PS = _tables_from_tail(ncry, ncov, body[-ntail:])
    # PtransZig is the line before sf, i.e. two lines before the ExitStrategy keyword line
    zi = next(i for i, l in enumerate(lines) if l.split() and l.split()[0] in ("Iter", "Time")) - 2
    return dict(listcrystal=listcrystal, listcov=listcov, PtransE=PE, PtransD=PD, PtransAbs=PA, PtransDes=PS,
                PtransZig=float(lines[zi].split()[0]), sf=float(lines[zi + 1].split()[0]))


def rates_for(case, path=None):
    """Tables of a named workload ("surface": 3C-SiC PVD, "Si-SL": Si PVD, "CVD-Si-SiH2": Si CVD with H coverage), or of
    the start.dat at `path`."""
    if path is not None:
        w = parse_start_tables(path)
    else:
        if case not in _TABLES:
            raise KeyError(f"unknown workload {case!r}: pass path= to read the tables of a start.dat")
        ncry, ncov, lc, lv, zig, sf, tail = _TABLES[case]
        PE, PD, PA, PS = _tables_from_tail(ncry, ncov, [l for l in tail.split("\n") if l.strip()])
        w = dict(listcrystal=list(lc), listcov=list(lv), PtransE=PE, PtransD=PD, PtransAbs=PA, PtransDes=PS, PtransZig=zig, sf=sf)
    w["walls"] = len(w["listcrystal"]) == 1                 # the Si builder (FS) has walls, the SiC one (F) has none
    w["two_species"] = len(w["listcrystal"]) == 2
    return w
