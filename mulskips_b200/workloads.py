"""Literal rate tables of the reference's regression start files, as numpy arrays in the module-array
layout (PtransE(3,0:3 x6) etc., Fortran order) -- used to set up synthetic workloads without parsing
a start.dat.  Values are those of regression-tests/input-and-references-files/start-<case>.dat."""
from __future__ import annotations

import os

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def idxE(sp, i1=0, i2=0, i3=0, i4=0, i5=0, i6=0):
    return (sp - 1) + 3 * (i1 + 4 * (i2 + 4 * (i3 + 4 * (i4 + 4 * (i5 + 4 * i6)))))


def parse_start_tables(path):
    """Rate-table tail of a start.dat (mcevent/SetProbability.f:21-143) + species header (main.f:83-100)."""
    lines = [l for l in open(path).read().split("\n")]
    ncry, ncov = (int(v) for v in lines[0].split()[:2])
    listcrystal = [int(v) for v in lines[1].split()[:ncry]]
    listcov = [int(v) for v in lines[2].split()[:ncov]] if ncov else []
    nE = {(1, 0): 3, (2, 0): 18, (3, 0): 57, (1, 1): 9, (2, 1): 50, (3, 1): 150, (1, 2): 19, (2, 2): 100, (3, 2): 288, (1, 3): 34, (2, 3): 172, (3, 3): 480}[(ncry, ncov)]
    ntail = nE + 3 * ncry + 6 * ncov
    body = [l for l in lines if l.strip()]
    tail = body[-ntail:]
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
    # PtransZig is the line before sf, i.e. two lines before the ExitStrategy keyword line
    zi = next(i for i, l in enumerate(lines) if l.split() and l.split()[0] in ("Iter", "Time")) - 2
    return dict(listcrystal=listcrystal, listcov=listcov, PtransE=PE, PtransD=PD, PtransAbs=PA, PtransDes=PS,
                PtransZig=float(lines[zi].split()[0]), sf=float(lines[zi + 1].split()[0]))


def rates_for(case):
    w = parse_start_tables(os.path.join(_GOLDEN, f"start-{case}.dat"))
    w["walls"] = case in ("Si-SL", "CVD-Si-SiH2")          # FS builder has walls, F has none
    w["two_species"] = len(w["listcrystal"]) == 2
    return w
