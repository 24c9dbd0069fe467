"""Test-case helpers: reference start.dat files (verbatim copies under tests/golden/) and reduced variants."""
from __future__ import annotations

import os

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
CASES = ["Si-SL", "surface", "cube", "cube-zigzag", "sphere", "inverted-pyramid", "finfet", "CVD-Si-SiH2", "CAD", "LA"]


def start_path(case):
    return os.path.join(GOLDEN, f"start-{case}.dat")


def make_start(case, dst, lens=None, maxiter=None, outmolmol=None, runtype=None, idum=None, ptranszig=None, maxtime=None, outtime=None, paths=None):
    """Copy tests/golden/start-<case>.dat to dst with some of the positional lines replaced.
    The grammar is the one of mulskips-source/main.f:83-212."""
    lines = open(start_path(case)).read().split("\n")
    ncov = int(lines[0].split()[1])
    i = 2 + (1 if ncov > 0 else 0)          # InitSt line
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
