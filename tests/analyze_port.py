"""The reference's frame readers (pymulskips/analyze.py), restated for the GPU tests -- the GPU box has no /root/reference.
TEST INFRASTRUCTURE.  Each function cites what it follows; tests/test_analyze_port.py (CPU, this container) pins them on
frames written by the oracle against the real pymulskips.analyze."""
from __future__ import annotations

import glob
import os

import numpy as np


def read_xyz(path):
    """(count, box [A], species list, z [A] array, KMC time) of one frame (analyze.py:129-155,275-281)"""
    with open(path) as f:
        content = f.read().splitlines()
    head = content[0].split()
    nat, time = int(head[0]), float(head[3])
    box = [float(v) for v in content[1].split()[1:4]]
    sp, z = [], []
    for line in content[2:]:
        t = line.split()
        if len(t) >= 4:
            sp.append(t[0]); z.append(float(t[3]))
    assert len(z) == nat, (path, len(z), nat)
    return nat, box, sp, np.array(z), time


def _seqsum(vals):
    """left-to-right float accumulation, as the reference's loops do (Python >= 3.12's sum() compensates)"""
    acc = 0.0
    for v in vals:
        acc = acc + v
    return acc


def get_surface_height(path, bin_size=5.0, surface_roughness=20.0):
    """analyze.py:124-207: histogram of z in bins of bin_size, mean z inside the fullest bin, then the mean z of every atom
    within surface_roughness of that"""
    nat, box, sp, z, _ = read_xyz(path)
    zl = box[2]
    if np.all(z == z[0]) or (z.max() - z.min()) < bin_size:
        return float(sum(z.tolist()) / len(z))   # (the reference uses sum() here)
    databins = int(zl / bin_size)
    hist = np.histogram(z, bins=databins, range=(0.0, zl))
    imax = [i for i, j in enumerate(hist[0]) if j == max(hist[0])]
    z_min = float(bin_size * imax[0]); z_max = float(z_min + bin_size)
    if z_max < z.min():
        z_min += bin_size; z_max += bin_size
    sel = z[(z > z_min) & (z < z_max)]
    z_ave = float(_seqsum(sel.tolist()) / float(len(sel)))
    sel = z[(z > z_ave - surface_roughness) & (z < z_ave + surface_roughness)]
    return float(_seqsum(sel.tolist()) / float(len(sel)))


def frames(rundir, what="undercoordinated"):
    """analyze.py:10-30"""
    pat = {"undercoordinated": "*[0-9].xyz", "defects": "*[0-9]_d.xyz", "wrong": "*[0-9]_w.xyz", "vacancies": "*[0-9]_v.xyz"}[what]
    return sorted(glob.glob(os.path.join(rundir, pat)))


def growth_rate(rundir, bin_size=5.0, surface_roughness=20.0, Nexclude=2):
    """analyze.py:209-330, method 'finitediff': surface height of every frame against the KMC time of its header, centred
    finite differences (one-sided at the ends), mean over the frames Nexclude in from both ends; micron/hour"""
    files = frames(rundir)
    h = np.array([get_surface_height(f, bin_size, surface_roughness) for f in files])
    t = np.array([read_xyz(f)[4] for f in files])
    gr = [(h[1] - h[0]) / (t[1] - t[0]) * 3600 * 1e-4]
    for i in range(1, len(t) - 1):
        gr.append((h[i + 1] - h[i - 1]) / (t[i + 1] - t[i - 1]) * 3600 * 1e-4)
    gr.append((h[-1] - h[-2]) / (t[-1] - t[-2]) * 3600 * 1e-4)
    return float(np.mean(np.array(gr[Nexclude:-Nexclude]))), h, t


def get_coverage(path):
    """analyze.py:418-481 (Miller 100: one dangling bond per surface atom): H and Cl per surface atom, where the surface atoms
    are the under-coordinated crystal atoms minus four per vacancy ('O' lines)"""
    nat, box, sp, z, _ = read_xyz(path)
    nH, nCl, nvac = sp.count("H"), sp.count("Cl"), sp.count("O")
    nunder = len(sp) - nH - nCl - nvac
    surface_atoms = nunder - 4 * nvac
    return {"H": nH / surface_atoms, "Cl": nCl / surface_atoms, "surface_atoms": surface_atoms, "vacancies": nvac}


def count_lines(path):
    """analyze.py:584-603 (`tail -n +3 | wc -l`): entries of a _v.xyz / _d.xyz frame"""
    with open(path) as f:
        return max(len(f.read().splitlines()) - 2, 0)
