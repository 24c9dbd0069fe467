"""x-decomposed slab mode on ONE GPU: with nranks == 1 the rank is its own left and right neighbour, so the
halo machinery (dirty-cell log, pack, apply) runs exactly as between two GPUs, minus the NCCL transport.
Checked: (1) after any number of sweeps both copies of each shared tile column are identical -- lattice
words and bonds -- i.e. every change made on one side of the seam reached the other side; (2) the growth
statistics of the slab run equal those of the plain periodic box of the same global size."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GX, GY, GZ, LEN1 = 192, 96, 144, 48


def make(slab, seed=None):
    from mulskips_b200 import Mskc
    from mulskips_b200.workloads import rates_for
    w = rates_for("surface"); w["PtransZig"] = 0.8
    ctx = Mskc(GX, GY, GZ, w["listcrystal"], w["listcov"], slab=(0, 1) if slab else None)
    ctx.set_rates(w["PtransE"], w["PtransD"], w["PtransAbs"], w["PtransDes"], w["PtransZig"])
    ctx.init_flat(LEN1, walls=False, two_species=True)
    return ctx


def test_seam_copies_identical():
    ctx = make(True)
    assert ctx.LenX == GX + 48
    tot = 0
    for i in range(6):
        r = ctx.run_parallel(40, events_per_tile=4.0, seed=10 + i)
        tot += r.events_done
        L = ctx.download_lattice()
        own = GX
        assert np.array_equal(L[:, :, 0:24], L[:, :, own:own + 24]), f"left halo != last owned column after {tot} events"
        assert np.array_equal(L[:, :, own + 24:own + 48], L[:, :, 24:48]), f"right halo != first owned column after {tot} events"
        # bonds of every non-empty cell of the shared columns
        # (x in 9..38 of the strip: bond targets stay inside 0..47, so absolute targets compare without wrap artefacts)
        z, y, x = np.nonzero(L[:, :, 9:39])
        a = np.stack([x + 9, y, z], 1).astype(np.int32)
        b = a.copy(); b[:, 0] += own
        na, nb = ctx.download_bonds(a), ctx.download_bonds(b)
        unset_a, unset_b = (na == 0).all(axis=2), (nb == 0).all(axis=2)
        assert np.array_equal(unset_a, unset_b)
        da = np.where(unset_a[..., None], 0, na - a[:, None, :]); db = np.where(unset_b[..., None], 0, nb - b[:, None, :])
        assert np.array_equal(da, db), "bond vectors differ across the seam"
    assert tot > 50000
    ctx.close()


def growth(slab, seed, sweeps=400):
    ctx = make(slab)
    r = None
    t = 0.0
    while t < 6.0e5:
        r = ctx.run_parallel(20, events_per_tile=4.0, seed=seed)
        t = r.time
    c, _ = ctx.census()
    ctx.close()
    return c.occupied, c.undercoordinated, r.time


def test_slab_statistics_match_periodic_box():
    a = np.array([growth(True, 100 + s)[:2] for s in range(8)], float)
    b = np.array([growth(False, 200 + s)[:2] for s in range(8)], float)
    base = 2 * (GX // 12) * (GY // 12) * (LEN1 // 3 + 1)
    ga, gb = a[:, 0] - base, b[:, 0] - base
    print(f"\ngrowth slab {ga.mean():.1f} +- {ga.std(ddof=1):.1f} | periodic {gb.mean():.1f} +- {gb.std(ddof=1):.1f}; surface sites {a[:, 1].mean():.1f} | {b[:, 1].mean():.1f}")
    se = np.hypot(ga.std(ddof=1), gb.std(ddof=1)) / np.sqrt(8)
    assert gb.mean() > 1000
    assert abs(ga.mean() - gb.mean()) <= 0.02 * gb.mean() + 3 * se
    assert abs(a[:, 1].mean() - b[:, 1].mean()) <= 3 * np.hypot(a[:, 1].std(ddof=1), b[:, 1].std(ddof=1))


def test_slab_long_growth():
    """Twenty atomic layers in slab mode (round-1 advisory: the mirrored halo lists only ever received entries -- their
    tombstones are now dropped once per sweep -- so a long growth run must neither fill a halo list nor let the two copies
    of a shared column drift apart)."""
    ctx = make(True)
    base = 2 * (GX // 12) * (GY // 12) * (LEN1 // 3 + 1)
    layer = 2 * (GX // 12) * (GY // 12)
    grown, calls = 0, 0
    while grown < 20 * layer:
        ctx.run_parallel(200, events_per_tile=4.0, seed=500 + calls)
        calls += 1
        c, _ = ctx.census()
        grown = c.occupied - base
        assert calls < 400, f"only {grown} atoms after {calls} calls"
    L = ctx.download_lattice()
    assert np.array_equal(L[:, :, 0:24], L[:, :, GX:GX + 24]) and np.array_equal(L[:, :, GX + 24:GX + 48], L[:, :, 24:48])
    print(f"\n{grown} atoms ({grown / layer:.1f} layers) in {calls} calls of 200 sweeps")
    ctx.close()
