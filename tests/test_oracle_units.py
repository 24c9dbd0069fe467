"""Unit-level checks of the oracle's building blocks against known answers / properties."""
import ctypes as C

import numpy as np

from cases import make_start
from oracle_binding import Oracle, load_oracle


def test_numberslist_table():
    lib = load_oracle()
    # PickTreeEvent.f:77 first and :1076 last literal, REAL(4) widened
    assert lib.orc_numberslist(1) == float(np.float32(0.22188498))
    assert lib.orc_numberslist(1000) == float(np.float32(0.73450478))
    assert lib.orc_numberslist(1001) == lib.orc_numberslist(1)


def test_ran3_stream(tmp_path):
    """ran3 with IDUM=9117116: in (0,1), multiples of the REAL(4) 1e-9, reproducible, Knuth recurrence."""
    s = make_start("Si-SL", str(tmp_path / "s.dat"))
    a, b = Oracle(48, 48, 48, s), Oracle(48, 48, 48, s)
    xs = [a.lib.orc_ran3(a.o) for _ in range(200)]
    ys = [b.lib.orc_ran3(b.o) for _ in range(200)]
    assert xs == ys and all(0 <= x < 1 for x in xs)
    fac = float(np.float32(1.0) / np.float32(1e9))
    ints = [round(x / fac) for x in xs]
    assert all(abs(i * fac - x) == 0 for i, x in zip(ints, xs))
    for k in range(55, 200):          # subtractive lagged Fibonacci: x[k] = x[k-55] - x[k-24] mod 1e9
        assert ints[k] == (ints[k - 55] - ints[k - 24]) % 1000000000


def test_find_nn_tetrahedron():
    lib = load_oracle()
    c = np.array([12, 12, 12], np.int32)
    jd = np.array([[3, 3, 3], [3, -3, -3], [-3, -3, 3], [-3, 3, -3]], np.int32)
    out = np.zeros((2, 3), np.int32)
    s1, s2 = (c + jd[0]).astype(np.int32), (c + jd[1]).astype(np.int32)
    lib.orc_find_nn(s1.ctypes.data, s2.ctypes.data, c.ctypes.data, out.ctypes.data, 480, 480)
    got = {tuple(r - c) for r in out}
    assert got == {tuple(jd[2]), tuple(jd[3])}


def test_find_sn_zigzag_is_cubic():
    lib = load_oracle()
    site = np.array([15, 15, 15], np.int32); sc = np.array([12, 12, 12], np.int32)
    sn = (sc + np.array([[3, -3, -3], [-3, -3, 3], [-3, 3, -3]], np.int32)).astype(np.int32)
    zig, arm = np.zeros((3, 3), np.int32), np.zeros((3, 3), np.int32)
    lib.orc_find_sn(site.ctypes.data, sc.ctypes.data, sn.ctypes.data, zig.ctypes.data, arm.ctypes.data, 480, 480)
    # zig-zag = inversion of the second neighbours through the bond centre -> the diamond (cubic) continuation
    assert {tuple(r - site) for r in zig} == {(-3, 3, 3), (3, 3, -3), (3, -3, 3)}
    # arm-chair = second neighbours translated along the bond by 5*sqrt(3) -> eclipsed (hexagonal) continuation
    assert {tuple(r - site) for r in arm} == {(5, -1, -1), (-1, -1, 5), (-1, 5, -1)}


def test_tree_invariant_and_time(tmp_path):
    s = make_start("surface", str(tmp_path / "s.dat"), lens=(48, 0, 0), maxiter=10**9, outmolmol=10**9)
    o = Oracle(96, 96, 240, s)
    o.setup()
    o.set_trace(5000)
    o.run(5000)
    tr = o.trace()
    assert len(tr) == 5000
    t = 0.0
    for it, ind, site, itrans, tree1, time in tr:
        t = t + 1.0 / tree1
        assert t == time                       # Time += 1/Tree(1), main.f:313
    d = o.lists()
    assert abs(d["mc_rate"].sum() - o.tree_root) <= 1e-9 * o.tree_root
    q = o.quirks()
    assert q["oob"] == 0 and q["stale_nbuff"] == 0
