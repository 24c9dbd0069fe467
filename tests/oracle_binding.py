"""ctypes binding of the CPU oracle (oracle/libmskc_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None


class OrcTrace(C.Structure):
    _fields_ = [("iter", C.c_int64), ("ind", C.c_int32), ("site", C.c_int32 * 3), ("itrans", C.c_int32), ("pad", C.c_int32),
                ("tree1", C.c_double), ("time", C.c_double)]


def build_oracle(native=False):
    """native=True: additionally build oracle/libmskc_oracle_native.so (-march=native) ON THIS machine and make it the library
    load_oracle() picks (bench.py's CPU arm; the default library is built for a portable target because it travels)"""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if native:
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "native"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        os.environ["MSKC_ORACLE_LIB"] = os.path.join(ORACLE_DIR, "libmskc_oracle_native.so")


def load_oracle():
    global _LIB
    if _LIB is not None:
        return _LIB
    p = os.environ.get("MSKC_ORACLE_LIB") or os.path.join(ORACLE_DIR, "libmskc_oracle.so")
    if not os.path.exists(p):
        build_oracle()
    lib = C.CDLL(p)
    vp = C.c_void_p
    lib.orc_create.restype = vp
    lib.orc_create.argtypes = [C.c_int] * 4
    lib.orc_destroy.argtypes = [vp]
    lib.orc_read_start.argtypes = [vp, C.c_char_p]
    lib.orc_setup.argtypes = [vp]
    lib.orc_iterate.argtypes = [vp]
    lib.orc_run.argtypes = [vp, C.c_int64]
    lib.orc_run.restype = C.c_int64
    lib.orc_finish.argtypes = [vp]
    lib.orc_write_frame.argtypes = [vp, C.c_double, C.c_int64, C.c_int64]
    lib.orc_state_hash.argtypes = [vp]
    lib.orc_state_hash.restype = C.c_uint64
    lib.orc_lattcoo.argtypes = [vp]
    lib.orc_lattcoo.restype = C.POINTER(C.c_int32)
    lib.orc_get_sizes.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.orc_get_time.argtypes = [vp]
    lib.orc_get_time.restype = C.c_double
    lib.orc_get_tree_root.argtypes = [vp]
    lib.orc_get_tree_root.restype = C.c_double
    lib.orc_get_last_site.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.orc_export_lists.argtypes = [vp] + [C.c_void_p] * 8
    lib.orc_get_tables.argtypes = [vp] + [C.c_void_p] * 5
    lib.orc_get_species.argtypes = [vp, C.c_void_p]
    lib.orc_get_field.argtypes = [vp] + [C.c_void_p] * 6
    lib.orc_get_fork.argtypes = [vp, C.c_void_p, C.c_void_p]
    lib.orc_set_variant.argtypes = [vp, C.c_int]
    lib.orc_set_numocc0.argtypes = [vp, C.c_int]
    lib.orc_next_leg.argtypes = [vp, C.c_void_p]
    lib.orc_get_run.argtypes = [vp, C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p]
    lib.orc_get_rng.argtypes = [vp, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    lib.orc_get_counts.argtypes = [vp, C.c_void_p, C.c_void_p]
    lib.orc_get_quirks.argtypes = [vp, C.c_void_p]
    lib.orc_set_io.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int]
    lib.orc_set_run.argtypes = [vp, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int]
    lib.orc_set_trace.argtypes = [vp, C.c_int64]
    lib.orc_get_trace.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.orc_get_trace.restype = C.POINTER(OrcTrace)
    lib.orc_errmsg.argtypes = [vp]
    lib.orc_errmsg.restype = C.c_char_p
    lib.orc_ran3.argtypes = [vp]
    lib.orc_ran3.restype = C.c_double
    lib.orc_numberslist.argtypes = [C.c_int64]
    lib.orc_numberslist.restype = C.c_double
    lib.orc_find_nn.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int]
    lib.orc_find_sn.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_int]
    _LIB = lib
    return lib


class Oracle:
    def __init__(self, lenx, leny, lenz, start_path, nummcmax=0, outdir=None, logpath=None, write_files=False, la_tree=False):
        self.lib = load_oracle()
        self.o = self.lib.orc_create(lenx, leny, lenz, nummcmax)
        if not self.o:
            raise MemoryError("orc_create")
        self.la_tree = bool(la_tree)
        self.lib.orc_set_variant(self.o, int(self.la_tree))     # True: follow mulskips-source-LaserAnnealing instead of mulskips-source
        self.LenX, self.LenY, self.LenZ = lenx, leny, lenz
        self.lib.orc_set_io(self.o, outdir.encode() if outdir else None, logpath.encode() if logpath else None, int(write_files))
        if self.lib.orc_read_start(self.o, start_path.encode()) != 0:
            raise RuntimeError("oracle start.dat: " + self.lib.orc_errmsg(self.o).decode())

    def close(self):
        if self.o:
            self.lib.orc_set_io(self.o, None, None, 0)
            self.lib.orc_destroy(self.o)
            self.o = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_run(self, maxiter=0, outmolmol=0, runtype_T=-1, idum=0, faithful_cost=0):
        self.lib.orc_set_run(self.o, maxiter, outmolmol, runtype_T, idum, faithful_cost)

    def setup(self):
        if self.lib.orc_setup(self.o) != 0:
            raise RuntimeError("oracle setup: " + self.lib.orc_errmsg(self.o).decode())

    def run(self, n):
        r = self.lib.orc_run(self.o, n)
        if r < 0:
            raise RuntimeError("oracle run: " + self.lib.orc_errmsg(self.o).decode())
        return r

    def finish(self):
        self.lib.orc_finish(self.o)

    def write_frame(self, iout):
        """one frame through the oracle's WriteMolMolSource / WriteMolMolXYZFile (into outdir when write_files is set)"""
        s = self.sizes()
        if self.lib.orc_write_frame(self.o, C.c_double(self.time), C.c_int64(s["Iter"]), C.c_int64(iout)) != 0:
            raise RuntimeError("oracle frame: " + self.lib.orc_errmsg(self.o).decode())

    def sizes(self):
        a = (C.c_int64 * 8)()
        self.lib.orc_get_sizes(self.o, a)
        return dict(NumAdAtom=a[0], NumAtoms=a[1], NumVoids=a[2], Iter=a[3], IterOut=a[4], ncell=a[5], Levels=a[6], done=a[7])

    @property
    def time(self):
        return self.lib.orc_get_time(self.o)

    @property
    def tree_root(self):
        return self.lib.orc_get_tree_root(self.o)

    def last_site(self):
        a = (C.c_int32 * 3)()
        self.lib.orc_get_last_site(self.o, a)
        return tuple(a)

    def hash(self):
        return self.lib.orc_state_hash(self.o)

    def lattcoo(self):
        n = (self.LenZ + 1) * self.LenY * self.LenX
        return np.ctypeslib.as_array(self.lib.orc_lattcoo(self.o), shape=(n,)).reshape(self.LenZ + 1, self.LenY, self.LenX)

    def lists(self):
        s = self.sizes()
        nm, nb, nv = s["NumAdAtom"], s["NumAtoms"], s["NumVoids"]
        d = dict(mc_xyz=np.zeros((nm, 3), np.int32), mc_nn=np.zeros((nm, 4, 3), np.int32), mc_event=np.zeros(nm, np.int32),
                 mc_rate=np.zeros(nm, np.float64), bulk_xyz=np.zeros((nb, 3), np.int32), bulk_nn=np.zeros((nb, 4, 3), np.int32),
                 void_xyz=np.zeros((nv, 3), np.int32), void_nn=np.zeros((nv, 4, 3), np.int32))
        self.lib.orc_export_lists(self.o, *[d[k].ctypes.data for k in ("mc_xyz", "mc_nn", "mc_event", "mc_rate", "bulk_xyz", "bulk_nn", "void_xyz", "void_nn")])
        return d

    def tables(self):
        PE, PD, PA, PS, sc = np.zeros(12288), np.zeros(9), np.zeros(12), np.zeros(12), np.zeros(8)
        self.lib.orc_get_tables(self.o, PE.ctypes.data, PD.ctypes.data, PA.ctypes.data, PS.ctypes.data, sc.ctypes.data)
        return dict(PtransE=PE, PtransD=PD, PtransAbs=PA, PtransDes=PS, PtransZig=sc[0], sf=sc[1], Tm=sc[2], P0=sc[3], sigma=sc[4], Tflex=sc[5],
                    MaxTime=sc[6], OutTime=sc[7])

    def field(self):
        """'ST' mode: the tabulated rate blocks of the probability-table file (SetProbabilityFromField.f)."""
        info = np.zeros(4, np.int32)
        self.lib.orc_get_field(self.o, info.ctypes.data, None, None, None, None, None)
        n = int(info[0])
        rt, E, D, A, S = np.zeros(n, np.int32), np.zeros((n, 12288)), np.zeros((n, 9)), np.zeros((n, 12)), np.zeros((n, 12))
        self.lib.orc_get_field(self.o, info.ctypes.data, rt.ctypes.data, E.ctypes.data, D.ctypes.data, A.ctypes.data, S.ctypes.data)
        return dict(ntemp=n, fixedT=int(info[1]), Ttop=int(info[2]), reftemp=rt, refPE=E, refPD=D, refPAbs=A, refPDes=S)

    def next_leg(self, T):
        """second and later legs of a socket-driven laser-annealing run (la/main.f:616-631): T is [LenZ][LenY*LenX] int32"""
        T = np.ascontiguousarray(T, dtype=np.int32)
        assert T.size == self.LenZ * self.LenY * self.LenX
        self.lib.orc_next_leg(self.o, T.ctypes.data)

    def set_numocc0(self, v):
        self.lib.orc_set_numocc0(self.o, int(v))

    def fork(self):
        """LaserAnnealing-fork bookkeeping (la/main.f:592-611, la/modules/defsystem.f:36-37,46,82)."""
        iv, dv = np.zeros(10, np.int32), np.zeros(12)
        self.lib.orc_get_fork(self.o, iv.ctypes.data, dv.ctypes.data)
        return dict(la_tree=int(iv[0]), NumAdAtomOcc=int(iv[1]), NumOcc0=int(iv[2]), NumOcc1=int(iv[3]), InitCountCrystal=iv[4:7].copy(),
                    CountLiquid=iv[7:10].copy(), Tm=dv[0:3].copy(), P0=dv[3:6].copy(), sigma=dv[6:9].copy(), Tflex=dv[9:12].copy())

    def species(self):
        a = np.zeros(8, np.int32)
        self.lib.orc_get_species(self.o, a.ctypes.data)
        nc, nv = int(a[0]), int(a[1])
        return [int(x) for x in a[2:2 + nc]], [int(x) for x in a[5:5 + nv]]

    def run_params(self):
        a = np.zeros(4, np.int64)
        i, r, e = C.create_string_buffer(4), C.create_string_buffer(4), C.create_string_buffer(8)
        self.lib.orc_get_run(self.o, a.ctypes.data, i, r, e)
        return dict(MaxIter=int(a[0]), OutMolMol=int(a[1]), Idum=int(a[2]), counter=int(a[3]), InitSt=i.value.decode(), RunType=r.value.decode(),
                    ExitStrategy=e.value.decode())

    def rng(self):
        st = np.zeros(58, np.int32); c = C.c_int64(); i = C.c_int32()
        self.lib.orc_get_rng(self.o, st.ctypes.data, C.byref(c), C.byref(i))
        return st, c.value, i.value

    def counts(self):
        cc, cv = np.zeros(3, np.int32), np.zeros(3, np.int32)
        self.lib.orc_get_counts(self.o, cc.ctypes.data, cv.ctypes.data)
        return cc, cv

    def quirks(self):
        q = np.zeros(8, np.int64)
        self.lib.orc_get_quirks(self.o, q.ctypes.data)
        return dict(stale_nbuff=int(q[0]), wl_iso=int(q[1]), wl_cov=int(q[2]), wl_dep=int(q[3]), wl_des=int(q[4]), oob=int(q[5]), anom=int(q[6]), n_getprob=int(q[7]))

    def set_trace(self, cap):
        self.lib.orc_set_trace(self.o, cap)

    def trace(self):
        n = C.c_int64()
        p = self.lib.orc_get_trace(self.o, C.byref(n))
        return [(p[i].iter, p[i].ind, tuple(p[i].site), p[i].itrans, p[i].tree1, p[i].time) for i in range(n.value)]


def trace_arrays(orc: Oracle):
    """The oracle's event trace as numpy arrays: site [n,3], event code [n], arm decision [n] (0 none, 1 zig-zag, 2 arm-chair)."""
    n = C.c_int64()
    p = orc.lib.orc_get_trace(orc.o, C.byref(n))
    dt = np.dtype([("iter", "<i8"), ("ind", "<i4"), ("site", "<i4", 3), ("itrans", "<i4"), ("arm", "<i4"), ("tree1", "<f8"), ("time", "<f8")])
    assert dt.itemsize == C.sizeof(OrcTrace)
    a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n.value * dt.itemsize,)).view(dt).copy() if n.value else np.zeros(0, dt)
    return a["site"].copy(), a["itrans"].copy(), a["arm"].copy(), a["tree1"].copy()


def upload_oracle_state(orc: Oracle, ctx, trange=None):
    """Hand the oracle's current state to a CUDA context: the tests' way of giving both sides the same input.
    trange = (tmin, tmax): temperature range of the per-kelvin LA tables (default: the range found in the lattice)."""
    t = orc.tables()
    L = orc.lattcoo()
    rp = orc.run_params()
    if rp["InitSt"] == "ST":
        fl = orc.field()
        T = (L[:-1] >> 9) & 8191
        tv = T[T > 0]
        ctx.set_rate_field(fl["reftemp"], fl["refPE"], fl["refPD"], fl["refPAbs"], fl["refPDes"], t["PtransZig"], int(tv.min()), int(tv.max()),
                           fixed_temperature=fl["Ttop"] if fl["fixedT"] else 0)
    else:
        ctx.set_rates(t["PtransE"], t["PtransD"], t["PtransAbs"], t["PtransDes"], t["PtransZig"])
    if orc.la_tree:
        ctx.set_tree_variant(1)
    if rp["InitSt"] == "LA" and orc.la_tree:
        fk = orc.fork()
        T = (L[:-1] >> 9) & 8191
        tv = T[T > 0]
        ctx.set_la_fork(fk["Tm"], fk["P0"], fk["sigma"], fk["Tflex"], int(tv.min()), int(tv.max()), fk["InitCountCrystal"], fk["NumOcc0"], fk["NumAdAtomOcc"])
    elif rp["InitSt"] == "LA":
        T = (L[:-1] >> 9) & 8191
        tv = T[T > 0]
        lo, hi = trange if trange else (int(tv.min()), int(tv.max()))
        ctx.set_la(t["Tm"], t["P0"], t["sigma"], t["Tflex"], lo, hi)
    d = orc.lists()
    st, counter, idum = orc.rng()
    ctx.seed(idum, rp["RunType"][0], counter, st)
    ctx.upload_state(L, d["mc_xyz"], d["mc_nn"], d["mc_event"], d["mc_rate"], d["bulk_xyz"], d["bulk_nn"], d["void_xyz"], d["void_nn"])
    cc, cv = orc.counts()
    s = orc.sizes()
    ctx.set_clock(s["Iter"], orc.time, cc, cv)
