"""ctypes mirror of include/mskc.h (the drop-in boundary of the KMC event loop).

Names and argument meaning follow the reference's module globals: LattCoo words
(mulskips-source/modules/defsystem.f:49,87-91), ListAdAtom/ListAtom/ListVoid records
(modules/defdertype.f:26-48), PtransE/D/Abs/Des tables (defsystem.f:60-65).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path() -> str:
    # MSKC_LIB: an A/B build of the same library (scripts/probe_*.py); the default is the in-tree build
    return os.environ.get("MSKC_LIB") or os.path.join(_HERE, "lib", "libmskc.so")


class MskcError(RuntimeError):
    pass


class _ParityArgs(C.Structure):
    _fields_ = [("max_iter", C.c_int64), ("max_time", C.c_double), ("exit_on_time", C.c_int32), ("pad", C.c_int32),
                ("out_every", C.c_int64), ("out_time", C.c_double), ("framecounter", C.c_int64), ("event_budget", C.c_int64)]


class ParityResult(C.Structure):
    _fields_ = [("iter", C.c_int64), ("iter_out", C.c_int64), ("time", C.c_double), ("last_site", C.c_int32 * 3),
                ("stop_reason", C.c_int32), ("num_atoms", C.c_int32), ("num_voids", C.c_int32), ("num_adatom", C.c_int32),
                ("pad", C.c_int32), ("events_done", C.c_int64), ("site_updates", C.c_int64), ("tree_root", C.c_double),
                ("kernel_ms", C.c_float), ("launches", C.c_int32)]


class _ParallelArgs(C.Structure):
    _fields_ = [("n_sweeps", C.c_int64), ("events_per_tile", C.c_double), ("tau_fixed", C.c_double), ("seed", C.c_uint64),
                ("tile", C.c_int32), ("flags", C.c_int32)]


class ParallelResult(C.Structure):
    _fields_ = [("events_done", C.c_int64), ("site_updates", C.c_int64), ("sweeps_done", C.c_int64), ("time", C.c_double),
                ("total_rate", C.c_double), ("n_active", C.c_int64), ("count_dep", C.c_int64 * 3), ("count_evap", C.c_int64),
                ("count_abs", C.c_int64), ("count_des", C.c_int64), ("kernel_ms", C.c_float), ("launches", C.c_int32),
                ("sweep_kernel_ms", C.c_float), ("stop_reason", C.c_int32), ("sampler_launches", C.c_int32), ("max_tile_events", C.c_int32), ("tau", C.c_double)]


class Counters(C.Structure):
    _fields_ = [("num_atoms", C.c_int32), ("num_voids", C.c_int32), ("num_adatom", C.c_int32), ("num_adatom_occ", C.c_int32),
                ("CountCrystal", C.c_int32 * 3), ("CountCov", C.c_int32 * 3), ("iter", C.c_int64), ("time", C.c_double)]


class Census(C.Structure):
    _fields_ = [("undercoordinated", C.c_int64), ("coverage", C.c_int64), ("noepi_bulk", C.c_int64), ("bulk", C.c_int64),
                ("occupied", C.c_int64), ("walls", C.c_int64), ("mean_height", C.c_double), ("sq_height", C.c_double)]


STOP_BUDGET, STOP_MAXITER, STOP_MAXTIME, STOP_FRAME, STOP_FEWMC, STOP_RESOLIDIFIED, STOP_ERROR = 0, 1, 2, 3, 4, 5, -1
TREE_SRC, TREE_LA = 0, 1

_LIB = None

# every symbol include/mskc.h declares
SYMBOLS = ["mskc_create", "mskc_destroy", "mskc_last_error", "mskc_set_rates", "mskc_set_la", "mskc_set_rate_field", "mskc_set_tree_variant", "mskc_set_la_fork", "mskc_upload_state",
           "mskc_init_flat", "mskc_seed", "mskc_get_rng", "mskc_set_clock", "mskc_run_parity", "mskc_run_parallel",
           "mskc_rerate_all", "mskc_set_temperature", "mskc_download_lattice", "mskc_patch_lattice", "mskc_get_counters",
           "mskc_download_lists", "mskc_state_hash", "mskc_census_lattice", "mskc_debug_tile_stats",
           "mskc_replay_events", "mskc_download_active", "mskc_download_bonds",
           "mskc_create_slab", "mskc_comm_unique_id", "mskc_comm_connect", "mskc_extract_frame", "mskc_extract_frame_all", "mskc_extract_antisites", "mskc_gather_lattice", "mskc_check_flags", "mskc_replay_teams", "mskc_set_temperature_u16", "mskc_download_phases", "mskc_last_dense_ms"]


def load_library():
    """Load lib/libmskc.so; fails loudly if the CUDA extension has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise MskcError(f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback for the KMC event loop)")
    lib = C.CDLL(p)
    vp, i32p, f64p, i64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int64)
    lib.mskc_create.argtypes = [C.c_int] * 5 + [i32p, i32p, C.c_int, C.c_int, C.POINTER(vp)]
    lib.mskc_destroy.argtypes = [vp]
    lib.mskc_last_error.argtypes = [vp]
    lib.mskc_last_error.restype = C.c_char_p
    lib.mskc_set_rates.argtypes = [vp, f64p, f64p, f64p, f64p, C.c_double]
    lib.mskc_set_la.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
    lib.mskc_set_rate_field.argtypes = [vp, C.c_int, i32p, f64p, f64p, f64p, f64p, C.c_double, C.c_int, C.c_int, C.c_int]
    lib.mskc_set_tree_variant.argtypes = [vp, C.c_int]
    lib.mskc_set_la_fork.argtypes = [vp, C.c_int, f64p, f64p, f64p, f64p, C.c_int, C.c_int, i32p, C.c_int, C.c_int]
    lib.mskc_upload_state.argtypes = [vp, i32p, C.c_int, i32p, i32p, i32p, f64p, C.c_int, i32p, i32p, C.c_int, i32p, i32p]
    lib.mskc_init_flat.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.mskc_seed.argtypes = [vp, C.c_int, C.c_int, C.c_int64, i32p]
    lib.mskc_get_rng.argtypes = [vp, i32p, i64p, i32p]
    lib.mskc_set_clock.argtypes = [vp, C.c_int64, C.c_double, i32p, i32p]
    lib.mskc_run_parity.argtypes = [vp, C.POINTER(_ParityArgs), C.POINTER(ParityResult)]
    lib.mskc_run_parallel.argtypes = [vp, C.POINTER(_ParallelArgs), C.POINTER(ParallelResult)]
    lib.mskc_rerate_all.argtypes = [vp, i64p, f64p, C.POINTER(C.c_float)]
    lib.mskc_set_temperature.argtypes = [vp, i32p]
    lib.mskc_download_lattice.argtypes = [vp, i32p]
    lib.mskc_patch_lattice.argtypes = [vp, C.c_int, i64p, i32p]
    lib.mskc_get_counters.argtypes = [vp, C.POINTER(Counters)]
    lib.mskc_download_lists.argtypes = [vp, i32p, i32p, i32p, f64p, i32p, i32p, i32p, i32p, C.c_int64, i64p]
    lib.mskc_state_hash.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.mskc_census_lattice.argtypes = [vp, C.POINTER(Census), C.POINTER(C.c_float)]
    lib.mskc_debug_tile_stats.argtypes = [vp, C.c_void_p, C.c_int64, i64p]
    lib.mskc_replay_events.argtypes = [vp, C.c_int64, i32p, i32p, i32p, C.c_double]
    lib.mskc_download_active.argtypes = [vp, C.c_int64, i32p, i32p, f64p, i64p]
    lib.mskc_download_bonds.argtypes = [vp, C.c_int64, i32p, i32p]
    lib.mskc_create_slab.argtypes = [C.c_int] * 5 + [i32p, i32p, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    lib.mskc_comm_unique_id.argtypes = [C.c_void_p]
    lib.mskc_comm_connect.argtypes = [vp, C.c_void_p]
    lib.mskc_extract_frame.argtypes = [vp, C.c_int64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), i64p, C.POINTER(C.c_float)]
    lib.mskc_extract_frame_all.argtypes = [vp, C.c_int64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), i64p, C.POINTER(C.c_float)]
    lib.mskc_extract_antisites.argtypes = [vp, C.c_int, C.c_int64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), i64p, C.POINTER(C.c_float)]
    lib.mskc_gather_lattice.argtypes = [vp, C.c_int64, i64p, C.POINTER(C.c_int32)]
    lib.mskc_set_temperature_u16.argtypes = [vp, C.POINTER(C.c_uint16)]
    lib.mskc_download_phases.argtypes = [vp, C.POINTER(C.c_uint8)]
    lib.mskc_last_dense_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.mskc_replay_teams.argtypes = [vp, C.c_int64, i32p, i32p, i32p, C.c_double, i64p, i64p]
    lib.mskc_check_flags.argtypes = [vp, i64p, i64p, i64p, C.POINTER(C.c_float)]
    for s in SYMBOLS:
        if s != "mskc_last_error":
            getattr(lib, s).restype = C.c_int
    _LIB = lib
    return lib


def _p(a, ty):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Mskc:
    """One KMC context on one GPU (the module globals of defsystem.f, device resident)."""

    def __init__(self, lenx, leny, lenz, listcrystal, listcov=(), device=0, nummcmax=0, slab=None):
        """slab=(rank, nranks): lenx is the GLOBAL LenX of an x-decomposed box; this context holds rank's block plus
        one 24-unit halo tile column on each side (local LenX = lenx/nranks + 48)."""
        self.lib = load_library()
        lc, lv = _i32(list(listcrystal)), _i32(list(listcov) if len(listcov) else [0])
        h = C.c_void_p()
        self.slab = slab
        if slab is not None:
            rank, nranks = slab
            self.global_LenX = lenx
            rc = self.lib.mskc_create_slab(lenx, leny, lenz, len(listcrystal), len(listcov), _p(lc, C.c_int32), _p(lv, C.c_int32), rank, nranks, device, C.byref(h))
            lenx = lenx // nranks + 48
        else:
            rc = self.lib.mskc_create(lenx, leny, lenz, len(listcrystal), len(listcov), _p(lc, C.c_int32), _p(lv, C.c_int32),
                                      device, nummcmax, C.byref(h))
        self.shape = (lenz + 1, leny, lenx)
        self.LenX, self.LenY, self.LenZ = lenx, leny, lenz
        if rc != 0:
            raise MskcError(self.lib.mskc_last_error(None).decode())
        self.h = h

    def _ck(self, rc):
        if rc != 0:
            raise MskcError(self.lib.mskc_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.mskc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- several GPUs ------------------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        if self.lib.mskc_comm_unique_id(buf) != 0:
            raise MskcError(self.lib.mskc_last_error(None).decode())
        return buf.raw

    def comm_connect(self, id128: bytes):
        self._ck(self.lib.mskc_comm_connect(self.h, C.create_string_buffer(id128, 128)))

    # -- tables ------------------------------------------------------------------------------------
    def set_rates(self, PtransE, PtransD, PtransAbs=None, PtransDes=None, PtransZig=1.0):
        pe, pd = _f64(PtransE).reshape(-1), _f64(PtransD).reshape(-1)
        assert pe.size == 12288 and pd.size == 9
        pa = _f64(np.zeros(12) if PtransAbs is None else PtransAbs).reshape(-1)
        ps = _f64(np.zeros(12) if PtransDes is None else PtransDes).reshape(-1)
        self._ck(self.lib.mskc_set_rates(self.h, _p(pe, C.c_double), _p(pd, C.c_double), _p(pa, C.c_double), _p(ps, C.c_double), float(PtransZig)))

    def set_la(self, Tm, P0, sigma, Tflex, tmin, tmax, enable=True):
        self._ck(self.lib.mskc_set_la(self.h, int(enable), Tm, P0, sigma, Tflex, int(tmin), int(tmax)))

    def set_rate_field(self, reftemp, refPE, refPD, refPAbs=None, refPDes=None, PtransZig=1.0, tmin=0, tmax=0, fixed_temperature=0):
        """'ST' mode (SetProbabilityFromField.f / InterpProbFromField.f): ntemp tabulated blocks of the four tables."""
        rt = _i32(reftemp).reshape(-1)
        n = rt.size
        pe, pd = _f64(refPE).reshape(-1), _f64(refPD).reshape(-1)
        assert pe.size == n * 12288 and pd.size == n * 9
        pa = _f64(np.zeros(12 * n) if refPAbs is None else refPAbs).reshape(-1)
        ps = _f64(np.zeros(12 * n) if refPDes is None else refPDes).reshape(-1)
        assert pa.size == n * 12 and ps.size == n * 12
        self._ck(self.lib.mskc_set_rate_field(self.h, n, _p(rt, C.c_int32), _p(pe, C.c_double), _p(pd, C.c_double), _p(pa, C.c_double),
                                              _p(ps, C.c_double), float(PtransZig), int(tmin), int(tmax), int(fixed_temperature)))

    def set_tree_variant(self, variant):
        """TREE_LA: lattice words and event codes cross the boundary in the LaserAnnealing fork's layout."""
        self._ck(self.lib.mskc_set_tree_variant(self.h, int(variant)))

    def set_la_fork(self, Tm, P0, sigma, Tflex, tmin, tmax, InitCountCrystal, num_occ0, num_adatom_occ, enable=True):
        """The LaserAnnealing fork's rate law and loop rules (mulskips-source-LaserAnnealing/mcevent/Get_Prob.f:142-152, main.f:634-726)."""
        def arr(a):
            v = np.zeros(3); a = np.atleast_1d(np.asarray(a, dtype=np.float64)); v[:len(a)] = a; return v
        tm, p0, sg, tf = arr(Tm), arr(P0), arr(sigma), arr(Tflex)
        ic = np.zeros(3, np.int32); ic[:len(InitCountCrystal)] = InitCountCrystal
        self._ck(self.lib.mskc_set_la_fork(self.h, int(enable), _p(tm, C.c_double), _p(p0, C.c_double), _p(sg, C.c_double), _p(tf, C.c_double),
                                           int(tmin), int(tmax), _p(ic, C.c_int32), int(num_occ0), int(num_adatom_occ)))

    # -- state -------------------------------------------------------------------------------------
    def upload_state(self, LattCoo, mc_xyz, mc_nn, mc_event, mc_rate, bulk_xyz, bulk_nn, void_xyz=None, void_nn=None):
        L = _i32(LattCoo).reshape(-1)
        assert L.size == (self.LenZ + 1) * self.LenY * self.LenX
        mx, mn, me, mr = _i32(mc_xyz), _i32(mc_nn), _i32(mc_event), _f64(mc_rate)
        bx, bn = _i32(bulk_xyz), _i32(bulk_nn)
        nv = 0 if void_xyz is None else len(void_xyz)
        vx, vn = (_i32(void_xyz), _i32(void_nn)) if nv else (_i32(np.zeros((1, 3))), _i32(np.zeros((1, 12))))
        self._ck(self.lib.mskc_upload_state(self.h, _p(L, C.c_int32), len(mx), _p(mx, C.c_int32), _p(mn, C.c_int32), _p(me, C.c_int32),
                                            _p(mr, C.c_double), len(bx), _p(bx, C.c_int32), _p(bn, C.c_int32), nv, _p(vx, C.c_int32), _p(vn, C.c_int32)))

    def init_flat(self, len1, walls=True, two_species=False):
        self._ck(self.lib.mskc_init_flat(self.h, int(len1), int(walls), int(two_species)))

    def seed(self, idum, runtype="R", counter0=0, ran3_state=None):
        st = _i32(ran3_state)
        self._ck(self.lib.mskc_seed(self.h, int(idum), ord(runtype), int(counter0), _p(st, C.c_int32)))

    def get_rng(self):
        st = np.zeros(58, np.int32); cnt = C.c_int64(); idum = C.c_int32()
        self._ck(self.lib.mskc_get_rng(self.h, _p(st, C.c_int32), C.byref(cnt), C.byref(idum)))
        return st, cnt.value, idum.value

    def set_clock(self, iter_, time, CountCrystal=None, CountCov=None):
        cc, cv = _i32(CountCrystal), _i32(CountCov)
        self._ck(self.lib.mskc_set_clock(self.h, int(iter_), float(time), _p(cc, C.c_int32), _p(cv, C.c_int32)))

    def set_temperature(self, T):
        t = _i32(T).reshape(-1)
        assert t.size == self.LenZ * self.LenY * self.LenX
        self._ck(self.lib.mskc_set_temperature(self.h, _p(t, C.c_int32)))

    def set_temperature_u16(self, T):
        t = np.ascontiguousarray(T, dtype=np.uint16).reshape(-1)
        assert t.size == self.LenZ * self.LenY * self.LenX
        self._ck(self.lib.mskc_set_temperature_u16(self.h, t.ctypes.data_as(C.POINTER(C.c_uint16))))

    def download_phases(self):
        """la/main.f:806-814: coordination of occupied cells, 10 = wall, 0 elsewhere; [LenZ][LenY][LenX] uint8"""
        p = np.empty((self.LenZ, self.LenY, self.LenX), np.uint8)
        self._ck(self.lib.mskc_download_phases(self.h, p.ctypes.data_as(C.POINTER(C.c_uint8))))
        return p

    def last_dense_ms(self):
        ms = C.c_float()
        self._ck(self.lib.mskc_last_dense_ms(self.h, C.byref(ms)))
        return ms.value

    # -- the loop ----------------------------------------------------------------------------------
    def run_parity(self, max_iter=2100000000, max_time=1e8, exit_on_time=False, out_every=0, out_time=0.0, framecounter=0, event_budget=0) -> ParityResult:
        a = _ParityArgs(int(max_iter), float(max_time), int(exit_on_time), 0, int(out_every), float(out_time), int(framecounter), int(event_budget))
        r = ParityResult()
        self._ck(self.lib.mskc_run_parity(self.h, C.byref(a), C.byref(r)))
        return r

    def run_parallel(self, n_sweeps, events_per_tile=4.0, tau_fixed=0.0, seed=1, flags=0, time_kernels=False) -> ParallelResult:
        """time_kernels: every colour-step launch is timed with CUDA events (n_sweeps <= 64) -> sweep_kernel_ms"""
        a = _ParallelArgs(int(n_sweeps), float(events_per_tile), float(tau_fixed), int(seed), 0, int(flags) | (0x10000 if time_kernels else 0))
        r = ParallelResult()
        self._ck(self.lib.mskc_run_parallel(self.h, C.byref(a), C.byref(r)))
        return r

    def rerate_all(self):
        n = C.c_int64(); tot = C.c_double(); ms = C.c_float()
        self._ck(self.lib.mskc_rerate_all(self.h, C.byref(n), C.byref(tot), C.byref(ms)))
        return n.value, tot.value, ms.value

    # -- read back ---------------------------------------------------------------------------------
    def download_lattice(self):
        L = np.empty(self.shape, np.int32)
        self._ck(self.lib.mskc_download_lattice(self.h, _p(L, C.c_int32)))
        return L

    def patch_lattice(self, cells, words):
        c = np.ascontiguousarray(cells, np.int64); w = _i32(words)
        self._ck(self.lib.mskc_patch_lattice(self.h, len(c), _p(c, C.c_int64), _p(w, C.c_int32)))

    def counters(self) -> Counters:
        c = Counters()
        self._ck(self.lib.mskc_get_counters(self.h, C.byref(c)))
        return c

    def download_lists(self, bulk=False):
        c = self.counters()
        nm, nv = c.num_adatom, c.num_voids
        mx, mn = np.zeros((max(nm, 1), 3), np.int32), np.zeros((max(nm, 1), 4, 3), np.int32)
        me, mr = np.zeros(max(nm, 1), np.int32), np.zeros(max(nm, 1), np.float64)
        vx, vn = np.zeros((max(nv, 1), 3), np.int32), np.zeros((max(nv, 1), 4, 3), np.int32)
        out = dict(mc_xyz=mx[:nm], mc_nn=mn[:nm], mc_event=me[:nm], mc_rate=mr[:nm], void_xyz=vx[:nv], void_nn=vn[:nv])
        if bulk:
            nb = max(c.num_atoms, 1)
            bx, bn = np.zeros((nb, 3), np.int32), np.zeros((nb, 4, 3), np.int32)
            found = C.c_int64()
            self._ck(self.lib.mskc_download_lists(self.h, _p(mx, C.c_int32), _p(mn, C.c_int32), _p(me, C.c_int32), _p(mr, C.c_double),
                                                  _p(vx, C.c_int32), _p(vn, C.c_int32), _p(bx, C.c_int32), _p(bn, C.c_int32), nb, C.byref(found)))
            out.update(bulk_xyz=bx[:found.value], bulk_nn=bn[:found.value], n_bulk_found=found.value)
        else:
            self._ck(self.lib.mskc_download_lists(self.h, _p(mx, C.c_int32), _p(mn, C.c_int32), _p(me, C.c_int32), _p(mr, C.c_double),
                                                  _p(vx, C.c_int32), _p(vn, C.c_int32), None, None, 0, None))
        return out

    def state_hash(self) -> int:
        v = C.c_uint64()
        self._ck(self.lib.mskc_state_hash(self.h, C.byref(v)))
        return v.value

    def tile_stats(self):
        n = C.c_int64()
        self._ck(self.lib.mskc_debug_tile_stats(self.h, None, 0, C.byref(n)))
        dt = np.dtype([("events", "<i8"), ("site_updates", "<i8"), ("dep", "<i4", 3), ("evap", "<i4"), ("abs", "<i4"), ("des", "<i4"), ("dAtoms", "<i4"),
                       ("dVoids", "<i4"), ("dCrystal", "<i4", 3), ("dCov", "<i4", 3), ("maxev", "<i4"), ("pad", "<i4")])
        a = np.zeros(n.value, dt)
        self._ck(self.lib.mskc_debug_tile_stats(self.h, a.ctypes.data, n.value, C.byref(n)))
        return a

    # -- test/diagnostic entry points of the sampler's transition code -------------------------------
    def replay_events(self, site_xyz, event, arm, tree_root=1.0):
        sx, ev, ar = _i32(site_xyz).reshape(-1, 3), _i32(event).reshape(-1), _i32(arm).reshape(-1)
        assert len(sx) == len(ev) == len(ar)
        self._ck(self.lib.mskc_replay_events(self.h, len(ev), _p(sx, C.c_int32), _p(ev, C.c_int32), _p(ar, C.c_int32), float(tree_root)))

    def replay_teams(self, site_xyz, event, arm, tree_root=1.0):
        """the trace through coop_event<8, FEAT>, four teams in lock step; returns (iterations, iterations shared by >= 2 events)"""
        sx, ev, ar = _i32(site_xyz).reshape(-1, 3), _i32(event).reshape(-1), _i32(arm).reshape(-1)
        assert len(sx) == len(ev) == len(ar)
        nit, nsh = C.c_int64(), C.c_int64()
        self._ck(self.lib.mskc_replay_teams(self.h, len(ev), _p(sx, C.c_int32), _p(ev, C.c_int32), _p(ar, C.c_int32), float(tree_root), C.byref(nit), C.byref(nsh)))
        return nit.value, nsh.value

    def download_active(self):
        n = C.c_int64()
        self._ck(self.lib.mskc_download_active(self.h, 0, None, None, None, C.byref(n)))
        cap = max(n.value, 1)
        xyz, ev, rate = np.zeros((cap, 3), np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.float64)
        self._ck(self.lib.mskc_download_active(self.h, cap, _p(xyz, C.c_int32), _p(ev, C.c_int32), _p(rate, C.c_double), C.byref(n)))
        k = min(n.value, cap)
        return xyz[:k], ev[:k], rate[:k]

    def download_bonds(self, xyz):
        x = _i32(xyz).reshape(-1, 3)
        nn = np.zeros((max(len(x), 1), 4, 3), np.int32)
        if len(x):
            self._ck(self.lib.mskc_download_bonds(self.h, len(x), _p(x, C.c_int32), _p(nn, C.c_int32)))
        return nn[:len(x)]

    def extract_frame(self, cap=None):
        """(cells, words, kernel_ms) of every cell the reference's frame writers print, sorted by cell (z,y,x order)"""
        n = C.c_int64(); ms = C.c_float()
        if cap is None:
            self._ck(self.lib.mskc_extract_frame(self.h, 0, None, None, C.byref(n), C.byref(ms)))
            cap = n.value
        cells, words = np.zeros(max(cap, 1), np.uint32), np.zeros(max(cap, 1), np.uint32)
        self._ck(self.lib.mskc_extract_frame(self.h, cap, cells.ctypes.data_as(C.POINTER(C.c_uint32)), words.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(n), C.byref(ms)))
        k = min(n.value, cap)
        o = np.argsort(cells[:k], kind="stable")
        return cells[:k][o], words[:k][o], ms.value

    def _extract(self, fn, *pre):
        u32p = C.POINTER(C.c_uint32)
        n = C.c_int64(); ms = C.c_float()
        self._ck(fn(self.h, *pre, 0, None, None, C.byref(n), C.byref(ms)))
        cap = max(n.value, 1)
        cells, words = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
        self._ck(fn(self.h, *pre, cap, cells.ctypes.data_as(u32p), words.ctypes.data_as(u32p), C.byref(n), C.byref(ms)))
        k = min(n.value, cap)
        o = np.argsort(cells[:k], kind="stable")
        return cells[:k][o], words[:k][o], ms.value

    def extract_frame_all(self):
        """(cells, words, kernel_ms) of every cell any frame writer prints or counts, vacancies included, sorted by cell"""
        return self._extract(self.lib.mskc_extract_frame_all)

    def extract_antisites(self, target):
        """(cells, words, kernel_ms): 4-coordinated atoms of planes 1..LenZ-1 with `target` (0 or 3) neighbours of the other species"""
        return self._extract(self.lib.mskc_extract_antisites, int(target))

    def gather_lattice(self, cells):
        c = np.ascontiguousarray(cells, np.int64)
        w = np.zeros(max(len(c), 1), np.int32)
        if len(c):
            self._ck(self.lib.mskc_gather_lattice(self.h, len(c), c.ctypes.data_as(C.POINTER(C.c_int64)), w.ctypes.data_as(C.POINTER(C.c_int32))))
        return w[:len(c)]

    def check_flags(self):
        """(sites, stored census flags that differ from a fresh census, sites whose coordination differs from the census, ms)"""
        n, bad, coor, ms = C.c_int64(), C.c_int64(), C.c_int64(), C.c_float()
        self._ck(self.lib.mskc_check_flags(self.h, C.byref(n), C.byref(bad), C.byref(coor), C.byref(ms)))
        return n.value, bad.value, coor.value, ms.value

    def census(self):
        c = Census(); ms = C.c_float()
        self._ck(self.lib.mskc_census_lattice(self.h, C.byref(c), C.byref(ms)))
        return c, ms.value
